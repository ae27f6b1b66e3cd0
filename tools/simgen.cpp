// Synthetic HISAT2-style alignment records for benchmarks and large parity cases (TOOLING, not product).
//
// Reads are drawn from known alleles of a synthetic locus and written as the SAM text lines the typing path
// consumes: CIGAR (M/D), MD:Z, NM:i, NH:i and the Zs:Z tag formatted like HISAT2's emitter
// (reference hisat2/sam.h:983-1030: "offset|S/D|id", offset = read bases since the previous item), in the
// name-sorted order `samtools view | sort -k1,1 -s` delivers (reference hisatgenotype_typing_core.py:458-468).
// This mirrors what common.simulate_reads + HISAT2 produce for error-free placement (the simulator's read names
// carry the true alignment, hisatgenotype_typing_common.py:780-843); aligning tens of millions of reads is not
// the subject of the benchmark.  Known variants cost no edit distance (NM counts only sequencing errors).
//
// Build: g++ -O2 -shared -fPIC -o tools/libhgtsim.so tools/simgen.cpp
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9e3779b97f4a7c15ULL + 0x1234567ULL) {}
    uint64_t next() {
        s += 0x9e3779b97f4a7c15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return next() % n; }
};

struct Allele {
    std::string seq;            // allele sequence
    std::vector<int32_t> bpos;  // backbone position of each allele base
    std::vector<int32_t> vid;   // variant index of a `single` at this base, else -1
    std::vector<int32_t> del_after;  // deletion variant index between this base and the next, else -1
};

struct Emit {
    std::string cigar, md, zs;
    int nm = 0;
    int pos1 = 0;  // 1-based leftmost backbone position
};

}  // namespace

struct SimLocus {
    std::string backbone, backbone_name;
    std::vector<Allele> alleles;
    std::vector<std::string> var_ids;
    std::vector<int32_t> var_pos, var_len;
    std::vector<uint8_t> var_type;  // 0 single, 1 deletion
};

static void append_int(std::string &s, long long v) {
    char buf[24];
    snprintf(buf, sizeof(buf), "%lld", v);
    s += buf;
}

static void make_read(const SimLocus &L, const Allele &a, int start, int len, double err, Rng &rng, std::string &read,
                      Emit &e) {
    e.cigar.clear(); e.md.clear(); e.zs.clear();
    e.nm = 0;
    e.pos1 = a.bpos[start] + 1;
    read.resize(len);
    int m_run = 0, md_run = 0;
    int prev_type = -1, prev_pos = 0;  // previous Zs item
    auto zs_item = [&](int type, int rpos, int v) {
        int off = rpos;
        if (prev_type == 0) off = rpos - (prev_pos + 1);
        else if (prev_type == 1) off = rpos - prev_pos;
        if (!e.zs.empty()) e.zs += ',';
        append_int(e.zs, off);
        e.zs += type == 0 ? "|S|" : "|D|";
        e.zs += L.var_ids[v];
        prev_type = type;
        prev_pos = rpos;
    };
    for (int k = 0; k < len; k++) {
        const int i = start + k;
        char r = a.seq[i];
        bool mutated = false;
        if (err > 0.0 && rng.uniform() < err) {
            static const char nt[4] = {'A', 'C', 'G', 'T'};
            char alt;
            do {
                alt = nt[rng.below(4)];
            } while (alt == r);
            r = alt;
            mutated = true;
        }
        read[k] = r;
        const char b = L.backbone[a.bpos[i]];
        m_run++;
        if (r == b) {
            md_run++;
        } else {
            append_int(e.md, md_run);
            e.md += b;
            md_run = 0;
            if (!mutated && a.vid[i] >= 0) zs_item(0, k, a.vid[i]);
            else e.nm++;
        }
        if (k + 1 < len && a.del_after[i] >= 0) {
            const int v = a.del_after[i];
            append_int(e.cigar, m_run);
            e.cigar += 'M';
            m_run = 0;
            append_int(e.cigar, L.var_len[v]);
            e.cigar += 'D';
            append_int(e.md, md_run);
            md_run = 0;
            e.md += '^';
            e.md.append(L.backbone, L.var_pos[v], L.var_len[v]);
            zs_item(1, k + 1, v);
        }
    }
    append_int(e.cigar, m_run);
    e.cigar += 'M';
    append_int(e.md, md_run);
}

extern "C" {

SimLocus *hgtsim_locus_create(const char *backbone, int L, const char *backbone_name, int n_vars, const int32_t *var_pos,
                              const int32_t *var_len, const uint8_t *var_type, const char *var_base, const char *var_ids) {
    SimLocus *s = new SimLocus();
    s->backbone.assign(backbone, L);
    s->backbone_name = backbone_name;
    const char *p = var_ids;
    for (int i = 0; i < n_vars; i++) {
        s->var_pos.push_back(var_pos[i]);
        s->var_len.push_back(var_len[i]);
        s->var_type.push_back(var_type[i]);
        s->var_ids.push_back(p);
        p += s->var_ids.back().size() + 1;
    }
    (void)var_base;
    return s;
}

// add an allele given its (sorted, non-overlapping) variant indices; returns allele slot or -1 (insertions unsupported)
int hgtsim_add_allele(SimLocus *s, int n, const int32_t *vars, const char *var_base) {
    Allele a;
    int prev = 0;
    const int L = (int)s->backbone.size();
    int pending_del = -1;
    auto push = [&](char c, int bp, int vid) {
        if (pending_del >= 0 && !a.seq.empty()) a.del_after.back() = pending_del;
        pending_del = -1;
        a.seq.push_back(c);
        a.bpos.push_back(bp);
        a.vid.push_back(vid);
        a.del_after.push_back(-1);
    };
    for (int k = 0; k < n; k++) {
        const int v = vars[k];
        const int pos = s->var_pos[v];
        if (s->var_type[v] > 1) return -1;
        for (int b = prev; b < pos; b++) push(s->backbone[b], b, -1);
        if (s->var_type[v] == 0) {
            push(var_base[v], pos, v);
            prev = pos + 1;
        } else {
            pending_del = v;
            prev = pos + s->var_len[v];
        }
    }
    for (int b = prev; b < L; b++) push(s->backbone[b], b, -1);
    s->alleles.push_back(std::move(a));
    return (int)s->alleles.size() - 1;
}

void hgtsim_locus_free(SimLocus *s) { delete s; }

// Writes n_pairs read pairs (or single reads when paired == 0) drawn uniformly from the given allele slots.
// Returns bytes written, or -(bytes needed) if out_cap is too small.
long long hgtsim_generate(const SimLocus *s, int n_slots, const int32_t *slots, long long n_pairs, int read_len,
                          int frag_len, double err, uint64_t seed, int paired, long long id_start, const char *id_prefix,
                          char *out, long long out_cap) {
    Rng rng(seed);
    std::string r1, r2, line;
    Emit e1, e2;
    std::string qual((size_t)read_len, 'I');
    long long used = 0;
    bool overflow = false;
    char name[64];
    for (long long p = 0; p < n_pairs; p++) {
        const Allele &a = s->alleles[slots[rng.below((uint64_t)n_slots)]];
        const int alen = (int)a.seq.size();
        const int span = paired ? frag_len : read_len;
        if (alen < span) return 0;
        const int start = (int)rng.below((uint64_t)(alen - span + 1));
        snprintf(name, sizeof(name), "%s%010lld", id_prefix, id_start + p);
        make_read(*s, a, start, read_len, err, rng, r1, e1);
        line.clear();
        auto emit = [&](const Emit &e, const std::string &seq, int flag, int pnext, int tlen) {
            line += name; line += '\t';
            append_int(line, flag); line += '\t';
            line += s->backbone_name; line += '\t';
            append_int(line, e.pos1); line += "\t60\t";
            line += e.cigar; line += '\t';
            if (paired) { line += "=\t"; append_int(line, pnext); line += '\t'; append_int(line, tlen); }
            else line += "*\t0\t0";
            line += '\t'; line += seq; line += '\t'; line += qual;
            line += "\tAS:i:0\tXN:i:0\tXM:i:0\tXO:i:0\tXG:i:0\tNM:i:"; append_int(line, e.nm);
            line += "\tMD:Z:"; line += e.md;
            line += paired ? "\tYS:i:0\tYT:Z:CP\tNH:i:1" : "\tYT:Z:UU\tNH:i:1";
            if (!e.zs.empty()) { line += "\tZs:Z:"; line += e.zs; }
            line += '\n';
        };
        if (paired) {
            make_read(*s, a, start + frag_len - read_len, read_len, err, rng, r2, e2);
            const int tlen = (e2.pos1 - e1.pos1) + read_len;  // approximate template length, unused by typing
            emit(e1, r1, 99, e2.pos1, tlen);
            emit(e2, r2, 147, e1.pos1, -tlen);
        } else {
            emit(e1, r1, 0, 0, 0);
        }
        if (used + (long long)line.size() > out_cap) overflow = true;
        if (!overflow) memcpy(out + used, line.data(), line.size());
        used += (long long)line.size();
    }
    return overflow ? -used : used;
}

}  // extern "C"
