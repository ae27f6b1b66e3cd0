// Stage (a): per-read allele compatibility -> Gene_cmpt / Gene_counts, on the GPU.
//
// Replaces the per-read loop of typing() (reference hisatgenotype_typing_core.py:598-1596) and get_mpileup
// (hisatgenotype_typing_common.py:1059-1184).  Split of work:
//   host  (walk.hpp)   text intake, filters, CIGAR x MD x Zs walk, error correction, ambiguity expansion, exon
//                      clipping  ->  per pair and table a list of haplotypes (left, right, sorted variant rows)
//   GPU   (this file)  pileup histogram + nt_set; haplotype -> allele bitset (add_count, core:626-677);
//                      per-pair arg-max class (add_stat, core:1171-1236); class de-duplication into
//                      (bitset, count, first-seen) tables; Gene_counts.
//
// HBM layout per locus (built once by hgt_locus_create):
//   var_pos[V] int32 (Var_list order), delr_{right,row}[D] deletions sorted by right end,
//   st[levels][V][wp] uint64: level 0 = Links as a bit matrix (one row per variant, one bit per allele),
//   level j = OR of 2^j consecutive rows (sparse table: OR is idempotent, so any contiguous range of
//   variant rows is the OR of two rows) -> a haplotype with k known variants touches 3k+2 rows instead of
//   every row its span covers.  mask[3][wp] = table masks (all alleles / exon reps / primary-exon reps).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "walk.hpp"

using namespace hgt;

// ================================================================================================================
// Locus
// ================================================================================================================
struct hgt_locus {
    LocusHost host;
    int A = 0, V = 0, L = 0, wp = 0, levels = 0, n_delr = 0;
    bool is_hla = false;
    std::vector<uint64_t> mask;  // [3][wp]
    std::vector<int32_t> gn_rank;
    std::vector<int32_t> delr_right, delr_row;
    hgt_ctx *ctx = nullptr;
    int32_t *d_var_pos = nullptr, *d_delr_right = nullptr, *d_delr_row = nullptr, *d_gn_rank = nullptr;
    uint64_t *d_st = nullptr, *d_mask = nullptr;
};

struct LocusDev {
    const int32_t *var_pos, *delr_right, *delr_row;
    const uint64_t *st, *mask;
    int V, wp, n_delr, levels;
};

static LocusDev locus_dev(const hgt_locus *l) {
    LocusDev d;
    d.var_pos = l->d_var_pos; d.delr_right = l->d_delr_right; d.delr_row = l->d_delr_row;
    d.st = l->d_st; d.mask = l->d_mask; d.V = l->V; d.wp = l->wp; d.n_delr = l->n_delr; d.levels = l->levels;
    return d;
}

__global__ void st_build_kernel(uint64_t *st, int V, int wp, int level) {
    const size_t n = (size_t)V * wp;
    const uint64_t *src = st + (size_t)(level - 1) * n;
    uint64_t *dst = st + (size_t)level * n;
    const int half = 1 << (level - 1);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / wp), j = (int)(i % wp);
        const int r2 = row + half;
        uint64_t v = src[i];
        if (r2 < V) v |= src[(size_t)r2 * wp + j];
        dst[i] = v;
    }
}

static bool split_tokens(const std::string &s, char sep, std::vector<std::string> *out) {
    out->clear();
    size_t p = 0;
    while (true) {
        size_t q = s.find(sep, p);
        if (q == std::string::npos) {
            out->push_back(s.substr(p));
            break;
        }
        out->push_back(s.substr(p, q - p));
        p = q + 1;
    }
    return true;
}

static int parse_alts(hgt_locus *loc, const char *text) {
    if (!text) return HGT_OK;
    const char *p = text;
    std::vector<std::string> toks;
    while (*p) {
        const char *e = strchr(p, '\n');
        if (!e) e = p + strlen(p);
        std::string line(p, e - p);
        p = *e ? e + 1 : e;
        if (line.empty()) continue;
        std::vector<std::string> cols;
        split_tokens(line, '\t', &cols);
        if (cols.size() != 3 || (cols[0] != "L" && cols[0] != "R")) {
            hgt_set_error("alts_text: malformed line '%s'", line.c_str());
            return HGT_ERR_ARG;
        }
        AltEntry ent;
        ent.key = cols[1];
        split_tokens(ent.key, '-', &ent.toks);
        if (ent.toks.size() < 2) {
            hgt_set_error("alts_text: malformed key '%s'", ent.key.c_str());
            return HGT_ERR_ARG;
        }
        ent.anchor = atoi(cols[0] == "L" ? ent.toks.back().c_str() : ent.toks.front().c_str());
        std::vector<std::string> alts;
        split_tokens(cols[2], ',', &alts);
        for (const std::string &a : alts) {
            split_tokens(a, '-', &toks);
            if (toks.size() < 2) {
                hgt_set_error("alts_text: malformed alternative '%s'", a.c_str());
                return HGT_ERR_ARG;
            }
            AltHap h;
            h.left = atoi(toks.front().c_str());
            h.right = atoi(toks.back().c_str());
            for (size_t i = 1; i + 1 < toks.size(); i++) {
                auto f = loc->host.row_of.find(toks[i]);
                if (f == loc->host.row_of.end()) {
                    hgt_set_error("alts_text: unknown variant id '%s'", toks[i].c_str());
                    return HGT_ERR_ARG;
                }
                if (loc->host.vars[f->second].type == T_INSERTION) {
                    hgt_set_error("alts_text: insertion '%s' inside an alternative haplotype", toks[i].c_str());
                    return HGT_ERR_ARG;
                }
                h.rows.push_back(f->second);
            }
            ent.alts.push_back(std::move(h));
        }
        (cols[0] == "L" ? loc->host.alts_left : loc->host.alts_right).push_back(std::move(ent));
    }
    auto by_anchor = [](const AltEntry &a, const AltEntry &b) { return a.anchor < b.anchor; };
    std::stable_sort(loc->host.alts_left.begin(), loc->host.alts_left.end(), by_anchor);
    std::stable_sort(loc->host.alts_right.begin(), loc->host.alts_right.end(), by_anchor);
    return HGT_OK;
}

extern "C" void hgt_locus_free(hgt_locus *l) {
    if (!l) return;
    if (l->ctx) {
        cudaSetDevice(l->ctx->device);
        cudaFree(l->d_var_pos); cudaFree(l->d_delr_right); cudaFree(l->d_delr_row); cudaFree(l->d_gn_rank);
        cudaFree(l->d_st); cudaFree(l->d_mask);
    }
    delete l;
}

extern "C" int hgt_locus_create(hgt_ctx *ctx, const hgt_locus_desc *d, hgt_locus **out) {
    if (!d || !out) {
        hgt_set_error("hgt_locus_create: null argument");
        return HGT_ERR_ARG;
    }
    *out = nullptr;
    if (d->n_alleles < 1 || d->n_vars < 0 || d->ref_len < 1 || !d->ref_seq) {
        hgt_set_error("hgt_locus_create: need n_alleles >= 1, ref_len >= 1");
        return HGT_ERR_ARG;
    }
    hgt_locus *l = new hgt_locus();
    l->ctx = ctx;
    l->A = d->n_alleles; l->V = d->n_vars; l->L = d->ref_len; l->wp = hgt_row_pitch(d->n_alleles);
    l->is_hla = d->is_hla != 0;
    LocusHost &h = l->host;
    h.is_hla = l->is_hla;
    h.ref.assign(d->ref_seq, d->ref_len);
    h.vars.resize(l->V);
    h.var_pos.resize(l->V);
    const char *ids = d->var_ids;
    for (int i = 0; i < l->V; i++) {
        VarInfo &v = h.vars[i];
        v.pos = d->var_pos[i]; v.len = d->var_len[i]; v.type = d->var_type[i]; v.base = d->var_base[i];
        v.in_links = d->var_flags[i] & 1; v.is_hv = (d->var_flags[i] >> 1) & 1;
        v.id = ids;
        ids += v.id.size() + 1;
        h.var_pos[i] = v.pos;
        h.row_of[v.id] = i;
        if (i > 0 && h.var_pos[i] < h.var_pos[i - 1]) {
            hgt_set_error("hgt_locus_create: variants must be sorted by position (Var_list order)");
            delete l;
            return HGT_ERR_ARG;
        }
    }
    for (int i = 0; i < d->n_exons; i++) h.exons.push_back({d->exons[2 * i], d->exons[2 * i + 1]});
    for (int i = 0; i < d->n_primary_exons; i++)
        h.primary_exons.push_back({d->primary_exons[2 * i], d->primary_exons[2 * i + 1]});
    int rc = parse_alts(l, d->alts_text);
    if (rc != HGT_OK) {
        delete l;
        return rc;
    }
    const int wp = l->wp;
    l->mask.assign((size_t)3 * wp, 0);
    for (int a = 0; a < l->A; a++) l->mask[a >> 6] |= 1ull << (a & 63);
    if (l->is_hla) {
        if (!d->exon_rep_mask || !d->primary_rep_mask) {
            hgt_set_error("hgt_locus_create: is_hla needs exon_rep_mask and primary_rep_mask");
            delete l;
            return HGT_ERR_ARG;
        }
        for (int j = 0; j < wp; j++) {
            l->mask[wp + j] = d->exon_rep_mask[j] & l->mask[j];
            l->mask[2 * wp + j] = d->primary_rep_mask[j] & l->mask[j];
        }
    }
    l->gn_rank.assign(l->A, 0);
    for (int a = 0; a < l->A; a++) l->gn_rank[a] = d->gene_names_rank ? d->gene_names_rank[a] : a;
    // deletions sorted by right end: a deletion starting left of a haplotype but ending inside it is negative
    // evidence too (core:663-669)
    std::vector<int32_t> order;
    for (int i = 0; i < l->V; i++)
        if (h.vars[i].type == T_DELETION && h.vars[i].in_links && h.vars[i].len > 1) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h.vars[a].right() < h.vars[b].right(); });
    for (int r : order) {
        l->delr_right.push_back(h.vars[r].right());
        l->delr_row.push_back(r);
    }
    l->n_delr = (int)order.size();
    if (!ctx) {
        *out = l;
        return HGT_OK;
    }
    // ---- device tables -----------------------------------------------------------------------------------
    cudaError_t e = cudaSetDevice(ctx->device);
    const int V = l->V;
    int levels = 1;
    while ((1 << levels) <= V) levels++;
    l->levels = levels;
    std::vector<uint64_t> lbits((size_t)std::max(V, 1) * wp, 0);
    for (int i = 0; i < V; i++) {
        if (!h.vars[i].in_links) continue;
        for (int64_t k = d->link_off[i]; k < d->link_off[i + 1]; k++) {
            const int a = d->link_allele[k];
            if (a < 0 || a >= l->A) {
                hgt_set_error("hgt_locus_create: link allele index %d out of range", a);
                delete l;
                return HGT_ERR_ARG;
            }
            lbits[(size_t)i * wp + (a >> 6)] |= 1ull << (a & 63);
        }
    }
    const size_t lvl_words = (size_t)std::max(V, 1) * wp;
#define LTRY(call)                                                                                 \
    if (e == cudaSuccess) {                                                                        \
        e = (call);                                                                                \
        if (e != cudaSuccess) hgt_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e)); \
    }
    LTRY(cudaMalloc(&l->d_var_pos, sizeof(int32_t) * std::max(V, 1)));
    LTRY(cudaMalloc(&l->d_delr_right, sizeof(int32_t) * std::max(l->n_delr, 1)));
    LTRY(cudaMalloc(&l->d_delr_row, sizeof(int32_t) * std::max(l->n_delr, 1)));
    LTRY(cudaMalloc(&l->d_gn_rank, sizeof(int32_t) * l->A));
    LTRY(cudaMalloc(&l->d_mask, sizeof(uint64_t) * 3 * wp));
    LTRY(cudaMalloc(&l->d_st, sizeof(uint64_t) * lvl_words * levels));
    cudaStream_t st = ctx->stream;
    if (V > 0) LTRY(cudaMemcpyAsync(l->d_var_pos, h.var_pos.data(), sizeof(int32_t) * V, cudaMemcpyHostToDevice, st));
    if (l->n_delr > 0) {
        LTRY(cudaMemcpyAsync(l->d_delr_right, l->delr_right.data(), sizeof(int32_t) * l->n_delr, cudaMemcpyHostToDevice, st));
        LTRY(cudaMemcpyAsync(l->d_delr_row, l->delr_row.data(), sizeof(int32_t) * l->n_delr, cudaMemcpyHostToDevice, st));
    }
    LTRY(cudaMemcpyAsync(l->d_gn_rank, l->gn_rank.data(), sizeof(int32_t) * l->A, cudaMemcpyHostToDevice, st));
    LTRY(cudaMemcpyAsync(l->d_mask, l->mask.data(), sizeof(uint64_t) * 3 * wp, cudaMemcpyHostToDevice, st));
    LTRY(cudaMemcpyAsync(l->d_st, lbits.data(), sizeof(uint64_t) * lvl_words, cudaMemcpyHostToDevice, st));
    if (e == cudaSuccess) {
        for (int lv = 1; lv < levels; lv++) {
            st_build_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(l->d_st, V, wp, lv);
            ctx->launches++;
        }
        LTRY(cudaGetLastError());
    }
    LTRY(cudaStreamSynchronize(st));
#undef LTRY
    if (e != cudaSuccess) {
        hgt_locus_free(l);
        return HGT_ERR_CUDA;
    }
    *out = l;
    return HGT_OK;
}

// ================================================================================================================
// Host pipeline: alignment text -> per-pair haplotype jobs
// ================================================================================================================
struct TableJobs {
    std::vector<int64_t> job_off{0};  // per pair: range of haplotypes
    std::vector<int32_t> hap_left, hap_right;
    std::vector<int64_t> row_off{0};
    std::vector<int32_t> rows;
    void add_hap(const LocusHost &L, const Haplotype &h, std::vector<int32_t> &tmp) {
        hap_left.push_back(h.left);
        hap_right.push_back(h.right);
        tmp.clear();
        for (int32_t id : h.ids)
            if (id >= 0 && L.vars[id].in_links) tmp.push_back(id);
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        rows.insert(rows.end(), tmp.begin(), tmp.end());
        row_off.push_back((int64_t)rows.size());
    }
    void end_job() { job_off.push_back((int64_t)hap_left.size()); }
};

struct HostOut {
    int64_t num_reads = 0, num_pairs = 0;
    TableJobs tb[3];
};

struct PileupIn {  // records that take part in the pileup (weaker filters, common:1084-1098)
    std::vector<int32_t> pos;
    std::vector<int64_t> cig_off{0}, seq_off{0};
    std::vector<uint32_t> cig;  // len << 4 | op (0 M, 1 I, 2 D, 3 S, 4 N)
    std::vector<char> seq;
};

struct Intake {
    std::vector<Record> recs;
};

static int intake(const char *sam, size_t n, const hgt_params &pr, Intake *in) {
    const char *p = sam, *end = sam + n;
    while (p < end) {
        const char *q = (const char *)memchr(p, '\n', end - p);
        if (!q) q = end;
        const char *a = p;
        while (a < q && is_ws(*a)) a++;
        if (a < q && *a != '@') {
            Record r;
            parse_record(a, q, pr.simulation != 0, pr.base_locus, &r);
            if (!r.ok) {
                hgt_set_error("malformed alignment record (fewer than 11 columns or bad integer): %.60s", a);
                return HGT_ERR_PARSE;
            }
            in->recs.push_back(r);
        }
        p = q + 1;
    }
    return HGT_OK;
}

static int op_code(char c) {
    switch (c) {
        case 'M': return 0;
        case 'I': return 1;
        case 'D': return 2;
        case 'S': return 3;
        case 'N': return 4;
        default: return -1;
    }
}

static int build_pileup_input(const Intake &in, const hgt_params &pr, PileupIn *pi) {
    std::vector<CigarOp> cig;
    for (const Record &r : in.recs) {
        if (r.flag & 0x4) continue;
        if (r.pos < 0) continue;
        if (!pr.allow_discordant && !(r.flag & 0x2)) continue;
        if (!parse_cigar(r.cigar, r.cigar_len, &cig)) {
            hgt_set_error("malformed CIGAR in read %.*s", r.qname_len, r.qname);
            return HGT_ERR_PARSE;
        }
        pi->pos.push_back(r.pos);
        for (const CigarOp &c : cig) {
            const int oc = op_code(c.op);
            if (oc < 0) continue;  // the reference's pileup ignores ops outside MIDNS (common:1107-1121)
            pi->cig.push_back(((uint32_t)c.len << 4) | (uint32_t)oc);
        }
        pi->cig_off.push_back((int64_t)pi->cig.size());
        pi->seq.insert(pi->seq.end(), r.seq, r.seq + r.seq_len);
        pi->seq_off.push_back((int64_t)pi->seq.size());
    }
    return HGT_OK;
}

static int host_walk(const hgt_locus *loc, const Intake &in, const hgt_params &pr, const PileupView &pu, HostOut *out) {
    const LocusHost &L = loc->host;
    std::unordered_set<std::string_view> seen[3];
    std::vector<NovelVar> novel;
    std::vector<Haplotype> left_hts, right_hts, all_hts, exon_tmp;
    std::vector<int32_t> tmp_rows;
    std::string seq;
    std::vector<CigarOp> cig;
    std::vector<ZsItem> zs;
    WalkResult w;
    WalkError err;
    Ambig amb;
    std::vector<Cmp> c2;
    std::string_view prev_id;
    bool have_prev = false;
    auto get_var = [&](int32_t id) {
        VarLite v;
        if (id >= 0) {
            v.type = L.vars[id].type; v.pos = L.vars[id].pos; v.len = L.vars[id].len;
        } else {
            const NovelVar &nv = novel[novel_index(id)];
            v.type = nv.type; v.pos = nv.pos; v.len = nv.len;
        }
        return v;
    };
    auto add_unique_ht = [](std::vector<Haplotype> &set, Haplotype &&h) {
        for (const Haplotype &x : set)
            if (x == h) return;
        set.push_back(std::move(h));
    };
    auto flush = [&]() {
        all_hts.clear();
        for (Haplotype &h : left_hts) add_unique_ht(all_hts, std::move(h));
        for (Haplotype &h : right_hts) add_unique_ht(all_hts, std::move(h));
        for (const Haplotype &h : all_hts) {
            if (loc->is_hla) {
                exon_tmp.clear();
                exon_haplotypes(h, L.primary_exons, get_var, &exon_tmp);
                for (const Haplotype &e : exon_tmp) out->tb[2].add_hap(L, e, tmp_rows);
                exon_tmp.clear();
                exon_haplotypes(h, L.exons, get_var, &exon_tmp);
                for (const Haplotype &e : exon_tmp) out->tb[1].add_hap(L, e, tmp_rows);
            }
            out->tb[0].add_hap(L, h, tmp_rows);
        }
        for (int t = 0; t < 3; t++) out->tb[t].end_job();
        out->num_pairs++;
        left_hts.clear();
        right_hts.clear();
    };
    for (const Record &r : in.recs) {
        if (r.pos < 0) continue;
        if (r.flag & 0x4) continue;
        if (!r.has_NM || !r.has_NH) {
            hgt_set_error("alignment of read %.*s lacks the NM or NH tag", r.qname_len, r.qname);
            return HGT_ERR_PARSE;
        }
        if (r.NM > pr.num_editdist) continue;
        if (r.NH > 1) continue;
        if (!pr.allow_discordant && !(r.flag & 0x2)) continue;
        const bool is_left = (r.flag & 0x40) != 0;
        const int kind = is_left ? 0 : ((r.flag & 0x80) ? 1 : 2);
        if (kind == 2 && !pr.allow_discordant) {
            hgt_set_error("read %.*s is neither first nor second mate and --discordant is off", r.qname_len, r.qname);
            return HGT_ERR_PARSE;
        }
        const std::string_view id(r.qname, r.qname_len);
        if (!seen[kind].insert(id).second) continue;
        if (!walk_record(L, r, pu, pr.error_correction != 0, seq, cig, zs, &w, &err)) {
            hgt_set_error("%s", err.msg.c_str());
            return err.code;
        }
        if (w.right_pos > (int32_t)L.ref.size()) continue;
        if (w.ncorr > std::max(1, pr.num_editdist)) continue;
        if (w.misaligned) continue;
        // novel variants (core:1126-1164): only indels keep an identity; it is (type, pos, len)
        for (Cmp &e : w.cmp) {
            if ((e.type == C_INSERTION || e.type == C_DELETION) && e.var == VAR_UNKNOWN) {
                const NovelVar nv{(uint8_t)(e.type == C_INSERTION ? T_INSERTION : T_DELETION), e.pos, e.len};
                size_t k = 0;
                for (; k < novel.size(); k++)
                    if (novel[k] == nv) break;
                if (k == novel.size()) novel.push_back(nv);
                e.var = VAR_NOVEL_BASE - (int32_t)k;
            }
        }
        out->num_reads++;
        if (!have_prev || id != prev_id) {
            if (have_prev) flush();
            left_hts.clear();
            right_hts.clear();
        }
        // cmp_list2 (core:1351-1368)
        c2.clear();
        for (const Cmp &e : w.cmp) {
            if (e.type == C_MATCH || (e.type == C_MISMATCH && e.var < 0)) {
                const int32_t ln = e.type == C_MATCH ? e.len : 1;
                if (!c2.empty() && c2.back().type == C_MATCH) c2.back().len += ln;
                else c2.push_back({C_MATCH, e.pos, ln, -1});
            } else {
                c2.push_back(e);
            }
        }
        if (c2.empty()) {
            hgt_set_error("read %.*s has an empty alignment", r.qname_len, r.qname);
            return HGT_ERR_PARSE;
        }
        if (!identify_ambiguous_diffs(L, c2, &amb, &err)) {
            hgt_set_error("%s (read %.*s)", err.msg.c_str(), r.qname_len, r.qname);
            return err.code;
        }
        std::vector<Haplotype> &dst = is_left ? left_hts : right_hts;
        for (const AltSide &a : amb.left) {
            for (const AltSide &b : amb.right) {
                Haplotype h;
                h.left = a.pos;
                h.right = b.pos;
                h.ids = a.ids;
                for (int32_t i = amb.cmp_left; i <= amb.cmp_right; i++)
                    if (c2[i].type != C_MATCH) h.ids.push_back(c2[i].var);
                h.ids.insert(h.ids.end(), b.ids.begin(), b.ids.end());
                add_unique_ht(dst, std::move(h));
            }
        }
        prev_id = id;
        have_prev = true;
    }
    if (have_prev) flush();
    return HGT_OK;
}

// ================================================================================================================
// Kernels
// ================================================================================================================
namespace {

constexpr int WARPS_PER_CTA = 8;

// ---- pileup (common:1100-1121): warp per record, lanes over the bases of each CIGAR op -------------------
__global__ void pileup_kernel(const int32_t *__restrict__ pos, const int64_t *__restrict__ cig_off,
                              const uint32_t *__restrict__ cig, const int64_t *__restrict__ seq_off,
                              const char *__restrict__ seq, int64_t n_rec, int L, uint32_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rec; r += nwarps) {
        int gpos = pos[r];
        int64_t rpos = seq_off[r];
        for (int64_t c = cig_off[r]; c < cig_off[r + 1]; c++) {
            const uint32_t x = cig[c];
            const int len = (int)(x >> 4), op = (int)(x & 15u);
            if (op == 0 || op == 2) {
                for (int j = lane; j < len; j += 32) {
                    const int g = gpos + j;
                    if (g < L) {
                        int code = 5;
                        if (op == 0) {
                            const char ch = seq[rpos + j];
                            code = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4;
                        }
                        atomicAdd(&counts[(size_t)g * 6 + code], 1u);
                    }
                }
            }
            if (op == 0 || op == 2 || op == 4) gpos += len;
            if (op == 0 || op == 1 || op == 3) rpos += len;
        }
    }
}

// representative bases (common:1124-1134): depth >= 20 and (count >= depth*0.2 or count >= 7)
__global__ void ntset_kernel(const uint32_t *__restrict__ counts, int L, uint8_t *__restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const uint32_t *c = counts + (size_t)i * 6;
    const uint32_t depth = c[0] + c[1] + c[2] + c[3] + c[4] + c[5];
    uint8_t m = 0;
    if (depth >= 20) {
        const double thr = (double)depth * 0.2;  // same IEEE product as Python's num_nt * 0.2
        for (int k = 0; k < 4; k++)
            if ((double)c[k] >= thr || c[k] >= 7) m |= (uint8_t)(1u << k);
    }
    mask[i] = m;
}

__device__ __forceinline__ int lower_bound_dev(const int32_t *a, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// ---- haplotype -> allele bitset (add_count, core:626-677; set form in SURVEY.md appendix A.5) ----------------
// warp per haplotype; lane l owns words l, l+32, ... of the row (WPL words per lane).
template <int WPL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
    compat_kernel(LocusDev loc, int table, const int32_t *__restrict__ hap_left, const int32_t *__restrict__ hap_right,
                  const int64_t *__restrict__ row_off, const int32_t *__restrict__ rows, int64_t n_haps,
                  uint64_t *__restrict__ out) {
    extern __shared__ __align__(16) int32_t s_pos[];  // var_pos staged once per CTA (persistent grid)
    for (int i = threadIdx.x; i < loc.V; i += blockDim.x) s_pos[i] = loc.var_pos[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wp = loc.wp;
    const size_t lvl = (size_t)max(loc.V, 1) * wp;
    const uint64_t *mask = loc.mask + (size_t)table * wp;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t h = warp0; h < n_haps; h += nwarps) {
        const int left = hap_left[h], right = hap_right[h];
        const int64_t r0 = row_off[h], r1 = row_off[h + 1];
        uint64_t acc[WPL], neg[WPL];
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const int j = lane + 32 * i;
            acc[i] = j < wp ? mask[j] : 0ull;
            neg[i] = 0ull;
        }
        // positives
        for (int64_t k = r0; k < r1; k++) {
            const uint64_t *row = loc.st + (size_t)rows[k] * wp;
#pragma unroll
            for (int i = 0; i < WPL; i++) {
                const int j = lane + 32 * i;
                if (j < wp) acc[i] &= row[j];
            }
        }
        // negatives with the left end inside [left, right]: rows [lo, hi) minus the haplotype's own rows
        const int lo = lower_bound_dev(s_pos, loc.V, left), hi = lower_bound_dev(s_pos, loc.V, right + 1);
        int prev = lo;
        for (int64_t k = r0; k <= r1; k++) {
            int endr = hi;
            if (k < r1) {
                endr = rows[k];
                if (endr < lo) continue;
                if (endr > hi) endr = hi;
            }
            if (endr > prev) {
                const int n = endr - prev;
                const int lv = 31 - __clz(n);
                const uint64_t *ra = loc.st + lv * lvl + (size_t)prev * wp;
                const uint64_t *rb = loc.st + lv * lvl + (size_t)(endr - (1 << lv)) * wp;
#pragma unroll
                for (int i = 0; i < WPL; i++) {
                    const int j = lane + 32 * i;
                    if (j < wp) neg[i] |= ra[j] | rb[j];
                }
            }
            prev = endr + 1;
            if (prev >= hi) break;
        }
        // deletions that start left of the haplotype and end inside it
        if (loc.n_delr > 0) {
            const int dlo = lower_bound_dev(loc.delr_right, loc.n_delr, left);
            const int dhi = lower_bound_dev(loc.delr_right, loc.n_delr, right + 1);
            for (int dd = dlo; dd < dhi; dd++) {
                const int row = loc.delr_row[dd];
                if (s_pos[row] >= left) continue;
                bool own = false;
                for (int64_t k = r0; k < r1; k++) own |= rows[k] == row;
                if (own) continue;
                const uint64_t *rp = loc.st + (size_t)row * wp;
#pragma unroll
                for (int i = 0; i < WPL; i++) {
                    const int j = lane + 32 * i;
                    if (j < wp) neg[i] |= rp[j];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const int j = lane + 32 * i;
            if (j < wp) out[(size_t)h * wp + j] = acc[i] & ~neg[i];
        }
    }
}

// ---- per-pair class (add_stat, core:1171-1236) + de-duplication --------------------------------------------
struct ClassPool {
    unsigned long long *keys;  // open-addressing table, 0 = empty
    int32_t *slot_class;       // class id once its row is written, -1 before
    uint32_t cap_mask;
    uint64_t *bits;            // [max_classes][wp]
    unsigned long long *count;
    int32_t *first;            // first pair index
    int32_t *table;            // table id of the class
    int32_t *n_classes;
    int32_t max_classes;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// warp per job; a job = (pair, table) with a list of haplotype bitsets; P bit-planes count up to 2^P-1
template <int WPL, int P>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
    class_kernel(int wp, const uint64_t *__restrict__ mask, int table, const int64_t *__restrict__ job_off,
                 const int32_t *__restrict__ job_list, int64_t n_jobs, const uint64_t *__restrict__ hapbits,
                 ClassPool pool) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp0; q < n_jobs; q += nwarps) {
        const int pair = job_list ? job_list[q] : (int)q;
        const int64_t h0 = job_off[pair], h1 = job_off[pair + 1];
        uint64_t plane[P][WPL], best[WPL];
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int i = 0; i < WPL; i++) plane[p][i] = 0ull;
        for (int64_t h = h0; h < h1; h++) {
#pragma unroll
            for (int i = 0; i < WPL; i++) {
                const int j = lane + 32 * i;
                uint64_t carry = j < wp ? hapbits[(size_t)h * wp + j] : 0ull;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const uint64_t t = plane[p][i] & carry;
                    plane[p][i] ^= carry;
                    carry = t;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const int j = lane + 32 * i;
            best[i] = j < wp ? mask[j] : 0ull;
        }
#pragma unroll
        for (int p = P - 1; p >= 0; p--) {
            uint64_t any = 0;
#pragma unroll
            for (int i = 0; i < WPL; i++) any |= best[i] & plane[p][i];
            if (__any_sync(0xffffffffu, any != 0ull)) {
#pragma unroll
                for (int i = 0; i < WPL; i++) best[i] &= plane[p][i];
            }
        }
        // hash of the class bitset
        uint64_t hsh = 0;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const int j = lane + 32 * i;
            if (j < wp) hsh += mix64(best[i] + 0x9e3779b97f4a7c15ULL * (uint64_t)(j + 1));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hsh += __shfl_xor_sync(0xffffffffu, hsh, o);
        const unsigned long long key = mix64(hsh + 0x632be59bd9b4e019ULL * (uint64_t)(table + 1)) | 1ull;
        uint32_t slot = (uint32_t)(key >> 20) & pool.cap_mask;
        int cid = -1;
        while (true) {
            unsigned long long prev = 0;
            if (lane == 0) prev = atomicCAS(&pool.keys[slot], 0ull, key);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == 0ull) {  // we own the slot: publish a new class
                int c = 0;
                if (lane == 0) c = atomicAdd(pool.n_classes, 1);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c < pool.max_classes) {
#pragma unroll
                    for (int i = 0; i < WPL; i++) {
                        const int j = lane + 32 * i;
                        if (j < wp) pool.bits[(size_t)c * wp + j] = best[i];
                    }
                    if (lane == 0) pool.table[c] = table;
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) atomicExch(&pool.slot_class[slot], c);
                }
                cid = c;
                break;
            }
            if (prev == key) {
                int c = -1;
                if (lane == 0) {
                    while ((c = *((volatile int32_t *)&pool.slot_class[slot])) < 0) {
                    }
                    __threadfence();
                }
                c = __shfl_sync(0xffffffffu, c, 0);
                bool same = __ldcg(&pool.table[c]) == table;
#pragma unroll
                for (int i = 0; i < WPL; i++) {
                    const int j = lane + 32 * i;
                    if (j < wp) same &= (__ldcg(&pool.bits[(size_t)c * wp + j]) == best[i]);
                }
                if (__all_sync(0xffffffffu, same)) {
                    cid = c;
                    break;
                }
            }
            slot = (slot + 1) & pool.cap_mask;
        }
        if (lane == 0 && cid < pool.max_classes) {
            atomicAdd(&pool.count[cid], 1ull);
            atomicMin(&pool.first[cid], pair);
        }
    }
}

// Gene_counts (core:1187-1190): counts[a] = sum of class counts over classes holding a; first[a] = first pair
__global__ void table_counts_kernel(int A, int wp, const int32_t *__restrict__ cls_idx, int n_cls,
                                    const uint64_t *__restrict__ bits, const unsigned long long *__restrict__ count,
                                    const int32_t *__restrict__ first, long long *__restrict__ a_count,
                                    long long *__restrict__ a_first) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    long long c = 0, f = 0x7fffffffffffffffLL;
    for (int k = 0; k < n_cls; k++) {
        const int id = cls_idx[k];
        const uint64_t w = bits[(size_t)id * wp + (a >> 6)];
        if ((w >> (a & 63)) & 1ull) {
            c += (long long)count[id];
            f = min(f, (long long)first[id]);
        }
    }
    a_count[a] = c;
    a_first[a] = c ? f : -1;
}

__global__ void gather_rows_kernel(int wp, const int32_t *__restrict__ idx, int n, const uint64_t *__restrict__ src,
                                   uint64_t *__restrict__ dst) {
    const size_t total = (size_t)n * wp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / wp), j = (int)(i % wp);
        dst[i] = src[(size_t)idx[r] * wp + j];
    }
}

}  // namespace

// ================================================================================================================
// Device arena helper
// ================================================================================================================
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n) {
        bytes = n;
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            hgt_set_error("cudaMalloc(%zu) -> %s", n, cudaGetErrorString(e));
            p = nullptr;
            return e == cudaErrorMemoryAllocation ? HGT_ERR_NOMEM : HGT_ERR_CUDA;
        }
        return HGT_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
    }
    template <class T>
    T *as() const { return static_cast<T *>(p); }
};

template <class T>
static int upload(DevBuf *b, const std::vector<T> &v, cudaStream_t st) {
    HGT_CHECK(b->alloc(v.size() * sizeof(T)));
    if (!v.empty()) HGT_CUDA(cudaMemcpyAsync(b->p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return HGT_OK;
}

// ================================================================================================================
// Typing result
// ================================================================================================================
struct hgt_typing {
    hgt_ctx *ctx = nullptr;
    const hgt_locus *locus = nullptr;
    int64_t num_reads = 0, num_pairs = 0;
    int32_t n_classes[3] = {0, 0, 0};
    std::vector<int32_t> cls_idx[3];  // class ids of each table in first-seen order
    std::vector<int64_t> cls_count[3], cls_first[3];
    std::vector<uint32_t> pile_counts;
    std::vector<uint8_t> pile_mask;
    DevBuf d_bits, d_count, d_first, d_table;  // class pool
    int32_t total_classes = 0;
};

extern "C" void hgt_typing_free(hgt_typing *t) {
    if (!t) return;
    if (t->ctx) cudaSetDevice(t->ctx->device);
    t->d_bits.release(); t->d_count.release(); t->d_first.release(); t->d_table.release();
    delete t;
}

static int run_pileup(hgt_ctx *ctx, const hgt_locus *loc, const PileupIn &pi, std::vector<uint32_t> *counts,
                      std::vector<uint8_t> *mask) {
    cudaStream_t st = ctx->stream;
    const int L = loc->L;
    counts->assign((size_t)L * 6, 0);
    mask->assign(L, 0);
    DevBuf d_pos, d_co, d_c, d_so, d_s, d_cnt, d_m;
    int rc = HGT_OK;
    do {
        if ((rc = upload(&d_pos, pi.pos, st)) != HGT_OK) break;
        if ((rc = upload(&d_co, pi.cig_off, st)) != HGT_OK) break;
        if ((rc = upload(&d_c, pi.cig, st)) != HGT_OK) break;
        if ((rc = upload(&d_so, pi.seq_off, st)) != HGT_OK) break;
        if ((rc = upload(&d_s, pi.seq, st)) != HGT_OK) break;
        if ((rc = d_cnt.alloc((size_t)L * 6 * 4)) != HGT_OK) break;
        if ((rc = d_m.alloc(L)) != HGT_OK) break;
        cudaError_t e = cudaMemsetAsync(d_cnt.p, 0, (size_t)L * 6 * 4, st);
        const int64_t n = (int64_t)pi.pos.size();
        if (e == cudaSuccess && n > 0) {
            const int ctas = (int)std::min<int64_t>((n + WARPS_PER_CTA - 1) / WARPS_PER_CTA, (int64_t)ctx->sm_count * 8);
            pileup_kernel<<<ctas, WARPS_PER_CTA * 32, 0, st>>>(d_pos.as<int32_t>(), d_co.as<int64_t>(), d_c.as<uint32_t>(),
                                                                d_so.as<int64_t>(), d_s.as<char>(), n, L, d_cnt.as<uint32_t>());
            ctx->launches++;
        }
        ntset_kernel<<<(L + 255) / 256, 256, 0, st>>>(d_cnt.as<uint32_t>(), L, d_m.as<uint8_t>());
        ctx->launches++;
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(counts->data(), d_cnt.p, (size_t)L * 6 * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(mask->data(), d_m.p, L, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            hgt_set_error("pileup: %s", cudaGetErrorString(e));
            rc = HGT_ERR_CUDA;
        }
    } while (0);
    d_pos.release(); d_co.release(); d_c.release(); d_so.release(); d_s.release(); d_cnt.release(); d_m.release();
    return rc;
}

template <int WPL>
static void launch_compat(hgt_ctx *ctx, cudaStream_t st, const LocusDev &ld, int table, const int32_t *hl,
                          const int32_t *hr, const int64_t *ro, const int32_t *rows, int64_t n, uint64_t *out) {
    const size_t smem = (size_t)std::max(ld.V, 1) * 4;
    cudaFuncSetAttribute(compat_kernel<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ctas = (int)std::min<int64_t>((n + WARPS_PER_CTA - 1) / WARPS_PER_CTA, (int64_t)ctx->sm_count * 4);
    compat_kernel<WPL><<<ctas, WARPS_PER_CTA * 32, smem, st>>>(ld, table, hl, hr, ro, rows, n, out);
    ctx->launches++;
}

template <int WPL, int P>
static void launch_class(hgt_ctx *ctx, cudaStream_t st, int wp, const uint64_t *mask, int table, const int64_t *job_off,
                         const int32_t *job_list, int64_t n, const uint64_t *hapbits, const ClassPool &pool) {
    if (n <= 0) return;
    const int ctas = (int)std::min<int64_t>((n + WARPS_PER_CTA - 1) / WARPS_PER_CTA, (int64_t)ctx->sm_count * 8);
    class_kernel<WPL, P><<<ctas, WARPS_PER_CTA * 32, 0, st>>>(wp, mask, table, job_off, job_list, n, hapbits, pool);
    ctx->launches++;
}

static int wpl_of(int wp) {
    const int w = (wp + 31) / 32;
    return w <= 1 ? 1 : w <= 2 ? 2 : w <= 4 ? 4 : 8;
}

extern "C" int hgt_typing_run(hgt_ctx *ctx, hgt_locus *loc, const char *sam, size_t n_bytes, const hgt_params *params,
                              hgt_typing **out) {
    if (!ctx || !loc || !params || !out || (!sam && n_bytes)) {
        hgt_set_error("hgt_typing_run: null argument");
        return HGT_ERR_ARG;
    }
    *out = nullptr;
    if (!loc->ctx) {
        hgt_set_error("hgt_typing_run: locus was created without a context");
        return HGT_ERR_ARG;
    }
    if (loc->wp > 256) {
        hgt_set_error("typing kernels support at most 16384 alleles per locus (got %d)", loc->A);
        return HGT_ERR_UNSUPPORTED;
    }
    HGT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Intake in;
    HGT_CHECK(intake(sam, n_bytes, *params, &in));
    PileupIn pi;
    HGT_CHECK(build_pileup_input(in, *params, &pi));
    hgt_typing *t = new hgt_typing();
    t->ctx = ctx;
    t->locus = loc;
    int rc = run_pileup(ctx, loc, pi, &t->pile_counts, &t->pile_mask);
    HostOut ho;
    if (rc == HGT_OK) {
        PileupView pu;
        pu.counts = t->pile_counts.data(); pu.nt_mask = t->pile_mask.data(); pu.L = loc->L;
        rc = host_walk(loc, in, *params, pu, &ho);
    }
    if (rc != HGT_OK) {
        hgt_typing_free(t);
        return rc;
    }
    t->num_reads = ho.num_reads;
    t->num_pairs = ho.num_pairs;
    const int n_tables = loc->is_hla ? 3 : 1;
    const int wp = loc->wp;
    const int64_t n_pairs = ho.num_pairs;
    const int64_t max_classes = std::max<int64_t>(1, n_pairs * n_tables);
    uint32_t cap = 64;
    while ((int64_t)cap < 2 * max_classes) cap <<= 1;
    DevBuf d_keys, d_slot, d_ncls;
    DevBuf d_hl[3], d_hr[3], d_ro[3], d_rows[3], d_jo[3], d_hb[3], d_jl[3];
    std::vector<int32_t> cls_table;
    do {
        if (n_pairs == 0) break;
        if ((rc = d_keys.alloc((size_t)cap * 8)) != HGT_OK) break;
        if ((rc = d_slot.alloc((size_t)cap * 4)) != HGT_OK) break;
        if ((rc = d_ncls.alloc(4)) != HGT_OK) break;
        if ((rc = t->d_bits.alloc((size_t)max_classes * wp * 8)) != HGT_OK) break;
        if ((rc = t->d_count.alloc((size_t)max_classes * 8)) != HGT_OK) break;
        if ((rc = t->d_first.alloc((size_t)max_classes * 4)) != HGT_OK) break;
        if ((rc = t->d_table.alloc((size_t)max_classes * 4)) != HGT_OK) break;
        cudaError_t e = cudaMemsetAsync(d_keys.p, 0, (size_t)cap * 8, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_slot.p, 0xff, (size_t)cap * 4, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_ncls.p, 0, 4, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(t->d_count.p, 0, (size_t)max_classes * 8, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(t->d_first.p, 0x7f, (size_t)max_classes * 4, st);
        if (e != cudaSuccess) {
            hgt_set_error("typing: %s", cudaGetErrorString(e));
            rc = HGT_ERR_CUDA;
            break;
        }
        ClassPool pool;
        pool.keys = d_keys.as<unsigned long long>(); pool.slot_class = d_slot.as<int32_t>(); pool.cap_mask = cap - 1;
        pool.bits = t->d_bits.as<uint64_t>(); pool.count = t->d_count.as<unsigned long long>();
        pool.first = t->d_first.as<int32_t>(); pool.table = t->d_table.as<int32_t>();
        pool.n_classes = d_ncls.as<int32_t>(); pool.max_classes = (int32_t)max_classes;
        const LocusDev ld = locus_dev(loc);
        const int wpl = wpl_of(wp);
        for (int tb = 0; tb < n_tables && rc == HGT_OK; tb++) {
            const TableJobs &J = ho.tb[tb];
            const int64_t H = (int64_t)J.hap_left.size();
            // jobs with more than 7 haplotypes need the wide counter
            std::vector<int32_t> small, big;
            for (int64_t p = 0; p < n_pairs; p++) {
                const int64_t k = J.job_off[p + 1] - J.job_off[p];
                if (k > 255) {
                    hgt_set_error("a read pair expands to %lld haplotypes (limit 255)", (long long)k);
                    rc = HGT_ERR_UNSUPPORTED;
                    break;
                }
                (k <= 7 ? small : big).push_back((int32_t)p);
            }
            if (rc != HGT_OK) break;
            if ((rc = upload(&d_hl[tb], J.hap_left, st)) != HGT_OK) break;
            if ((rc = upload(&d_hr[tb], J.hap_right, st)) != HGT_OK) break;
            if ((rc = upload(&d_ro[tb], J.row_off, st)) != HGT_OK) break;
            if ((rc = upload(&d_rows[tb], J.rows, st)) != HGT_OK) break;
            if ((rc = upload(&d_jo[tb], J.job_off, st)) != HGT_OK) break;
            if ((rc = d_hb[tb].alloc((size_t)std::max<int64_t>(H, 1) * wp * 8)) != HGT_OK) break;
            std::vector<int32_t> jl(small);
            jl.insert(jl.end(), big.begin(), big.end());
            if ((rc = upload(&d_jl[tb], jl, st)) != HGT_OK) break;
            const uint64_t *mask = loc->d_mask + (size_t)tb * wp;
#define DISPATCH(W)                                                                                                  \
    {                                                                                                                \
        if (H > 0)                                                                                                   \
            launch_compat<W>(ctx, st, ld, tb, d_hl[tb].as<int32_t>(), d_hr[tb].as<int32_t>(), d_ro[tb].as<int64_t>(), \
                             d_rows[tb].as<int32_t>(), H, d_hb[tb].as<uint64_t>());                                  \
        launch_class<W, 3>(ctx, st, wp, mask, tb, d_jo[tb].as<int64_t>(), d_jl[tb].as<int32_t>(),                    \
                           (int64_t)small.size(), d_hb[tb].as<uint64_t>(), pool);                                    \
        launch_class<W, 8>(ctx, st, wp, mask, tb, d_jo[tb].as<int64_t>(), d_jl[tb].as<int32_t>() + small.size(),     \
                           (int64_t)big.size(), d_hb[tb].as<uint64_t>(), pool);                                      \
    }
            switch (wpl) {
                case 1: DISPATCH(1); break;
                case 2: DISPATCH(2); break;
                case 4: DISPATCH(4); break;
                default: DISPATCH(8); break;
            }
#undef DISPATCH
            e = cudaGetLastError();
            if (e != cudaSuccess) {
                hgt_set_error("typing kernels: %s", cudaGetErrorString(e));
                rc = HGT_ERR_CUDA;
            }
        }
        if (rc != HGT_OK) break;
        int32_t ncls = 0;
        e = cudaMemcpyAsync(&ncls, d_ncls.p, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            hgt_set_error("typing: %s", cudaGetErrorString(e));
            rc = HGT_ERR_CUDA;
            break;
        }
        t->total_classes = ncls;
        std::vector<unsigned long long> cnt(ncls);
        std::vector<int32_t> first(ncls);
        cls_table.resize(ncls);
        if (ncls > 0) {
            e = cudaMemcpyAsync(cnt.data(), t->d_count.p, (size_t)ncls * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(first.data(), t->d_first.p, (size_t)ncls * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(cls_table.data(), t->d_table.p, (size_t)ncls * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                hgt_set_error("typing: %s", cudaGetErrorString(e));
                rc = HGT_ERR_CUDA;
                break;
            }
        }
        for (int c = 0; c < ncls; c++) t->cls_idx[cls_table[c]].push_back(c);
        for (int tb = 0; tb < 3; tb++) {
            std::vector<int32_t> &ix = t->cls_idx[tb];
            std::sort(ix.begin(), ix.end(), [&](int a, int b) { return first[a] < first[b]; });
            t->n_classes[tb] = (int32_t)ix.size();
            for (int c : ix) {
                t->cls_count[tb].push_back((int64_t)cnt[c]);
                t->cls_first[tb].push_back(first[c]);
            }
        }
    } while (0);
    d_keys.release(); d_slot.release(); d_ncls.release();
    for (int tb = 0; tb < 3; tb++) {
        d_hl[tb].release(); d_hr[tb].release(); d_ro[tb].release(); d_rows[tb].release(); d_jo[tb].release();
        d_hb[tb].release(); d_jl[tb].release();
    }
    if (rc != HGT_OK) {
        hgt_typing_free(t);
        return rc;
    }
    *out = t;
    return HGT_OK;
}

extern "C" int hgt_typing_summary(const hgt_typing *t, int64_t *num_reads, int64_t *num_pairs, int32_t n_classes[3]) {
    if (!t) return HGT_ERR_ARG;
    if (num_reads) *num_reads = t->num_reads;
    if (num_pairs) *num_pairs = t->num_pairs;
    if (n_classes)
        for (int i = 0; i < 3; i++) n_classes[i] = t->n_classes[i];
    return HGT_OK;
}

extern "C" int hgt_typing_pileup(const hgt_typing *t, uint32_t *counts, uint8_t *nt_mask) {
    if (!t) return HGT_ERR_ARG;
    if (counts) memcpy(counts, t->pile_counts.data(), t->pile_counts.size() * 4);
    if (nt_mask) memcpy(nt_mask, t->pile_mask.data(), t->pile_mask.size());
    return HGT_OK;
}

// class rows of one table, first-seen order, gathered into a contiguous device buffer
static int gather_table(const hgt_typing *t, int table, DevBuf *d_idx, DevBuf *d_rows) {
    hgt_ctx *ctx = t->ctx;
    cudaStream_t st = ctx->stream;
    const int n = t->n_classes[table], wp = t->locus->wp;
    HGT_CHECK(upload(d_idx, t->cls_idx[table], st));
    HGT_CHECK(d_rows->alloc((size_t)std::max(n, 1) * wp * 8));
    if (n > 0) {
        gather_rows_kernel<<<std::min(ctx->sm_count * 4, (n * wp + 255) / 256), 256, 0, st>>>(
            wp, d_idx->as<int32_t>(), n, t->d_bits.as<uint64_t>(), d_rows->as<uint64_t>());
        ctx->launches++;
        HGT_CUDA(cudaGetLastError());
    }
    return HGT_OK;
}

extern "C" int hgt_typing_table(const hgt_typing *t, int32_t table, uint64_t *class_bits, int64_t *class_count,
                                int64_t *class_first, int64_t *allele_count, int64_t *allele_first) {
    if (!t || table < 0 || table > 2) {
        hgt_set_error("hgt_typing_table: bad argument");
        return HGT_ERR_ARG;
    }
    hgt_ctx *ctx = t->ctx;
    HGT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = t->n_classes[table], wp = t->locus->wp, A = t->locus->A;
    if (class_count) memcpy(class_count, t->cls_count[table].data(), (size_t)n * 8);
    if (class_first) memcpy(class_first, t->cls_first[table].data(), (size_t)n * 8);
    DevBuf d_idx, d_rows, d_ac, d_af;
    int rc = HGT_OK;
    do {
        if (class_bits && n > 0) {
            if ((rc = gather_table(t, table, &d_idx, &d_rows)) != HGT_OK) break;
            cudaError_t e = cudaMemcpyAsync(class_bits, d_rows.p, (size_t)n * wp * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { hgt_set_error("table readback: %s", cudaGetErrorString(e)); rc = HGT_ERR_CUDA; break; }
        }
        if (allele_count || allele_first) {
            if (!d_idx.p && (rc = upload(&d_idx, t->cls_idx[table], st)) != HGT_OK) break;
            if ((rc = d_ac.alloc((size_t)A * 8)) != HGT_OK) break;
            if ((rc = d_af.alloc((size_t)A * 8)) != HGT_OK) break;
            table_counts_kernel<<<(A + 127) / 128, 128, 0, st>>>(A, wp, d_idx.as<int32_t>(), n, t->d_bits.as<uint64_t>(),
                                                                 t->d_count.as<unsigned long long>(), t->d_first.as<int32_t>(),
                                                                 d_ac.as<long long>(), d_af.as<long long>());
            ctx->launches++;
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess && allele_count) e = cudaMemcpyAsync(allele_count, d_ac.p, (size_t)A * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && allele_first) e = cudaMemcpyAsync(allele_first, d_af.p, (size_t)A * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { hgt_set_error("table counts: %s", cudaGetErrorString(e)); rc = HGT_ERR_CUDA; break; }
        }
    } while (0);
    d_idx.release(); d_rows.release(); d_ac.release(); d_af.release();
    return rc;
}

extern "C" int hgt_typing_em(hgt_ctx *ctx, const hgt_typing *t, int32_t table, const uint64_t *keep_mask,
                             const double *allele_len, int32_t remove_low, double *prob, uint8_t *in_result,
                             int32_t *first_class, int32_t *iters) {
    if (!ctx || !t || table < 0 || table > 2) {
        hgt_set_error("hgt_typing_em: bad argument");
        return HGT_ERR_ARG;
    }
    const int n = t->n_classes[table], wp = t->locus->wp, A = t->locus->A;
    std::vector<uint64_t> bits((size_t)std::max(n, 1) * wp);
    std::vector<int64_t> cnt(std::max(n, 1));
    HGT_CHECK(hgt_typing_table(t, table, bits.data(), cnt.data(), nullptr, nullptr, nullptr));
    int m = n;
    if (keep_mask) {
        // project classes onto the kept alleles, merge equal keys in first-seen order (core:1753-1766)
        std::unordered_map<std::string, int> where;
        std::vector<uint64_t> b2;
        std::vector<int64_t> c2;
        for (int k = 0; k < n; k++) {
            bool any = false;
            for (int j = 0; j < wp; j++) {
                bits[(size_t)k * wp + j] &= keep_mask[j];
                any |= bits[(size_t)k * wp + j] != 0;
            }
            if (!any) continue;
            std::string key(reinterpret_cast<const char *>(&bits[(size_t)k * wp]), (size_t)wp * 8);
            auto f = where.find(key);
            if (f == where.end()) {
                where.emplace(std::move(key), (int)c2.size());
                b2.insert(b2.end(), bits.begin() + (size_t)k * wp, bits.begin() + (size_t)(k + 1) * wp);
                c2.push_back(cnt[k]);
            } else {
                c2[f->second] += cnt[k];
            }
        }
        bits.swap(b2);
        cnt.swap(c2);
        m = (int)cnt.size();
    }
    return hgt_em(ctx, bits.data(), cnt.data(), m, A, wp, allele_len, remove_low, prob, in_result, first_class, iters);
}

// ================================================================================================================
// Host-only walk (tests of the host logic without a GPU)
// ================================================================================================================
struct hgt_walk {
    HostOut ho;
};

extern "C" int hgt_host_walk(hgt_locus *loc, const char *sam, size_t n_bytes, const hgt_params *params,
                             const uint32_t *counts, const uint8_t *nt_mask, hgt_walk **out) {
    if (!loc || !params || !out || !counts || !nt_mask) {
        hgt_set_error("hgt_host_walk: null argument");
        return HGT_ERR_ARG;
    }
    *out = nullptr;
    Intake in;
    HGT_CHECK(intake(sam, n_bytes, *params, &in));
    hgt_walk *w = new hgt_walk();
    PileupView pu;
    pu.counts = counts; pu.nt_mask = nt_mask; pu.L = loc->L;
    int rc = host_walk(loc, in, *params, pu, &w->ho);
    if (rc != HGT_OK) {
        delete w;
        return rc;
    }
    *out = w;
    return HGT_OK;
}

extern "C" int hgt_walk_summary(const hgt_walk *w, int64_t *num_reads, int64_t *num_pairs, int64_t n_haps[3],
                                int64_t n_rows[3]) {
    if (!w) return HGT_ERR_ARG;
    if (num_reads) *num_reads = w->ho.num_reads;
    if (num_pairs) *num_pairs = w->ho.num_pairs;
    for (int t = 0; t < 3; t++) {
        if (n_haps) n_haps[t] = (int64_t)w->ho.tb[t].hap_left.size();
        if (n_rows) n_rows[t] = (int64_t)w->ho.tb[t].rows.size();
    }
    return HGT_OK;
}

extern "C" int hgt_walk_table(const hgt_walk *w, int32_t table, int64_t *job_off, int32_t *hap_left, int32_t *hap_right,
                              int64_t *row_off, int32_t *rows) {
    if (!w || table < 0 || table > 2) return HGT_ERR_ARG;
    const TableJobs &J = w->ho.tb[table];
    if (job_off) memcpy(job_off, J.job_off.data(), J.job_off.size() * 8);
    if (hap_left) memcpy(hap_left, J.hap_left.data(), J.hap_left.size() * 4);
    if (hap_right) memcpy(hap_right, J.hap_right.data(), J.hap_right.size() * 4);
    if (row_off) memcpy(row_off, J.row_off.data(), J.row_off.size() * 8);
    if (rows) memcpy(rows, J.rows.data(), J.rows.size() * 4);
    return HGT_OK;
}

extern "C" void hgt_walk_free(hgt_walk *w) { delete w; }
