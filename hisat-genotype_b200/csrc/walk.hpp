// Host side of stage (a): alignment-record intake and the CIGAR x MD x Zs walk that turns one HISAT2 record
// into haplotypes (left, variant rows..., right).  Pure C++ (no CUDA); the allele-set algebra on those
// haplotypes runs on the GPU (typing.cu).
//
// Restates, record by record, reference hisatgenotype_modules/hisatgenotype_typing_core.py:800-1406
//   (filters :815-874, walk :876-1095, error_correct :119-243, post-filters :1117-1124, novel variants
//   :1126-1164, cmp_list2 :1351-1368, haplotype assembly :1386-1406, get_exon_haplotypes :718-792) and
// hisatgenotype_typing_common.py:1663-1955 (identify_ambigious_diffs) + validation_check.py:313-341.
// Semantics are pinned by tests/golden/*.json.gz through the oracle (oracle/hgt_oracle.py).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

namespace hgt {

enum : uint8_t { T_SINGLE = 0, T_DELETION = 1, T_INSERTION = 2 };
enum : uint8_t { C_MATCH = 0, C_MISMATCH = 1, C_INSERTION = 2, C_DELETION = 3 };
constexpr int32_t VAR_UNKNOWN = -1;
constexpr int32_t VAR_NOVEL_BASE = -10;  // novel indel k of the unit is encoded as VAR_NOVEL_BASE - k
inline bool var_is_novel(int32_t v) { return v <= VAR_NOVEL_BASE; }
inline int32_t novel_index(int32_t v) { return VAR_NOVEL_BASE - v; }

struct VarInfo {
    int32_t pos = 0, len = 1;  // len: deletion length / inserted length / 1
    uint8_t type = T_SINGLE;
    char base = 0;  // alt base of a `single`
    bool in_links = false, is_hv = false;
    std::string id;
    int32_t right() const { return type == T_DELETION ? pos + len - 1 : pos; }
};

struct AltHap {  // one alternative of an Alts_* entry: left-ids...-right
    int32_t left = 0, right = 0;
    std::vector<int32_t> rows;
};
struct AltEntry {
    int32_t anchor = 0;  // last token (left table) / first token (right table)
    std::string key;     // "529-hv8-hv22-606"
    std::vector<std::string> toks;
    std::vector<int32_t> tok_rows;  // row of every token that is a variant id of the locus, else -1
    std::vector<AltHap> alts;
};

struct LocusHost {
    std::string ref;
    std::vector<VarInfo> vars;  // Var_list order
    std::vector<int32_t> var_pos;
    std::unordered_map<std::string, int32_t> row_of;
    std::vector<std::pair<int32_t, int32_t>> exons, primary_exons;
    std::vector<AltEntry> alts_left, alts_right;  // sorted by anchor
    bool is_hla = false;
    // acceleration tables built by finalize(): anchors of Alts_left / Alts_right below each position, and whether
    // every variant id is "<letters><digits>" with one common letter prefix (then the reference's substring test on
    // '-'-joined ids, common:1734/1856, is a token comparison: all but the last id equal, the last one a prefix)
    std::vector<int32_t> altl_below, altr_below;  // [ref+2] number of anchors < p
    bool simple_ids = false;
    std::string id_prefix;
    std::vector<int32_t> row_by_num;  // simple ids: row of "<prefix><n>" or -1
    void finalize() {
        const int32_t n = (int32_t)ref.size();
        auto below = [&](const std::vector<AltEntry> &v, std::vector<int32_t> *out) {
            out->assign(n + 2, 0);
            std::vector<int32_t> cnt(n + 2, 0);
            for (const AltEntry &e : v) cnt[std::min(std::max(e.anchor, 0), n)]++;
            int32_t run = 0;
            for (int32_t p = 0; p <= n + 1; p++) {
                (*out)[p] = run;
                if (p <= n) run += cnt[p];
            }
        };
        below(alts_left, &altl_below);
        below(alts_right, &altr_below);
        simple_ids = !vars.empty();
        id_prefix.clear();
        int64_t max_num = -1;
        for (size_t i = 0; i < vars.size() && simple_ids; i++) {
            const std::string &id = vars[i].id;
            size_t k = 0;
            while (k < id.size() && ((id[k] >= 'a' && id[k] <= 'z') || (id[k] >= 'A' && id[k] <= 'Z'))) k++;
            if (k == 0 || k == id.size() || id.size() - k > 9) { simple_ids = false; break; }
            for (size_t j = k; j < id.size(); j++)
                if (id[j] < '0' || id[j] > '9') simple_ids = false;
            if (id.size() > k + 1 && id[k] == '0') simple_ids = false;  // leading zeros would alias numbers
            if (i == 0) id_prefix = id.substr(0, k);
            else if (id.compare(0, k, id_prefix) != 0 || k != id_prefix.size()) simple_ids = false;
            if (simple_ids) max_num = std::max<int64_t>(max_num, atoll(id.c_str() + k));
        }
        row_by_num.clear();
        if (simple_ids && max_num <= 50000000) {
            row_by_num.assign((size_t)max_num + 1, -1);
            for (size_t i = 0; i < vars.size(); i++) row_by_num[atoll(vars[i].id.c_str() + id_prefix.size())] = (int32_t)i;
        } else {
            simple_ids = false;
        }
        for (std::vector<AltEntry> *tab : {&alts_left, &alts_right})
            for (AltEntry &e : *tab) {
                e.tok_rows.assign(e.toks.size(), -1);
                for (size_t t = 0; t < e.toks.size(); t++) {
                    auto f = row_of.find(e.toks[t]);
                    if (f != row_of.end()) e.tok_rows[t] = f->second;
                }
            }
    }
    // row of a variant id given as characters (Zs tag), -3 when it is not a variant of this locus
    int32_t row_of_chars(const char *p, size_t n) const {
        if (simple_ids) {
            const size_t k = id_prefix.size();
            if (n <= k || n - k > 9 || memcmp(p, id_prefix.data(), k) != 0) return -3;
            if (n > k + 1 && p[k] == '0') return -3;
            int64_t v = 0;
            for (size_t j = k; j < n; j++) {
                if (p[j] < '0' || p[j] > '9') return -3;
                v = v * 10 + (p[j] - '0');
            }
            if (v >= (int64_t)row_by_num.size()) return -3;
            const int32_t r = row_by_num[(size_t)v];
            return r < 0 ? -3 : r;
        }
        auto f = row_of.find(std::string(p, n));
        return f == row_of.end() ? -3 : f->second;
    }
    int32_t lower_bound(int32_t pos) const {
        return (int32_t)(std::lower_bound(var_pos.begin(), var_pos.end(), pos) - var_pos.begin());
    }
};

struct Cmp {
    uint8_t type;
    int32_t pos, len;
    int32_t var;  // row >= 0, VAR_UNKNOWN, or a novel-indel code (var_is_novel); unused (-1) for matches
};

struct Record {
    const char *qname = nullptr;
    int32_t qname_len = 0;
    int32_t flag = 0;
    int32_t pos = 0;  // 0-based on the backbone, after subtracting base_locus+1
    const char *cigar = nullptr;
    int32_t cigar_len = 0;
    const char *seq = nullptr;
    int32_t seq_len = 0;
    int32_t NM = 0, NH = 0;
    bool has_NM = false, has_NH = false;
    const char *MD = nullptr;
    int32_t MD_len = 0;
    const char *Zs = nullptr;
    int32_t Zs_len = 0;
    bool ok = false;
};

struct Haplotype {
    int32_t left, right;
    std::vector<int32_t> ids;  // row >= 0 or a novel-indel code
    bool operator==(const Haplotype &o) const { return left == o.left && right == o.right && ids == o.ids; }
    bool operator<(const Haplotype &o) const {
        if (left != o.left) return left < o.left;
        if (right != o.right) return right < o.right;
        return ids < o.ids;
    }
};

struct NovelVar {
    uint8_t type;
    int32_t pos, len;
    bool operator==(const NovelVar &o) const { return type == o.type && pos == o.pos && len == o.len; }
};

struct WalkError {
    int code;
    std::string msg;
};

// ------------------------------------------------------------------------------------------------------------
// SAM line -> Record  (whitespace separated like Python's line.split(), core:804-841)
// ------------------------------------------------------------------------------------------------------------
inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

inline bool parse_int(const char *s, int n, int32_t *out) {
    if (n <= 0) return false;
    int i = 0;
    bool neg = false;
    if (s[0] == '-' || s[0] == '+') {
        neg = s[0] == '-';
        i = 1;
    }
    if (i >= n) return false;
    int64_t v = 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
        if (v > 2000000000LL) return false;
    }
    *out = (int32_t)(neg ? -v : v);
    return true;
}

inline void parse_record(const char *line, const char *end, bool simulation, int32_t base_locus, Record *r) {
    *r = Record();
    const char *p = line;
    int col = 0;
    while (p < end) {
        while (p < end && is_ws(*p)) p++;
        if (p >= end) break;
        const char *q = p;
        while (q < end && !is_ws(*q)) q++;
        const int n = (int)(q - p);
        switch (col) {
            case 0: {
                r->qname = p;
                r->qname_len = n;
                if (simulation) {  // read_id.split('|')[0]
                    const void *bar = memchr(p, '|', n);
                    if (bar) r->qname_len = (int)((const char *)bar - p);
                }
                break;
            }
            case 1: if (!parse_int(p, n, &r->flag)) return; break;
            case 3: if (!parse_int(p, n, &r->pos)) return; r->pos -= base_locus + 1; break;
            case 5: r->cigar = p; r->cigar_len = n; break;
            case 9: r->seq = p; r->seq_len = n; break;
            default:
                if (col >= 11 && n >= 5) {
                    if (p[0] == 'Z' && p[1] == 's') { r->Zs = p + 5; r->Zs_len = n - 5; }
                    else if (p[0] == 'M' && p[1] == 'D') { r->MD = p + 5; r->MD_len = n - 5; }
                    else if (p[0] == 'N' && p[1] == 'M') { r->has_NM = parse_int(p + 5, n - 5, &r->NM); }
                    else if (p[0] == 'N' && p[1] == 'H') { r->has_NH = parse_int(p + 5, n - 5, &r->NH); }
                }
                break;
        }
        col++;
        p = q;
    }
    r->ok = col >= 11;
}

struct CigarOp {
    char op;
    int32_t len;
};
inline bool parse_cigar(const char *s, int n, std::vector<CigarOp> *out) {
    out->clear();
    int32_t v = 0;
    bool have = false;
    for (int i = 0; i < n; i++) {
        const char c = s[i];
        if (c >= '0' && c <= '9') {
            v = v * 10 + (c - '0');
            have = true;
        } else {
            if (!have) return false;
            out->push_back({c, v});
            v = 0;
            have = false;
        }
    }
    return !have;
}

// ------------------------------------------------------------------------------------------------------------
// Pileup view handed to the walk: per position the nt_set bit mask (bit i = "ACGT"[i]) and the deletion-artefact
// flag; produced on the GPU (typing.cu pileup kernels), semantics of common:1059-1134.
// ------------------------------------------------------------------------------------------------------------
struct PileupView {
    const uint8_t *nt_mask = nullptr;       // [L] nt_set as a bit mask
    const uint8_t *del_artefact = nullptr;  // [L] 1 where del_count * 6 < nt_count (core:1064-1077)
    int32_t L = 0;
};
inline int nt_code(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 4;
    }
}

// Known `single` variant at pos with this base (core:159-169, 204-214, 949-961) else VAR_UNKNOWN.
inline int32_t known_single(const LocusHost &L, int32_t pos, char base) {
    for (int32_t j = L.lower_bound(pos); j < (int32_t)L.vars.size() && L.vars[j].pos == pos; j++)
        if (L.vars[j].type == T_SINGLE && L.vars[j].base == base) return j;
    return VAR_UNKNOWN;
}

// error_correct (core:119-243) on the cmp entries of one M segment; seq is the mutable read
inline int error_correct(const LocusHost &L, std::string &seq, int32_t read_pos, const PileupView &pu,
                         std::vector<Cmp> &cmp, size_t seg_start) {
    std::vector<Cmp> out;
    int ncorr = 0;
    const int32_t reflen = (int32_t)L.ref.size();
    for (size_t k = seg_start; k < cmp.size(); k++) {
        Cmp e = cmp[k];
        if (e.pos >= reflen) {
            out.insert(out.end(), cmp.begin() + k, cmp.end());
            break;
        }
        if (e.type == C_MATCH) {
            int32_t last = 0;
            for (int32_t j = 0; j < e.len; j++) {
                if (read_pos + j >= (int32_t)seq.size() || e.pos + j >= reflen) continue;
                char bp = seq[read_pos + j];
                const uint8_t m = pu.nt_mask[e.pos + j];
                const int c = nt_code(bp);
                if (m != 0 && !(c < 4 && ((m >> c) & 1))) {
                    bp = (m & (m - 1)) ? 'N' : "ACGT"[__builtin_ctz(m)];
                    seq[read_pos + j] = bp;
                    ncorr++;
                    const int32_t vid = bp != 'N' ? known_single(L, e.pos + j, bp) : VAR_UNKNOWN;
                    if (j > last) out.push_back({C_MATCH, e.pos + last, j - last, -1});
                    out.push_back({C_MISMATCH, e.pos + j, 1, vid});
                    last = j + 1;
                }
            }
            if (last < e.len) out.push_back({C_MATCH, e.pos + last, e.len - last, -1});
        } else {
            char bp = seq[read_pos];
            const char ref_bp = L.ref[e.pos];
            const uint8_t m = pu.nt_mask[e.pos];
            const int c = nt_code(bp);
            if (m != 0 && !(c < 4 && ((m >> c) & 1))) {
                bp = (m & (m - 1)) ? 'N' : "ACGT"[__builtin_ctz(m)];
                seq[read_pos] = bp;
                if (bp == 'N') e.var = VAR_UNKNOWN;
                else if (bp == ref_bp) {
                    e = {C_MATCH, e.pos, 1, -1};
                    ncorr++;
                } else e.var = known_single(L, e.pos, bp);
            }
            out.push_back(e);
        }
        read_pos += e.len;  // NB: uses the ORIGINAL entry length (core:222)
    }
    cmp.resize(seg_start);
    for (const Cmp &e : out) {
        if (e.type == C_MATCH && cmp.size() > seg_start && cmp.back().type == C_MATCH) cmp.back().len += e.len;
        else cmp.push_back(e);
    }
    return ncorr;
}

struct ZsItem {
    int32_t off;
    char kind;
    int32_t row;  // row of the id, or -3 if the id is not a variant of this locus
};

struct WalkResult {
    std::vector<Cmp> cmp;
    int32_t right_pos = 0;
    int ncorr = 0;
    bool misaligned = false;
};

// CIGAR x MD x Zs walk (core:876-1095).  Returns false + err on input the reference would assert on.
inline bool walk_record(const LocusHost &L, const Record &r, const PileupView &pu, bool error_correction,
                        std::string &seq, std::vector<CigarOp> &cig, std::vector<ZsItem> &zs, WalkResult *w,
                        WalkError *err) {
    auto fail = [&](const char *m) {
        err->code = -6;
        err->msg = std::string(m) + " (read " + std::string(r.qname, r.qname_len) + ")";
        return false;
    };
    if (!r.MD || r.MD_len == 0) return fail("MD tag missing");
    if (!parse_cigar(r.cigar, r.cigar_len, &cig)) return fail("malformed CIGAR");
    seq.assign(r.seq, r.seq_len);
    zs.clear();
    if (r.Zs_len > 0) {
        const char *p = r.Zs, *end = r.Zs + r.Zs_len;
        while (p < end) {
            const char *q = (const char *)memchr(p, ',', end - p);
            if (!q) q = end;
            const char *b1 = (const char *)memchr(p, '|', q - p);
            if (!b1) return fail("malformed Zs item");
            const char *b2 = (const char *)memchr(b1 + 1, '|', q - b1 - 1);
            if (!b2 || b2 != b1 + 2) return fail("malformed Zs item");
            ZsItem it;
            if (!parse_int(p, (int)(b1 - p), &it.off)) return fail("malformed Zs offset");
            it.kind = b1[1];
            it.row = L.row_of_chars(b2 + 1, (size_t)(q - b2 - 1));
            zs.push_back(it);
            p = q + 1;
        }
    }
    const char *MD = r.MD;
    const int32_t MDn = r.MD_len;
    int32_t md_i = 0, md_len = 0;
    size_t zs_i = 0;
    int32_t zs_pos = zs.empty() ? 0 : zs[0].off;
    int32_t read_pos = 0, right_pos = r.pos;
    w->cmp.clear();
    w->ncorr = 0;
    w->misaligned = false;
    auto isdig = [](char c) { return c >= '0' && c <= '9'; };
    auto isnt = [](char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; };
    for (size_t ci = 0; ci < cig.size(); ci++) {
        const char op = cig[ci].op;
        const int32_t length = cig[ci].len;
        if (op == 'M') {
            bool first = true;
            int32_t used = 0;
            const size_t seg_start = w->cmp.size();
            while (true) {
                if (!first || md_len == 0) {
                    if (md_i >= MDn) return fail("MD shorter than CIGAR");
                    if (isdig(MD[md_i])) {
                        int32_t num = 0;
                        while (md_i < MDn && isdig(MD[md_i])) num = num * 10 + (MD[md_i++] - '0');
                        md_len += num;
                    }
                }
                if (md_len >= length) {
                    md_len -= length;
                    if (length > used) w->cmp.push_back({C_MATCH, right_pos + used, length - used, -1});
                    break;
                }
                first = false;
                if (read_pos + md_len >= (int32_t)seq.size()) return fail("MD runs past the read");
                const char base = seq[read_pos + md_len];
                if (md_i >= MDn || !isnt(MD[md_i])) return fail("MD reference base is not ACGT");
                md_i++;
                if (md_len > used) w->cmp.push_back({C_MATCH, right_pos + used, md_len - used, -1});
                int32_t vid = VAR_UNKNOWN;
                if (read_pos + md_len == zs_pos && zs_i < zs.size()) {
                    if (zs[zs_i].kind != 'S') return fail("Zs item is not a substitution");
                    if (zs[zs_i].row < 0) return fail("Zs id is not a variant of this locus");
                    vid = zs[zs_i].row;
                    zs_i++;
                    zs_pos += 1;
                    if (zs_i < zs.size()) zs_pos += zs[zs_i].off;
                } else {
                    vid = known_single(L, right_pos + md_len, base);
                }
                w->cmp.push_back({C_MISMATCH, right_pos + md_len, 1, vid});
                used = md_len + 1;
                md_len += 1;
                if (md_len == length) {
                    md_len = 0;
                    break;
                }
            }
            if (error_correction) w->ncorr += error_correct(L, seq, read_pos, pu, w->cmp, seg_start);
        } else if (op == 'I') {
            int32_t vid = VAR_UNKNOWN;
            if (read_pos == zs_pos && zs_i < zs.size()) {
                if (zs[zs_i].kind != 'I') return fail("Zs item is not an insertion");
                if (zs[zs_i].row < 0) return fail("Zs id is not a variant of this locus");
                vid = zs[zs_i].row;
                zs_i++;
                if (zs_i < zs.size()) zs_pos += zs[zs_i].off;
            } else {
                for (int32_t j = L.lower_bound(right_pos); j < (int32_t)L.vars.size() && L.vars[j].pos == right_pos; j++)
                    if (L.vars[j].type == T_INSERTION && L.vars[j].len == length) {
                        vid = j;
                        break;
                    }
            }
            w->cmp.push_back({C_INSERTION, right_pos, length, vid});
            for (int32_t j = read_pos; j < read_pos + length && j < (int32_t)seq.size(); j++)
                if (seq[j] == 'N') w->misaligned = true;
        } else if (op == 'D') {
            if (md_i < MDn && MD[md_i] == '0') md_i++;
            if (md_i >= MDn || MD[md_i] != '^') return fail("MD lacks ^ for a deletion");
            md_i++;
            while (md_i < MDn && isnt(MD[md_i])) md_i++;
            int32_t vid = VAR_UNKNOWN;
            if (read_pos == zs_pos && zs_i < zs.size() && zs[zs_i].kind == 'D') {
                if (zs[zs_i].row < 0) return fail("Zs id is not a variant of this locus");
                vid = zs[zs_i].row;
                zs_i++;
                if (zs_i < zs.size()) zs_pos += zs[zs_i].off;
            } else {
                for (int32_t j = L.lower_bound(right_pos); j < (int32_t)L.vars.size() && L.vars[j].pos == right_pos; j++)
                    if (L.vars[j].type == T_DELETION && L.vars[j].len == length) {
                        vid = j;
                        break;
                    }
            }
            w->cmp.push_back({C_DELETION, right_pos, length, vid});
            // artificial-deletion rule, hla only (core:1064-1077)
            if (right_pos < pu.L && L.is_hla && pu.del_artefact[right_pos]) w->misaligned = true;
        } else if (op == 'S') {
            if (ci == 0) zs_pos += length;
            else if (ci + 1 != cig.size()) return fail("soft clip in the middle of a CIGAR");
        } else {
            return fail("unsupported CIGAR operation");
        }
        if (op == 'M' || op == 'N' || op == 'D') right_pos += length;
        if (op == 'M' || op == 'I' || op == 'S') read_pos += length;
    }
    w->right_pos = right_pos;
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// identify_ambigious_diffs (common:1663-1955) on cmp_list2.  Ids are compared as the reference compares them
// (substring search on '-'-joined id strings), hence the strings.
// ------------------------------------------------------------------------------------------------------------
struct AltSide {  // element of left_alt_set: pos-ids ; of right_alt_set: ids-pos
    int32_t pos;
    std::vector<int32_t> ids;
    bool operator==(const AltSide &o) const { return pos == o.pos && ids == o.ids; }
};

inline void add_unique(std::vector<AltSide> &set, AltSide &&x) {
    for (const AltSide &y : set)
        if (y == x) return;
    set.push_back(std::move(x));
}

struct Ambig {
    int32_t cmp_left = 0, cmp_right = 0;
    std::vector<AltSide> left, right;
};

inline int32_t alt_lower_bound(const std::vector<AltEntry> &v, int32_t pos) {
    int32_t lo = 0, hi = (int32_t)v.size();
    while (lo < hi) {
        const int32_t m = (lo + hi) / 2;
        if (v[m].anchor < pos) lo = m + 1;
        else hi = m;
    }
    return lo;
}

// Scratch that survives across reads (no per-read allocation in the common case).
struct AmbigScratch {
    std::vector<int32_t> all_ids, idn, nov, sl;
};

// Does the '-'-joined id string of `ids` occur inside the entry's key (common:1734, 1856)?
inline bool key_contains_ids(const LocusHost &L, const AltEntry &a, const int32_t *ids, size_t m) {
    if (L.simple_ids) {
        const size_t T = a.tok_rows.size();
        if (m > T) return false;
        const std::string &last = L.vars[ids[m - 1]].id;
        for (size_t t = 0; t + m <= T; t++) {
            bool ok = true;
            for (size_t k = 0; k + 1 < m && ok; k++) ok = a.tok_rows[t + k] == ids[k];
            if (!ok) continue;
            const std::string &kt = a.toks[t + m - 1];
            if (kt.size() >= last.size() && memcmp(kt.data(), last.data(), last.size()) == 0) return true;
        }
        return false;
    }
    std::string joined;
    for (size_t i = 0; i < m; i++) {
        if (i) joined += '-';
        joined += L.vars[ids[i]].id;
    }
    return a.key.find(joined) != std::string::npos;
}

inline bool identify_ambiguous_diffs(const LocusHost &L, const std::vector<Cmp> &c2, Ambig *out, WalkError *err,
                                     AmbigScratch *sc) {
    const int32_t n = (int32_t)c2.size();
    out->cmp_left = 0;
    out->cmp_right = n - 1;
    out->left.clear();
    out->right.clear();
    const int32_t left = c2[0].pos, right = c2[n - 1].pos + c2[n - 1].len - 1;
    const int32_t reflen = (int32_t)L.ref.size();
    auto id_is_hv = [&](const Cmp &e) { return e.var >= 0 && L.vars[e.var].is_hv; };
    // prefix sums over the entries: known ids, novel ids and sequence length of c2[0..i)
    // (get_haplotype_and_seq, common:1679-1700)
    bool prefix_ready = false;
    auto build_prefix = [&]() {
        if (prefix_ready) return;
        prefix_ready = true;
        sc->all_ids.clear();
        sc->idn.assign(n + 1, 0);
        sc->nov.assign(n + 1, 0);
        sc->sl.assign(n + 1, 0);
        for (int32_t i = 0; i < n; i++) {
            const Cmp &e = c2[i];
            int32_t len = 0;
            if (e.type == C_MATCH) {
                const int32_t a = std::min(std::max(e.pos, 0), reflen), b = std::min(std::max(e.pos + e.len, 0), reflen);
                len = std::max(0, b - a);
            } else if (e.type == C_MISMATCH) {
                len = 1;
            }
            int32_t novel = 0;
            if (e.type != C_MATCH && e.var != VAR_UNKNOWN) {
                if (e.var >= 0) sc->all_ids.push_back(e.var);
                else novel = 1;
            }
            sc->idn[i + 1] = (int32_t)sc->all_ids.size();
            sc->nov[i + 1] = sc->nov[i] + novel;
            sc->sl[i + 1] = sc->sl[i] + len;
        }
    };
    auto hv_between = [&](int32_t lo, int32_t hi, std::vector<int32_t> *ids) {
        for (int32_t j = lo; j < hi; j++)
            if (c2[j].type != C_MATCH && id_is_hv(c2[j])) ids->push_back(c2[j].var);
    };
    // any anchor of the table inside [lo, hi]?  (conservative outside the backbone)
    auto any_anchor = [&](const std::vector<int32_t> &below, int32_t lo, int32_t hi) {
        if (lo < 0 || hi >= reflen) return true;
        if (hi < lo) return false;
        return below[hi + 1] - below[lo] > 0;
    };
    // ---- left end ------------------------------------------------------------------------------------------
    bool found = false;
    if (!L.alts_left.empty()) {
        for (int32_t i = n - 1; i >= 0; i--) {
            const Cmp &e = c2[i];
            if (e.type != C_MATCH) {
                if (e.type == C_INSERTION || !id_is_hv(e)) continue;
            }
            const int32_t cur_left = e.pos;
            const int32_t cur_right = (e.type == C_MATCH || e.type == C_DELETION) ? e.pos + e.len - 1 : e.pos;
            if (!any_anchor(L.altl_below, cur_left, cur_right)) continue;
            const int32_t start = std::min(alt_lower_bound(L.alts_left, cur_right + 1) + 1, (int32_t)L.alts_left.size());
            bool candidates = false;
            for (int32_t j = start - 1; j >= 0; j--) {
                if (L.alts_left[j].anchor < cur_left) break;
                if (L.alts_left[j].anchor <= cur_right) {
                    candidates = true;
                    break;
                }
            }
            if (!candidates) continue;
            build_prefix();
            const bool has_novel = sc->nov[i + 1] > 0;
            const int32_t cur_len = sc->sl[i + 1];
            const int32_t *cur_ids = sc->all_ids.data();
            const size_t n_cur = (size_t)sc->idn[i + 1];
            const size_t n_ids = n_cur + (has_novel ? 1 : 0);  // novel ids count as ids that never match
            bool hit = false;
            for (int32_t j = start - 1; j >= 0; j--) {
                const AltEntry &a = L.alts_left[j];
                if (a.anchor < cur_left) break;
                if (a.anchor > cur_right) continue;
                if (n_ids > 0) {
                    if (has_novel || !key_contains_ids(L, a, cur_ids, n_cur)) continue;
                }
                const int32_t ntok = (int32_t)a.toks.size() - 1;  // key.split('-')[:-1]
                if ((int32_t)n_cur + 1 == ntok) {
                    if (left < atoi(a.toks[0].c_str())) continue;
                } else {
                    int32_t k = ntok - (int32_t)n_cur - 1;
                    if (k < 0) k += ntok;  // Python negative index
                    if (k < 0 || k >= ntok) { err->code = -6; err->msg = "alt haplotype index out of range"; return false; }
                    const int32_t row = a.tok_rows[k];
                    if (row < 0) { err->code = -6; err->msg = "alt haplotype token is not a variant"; return false; }
                    if (left <= L.vars[row].right()) continue;
                }
                hit = true;
                for (const AltHap &alt : a.alts) {
                    int32_t seq_pos = cur_right - alt.right, cur_pos = alt.right;
                    std::vector<int32_t> part;
                    for (int32_t t = (int32_t)alt.rows.size() - 1; t >= 0; t--) {
                        const VarInfo &v = L.vars[alt.rows[t]];
                        const int32_t vp = v.right();
                        int32_t nxt = seq_pos + (cur_pos - vp);
                        if (nxt >= cur_len) break;
                        int32_t npos;
                        if (v.type == T_SINGLE) { nxt += 1; npos = vp - 1; }
                        else npos = vp - v.len;
                        part.insert(part.begin(), alt.rows[t]);
                        if (nxt >= cur_len) break;
                        seq_pos = nxt;
                        cur_pos = npos;
                    }
                    if (!part.empty()) {
                        const int32_t seq_left = cur_len - seq_pos - 1;
                        AltSide s;
                        s.pos = cur_pos - seq_left;
                        s.ids = part;
                        if (found) hv_between(i + 1, out->cmp_left, &s.ids);
                        add_unique(out->left, std::move(s));
                    }
                }
            }
            if (hit) {
                if (!found) {
                    out->cmp_left = i + 1;
                    // cur_ht_str; a hit implies the slice holds no novel id (the substring test would fail)
                    AltSide s;
                    s.pos = left;
                    s.ids.assign(cur_ids, cur_ids + n_cur);
                    add_unique(out->left, std::move(s));
                }
                found = true;
            }
        }
    }
    if (!found) add_unique(out->left, AltSide{left, {}});
    // ---- right end -----------------------------------------------------------------------------------------
    found = false;
    if (!L.alts_right.empty()) {
        for (int32_t i = 0; i < n; i++) {
            const Cmp &e = c2[i];
            if (e.type != C_MATCH) {
                if (e.type == C_INSERTION || !id_is_hv(e)) continue;
            }
            const int32_t cur_left = e.pos;
            const int32_t cur_right = (e.type == C_MATCH || e.type == C_DELETION) ? e.pos + e.len - 1 : e.pos;
            if (!any_anchor(L.altr_below, cur_left, cur_right)) continue;
            const int32_t start = alt_lower_bound(L.alts_right, cur_left);
            if (start >= (int32_t)L.alts_right.size() || L.alts_right[start].anchor > cur_right) continue;
            build_prefix();
            const bool has_novel = sc->nov[n] - sc->nov[i] > 0;
            const int32_t cur_len = sc->sl[n] - sc->sl[i];
            const int32_t *cur_ids = sc->all_ids.data() + sc->idn[i];
            const size_t n_cur = (size_t)(sc->idn[n] - sc->idn[i]);
            const size_t n_ids = n_cur + (has_novel ? 1 : 0);
            bool hit = false;
            for (int32_t j = start; j < (int32_t)L.alts_right.size(); j++) {
                const AltEntry &a = L.alts_right[j];
                if (a.anchor > cur_right) break;
                if (a.anchor < cur_left) continue;
                if (n_ids > 0) {
                    if (has_novel || !key_contains_ids(L, a, cur_ids, n_cur)) continue;
                }
                const int32_t ntok = (int32_t)a.toks.size() - 1;  // key.split('-')[1:]
                if ((int32_t)n_cur + 1 == ntok) {
                    if (right > atoi(a.toks[ntok].c_str())) continue;
                } else {
                    const int32_t k = (int32_t)n_cur;
                    if (k >= ntok) { err->code = -6; err->msg = "alt haplotype index out of range"; return false; }
                    const int32_t row = a.tok_rows[1 + k];
                    if (row < 0) { err->code = -6; err->msg = "alt haplotype token is not a variant"; return false; }
                    if (right >= L.vars[row].pos) continue;
                }
                hit = true;
                for (const AltHap &alt : a.alts) {
                    int32_t seq_pos = alt.left - cur_left, cur_pos = alt.left;
                    std::vector<int32_t> part;
                    for (size_t t = 0; t < alt.rows.size(); t++) {
                        const VarInfo &v = L.vars[alt.rows[t]];
                        int32_t nxt = seq_pos + (v.pos - cur_pos);
                        if (nxt >= cur_len) break;
                        int32_t npos;
                        if (v.type == T_SINGLE) { nxt += 1; npos = v.pos + 1; }
                        else npos = v.pos + v.len;
                        part.push_back(alt.rows[t]);
                        if (nxt >= cur_len) break;
                        seq_pos = nxt;
                        cur_pos = npos;
                    }
                    if (!part.empty()) {
                        const int32_t seq_left = cur_len - seq_pos - 1;
                        AltSide s;
                        s.pos = cur_pos + seq_left;
                        if (found) hv_between(out->cmp_right + 1, i, &s.ids);
                        s.ids.insert(s.ids.end(), part.begin(), part.end());
                        add_unique(out->right, std::move(s));
                    }
                }
            }
            if (hit) {
                if (!found) {
                    out->cmp_right = i - 1;
                    AltSide s;
                    s.pos = right;
                    s.ids.assign(cur_ids, cur_ids + n_cur);
                    add_unique(out->right, std::move(s));
                }
                found = true;
            }
        }
    }
    if (!found) add_unique(out->right, AltSide{right, {}});
    if (out->cmp_right < out->cmp_left) {
        out->cmp_left = 0;
        out->left.clear();
        out->left.push_back(AltSide{left, {}});
    }
    // check_amb_uniqueness (validation_check.py:313-341)
    std::vector<const std::vector<int32_t> *> seen;
    auto dup = [&](const std::vector<int32_t> &ids) {
        if (ids.empty()) return false;
        for (auto *p : seen)
            if (*p == ids) return true;
        seen.push_back(&ids);
        return false;
    };
    for (const AltSide &s : out->left)
        if (dup(s.ids)) { err->code = -7; err->msg = "ambiguous alternative haplotype sets (check_amb_uniqueness)"; return false; }
    for (const AltSide &s : out->right)
        if (dup(s.ids)) { err->code = -7; err->msg = "ambiguous alternative haplotype sets (check_amb_uniqueness)"; return false; }
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// get_exon_haplotypes (core:718-792)
// ------------------------------------------------------------------------------------------------------------
struct VarLite {
    uint8_t type;
    int32_t pos, len;
};
template <class GetVar>
inline void exon_haplotypes(const Haplotype &ht, const std::vector<std::pair<int32_t, int32_t>> &exons, GetVar get,
                            std::vector<Haplotype> *out) {
    for (const auto &ex : exons) {
        const int32_t e_left = ex.first, e_right = ex.second;
        if (e_left > ht.right || e_right < ht.left) continue;
        int32_t left = ht.left, right = ht.right;
        size_t lo = 0, hi = ht.ids.size();  // kept id range [lo, hi)
        if (left < e_left) {
            bool split = false;
            for (size_t i = 0; i < ht.ids.size(); i++) {
                const VarLite v = get(ht.ids[i]);
                if ((v.type != T_DELETION && v.pos >= e_left) || (v.type == T_DELETION && v.pos - 1 >= e_left)) {
                    left = e_left; lo = i; split = true;
                    break;
                }
                if (v.type == T_DELETION && v.pos + v.len >= e_left) {
                    left = v.pos + v.len; lo = i + 1; split = true;
                    break;
                }
            }
            if (!split) { left = e_left; lo = hi = 0; }
        }
        if (right > e_right) {
            bool split = false;
            for (size_t i = hi; i-- > lo;) {
                const VarLite v = get(ht.ids[i]);
                const int32_t r = v.type == T_DELETION ? v.pos + v.len - 1 : v.pos;
                if ((v.type != T_DELETION && r <= e_right) || (v.type == T_DELETION && r + 1 <= e_right)) {
                    right = e_right; hi = i + 1; split = true;
                    break;
                }
                if (v.type == T_DELETION && r - v.len <= e_right) {
                    right = r - v.len; hi = i; split = true;
                    break;
                }
            }
            if (!split) { right = e_right; lo = hi = 0; }
        }
        Haplotype h;
        h.left = left;
        h.right = right;
        h.ids.assign(ht.ids.begin() + lo, ht.ids.begin() + hi);
        out->push_back(std::move(h));
    }
}

}  // namespace hgt
