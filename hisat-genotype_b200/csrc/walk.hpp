// Host-side description of one locus (what hgt_locus_create receives, before it is flattened for the device by
// build_walk_tables in typing.cu): variants in Var_list order, Alts_left / Alts_right entries, exon intervals.
// The record walk itself - round 1 ran it here on host threads - lives in walk_dev.cuh as __host__ __device__ code that
// the kernels of reads.cuh run on the GPU and hgt_host_walk runs in plain loops for the CPU-side tests.
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

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

}  // namespace hgt
