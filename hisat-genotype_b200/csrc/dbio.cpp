// Native reader of the typing database text files (include/hgt.h "hgt_db_*").
//
// The files are the ones genotyping_locus loads (reference hisatgenotype_modules/hisatgenotype_typing_core.py:2417-2485
// through hisatgenotype_typing_common.py:279-309 read_locus, :313-334 read_backbone_alleles, :339-368 read_variants,
// :388-403 read_links) plus the two index-side files hisat2-build consumes; their formats are the ones
// hisatgenotype_typing_process.py writes (:1055-1063 .locus, :1088-1102 .snp / .index.snp, :1105-1106 .link, :1215-1220
// .haplotype, :1242-1244 .allele / .partial, 60-column FASTA).  Host code only; hisat-genotype_b200/dbio.py is the
// Python reader of the same files and tests/test_dbio.py holds the two against each other.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hgt.h"

void hgt_set_error(const char *fmt, ...);

namespace {

struct Variant {
    std::string id, data;
    int type;  // 0 single, 1 deletion, 2 insertion
    int64_t pos;
    bool in_index;
};
struct Haplotype {
    std::string id, ids;  // ids: comma separated, as in the file (without the trailing comma)
    int64_t left, right;
};
struct Gene {
    std::string gene, name, chrom, seq, strand;
    int64_t left = 0, right = 0, length = 0;
    std::vector<int64_t> exons;  // (left, right, primary) triples
    std::vector<Variant> vars;   // Var_list order: by position, file order inside one position (stable)
    std::vector<std::string> links;  // per variant: allele names, space separated ("" = no .link line)
    std::vector<std::string> alleles;  // .allele entries of this gene, file order
    std::vector<std::string> partial;  // .partial entries of this gene, file order
    std::vector<Haplotype> haps;
};

bool slurp(const std::string &path, std::string *out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}
std::vector<std::string> lines_of(const std::string &text) {
    std::vector<std::string> v;
    size_t a = 0;
    while (a <= text.size()) {
        size_t b = text.find('\n', a);
        if (b == std::string::npos) b = text.size();
        std::string l = text.substr(a, b - a);
        if (!l.empty() && l.back() == '\r') l.pop_back();
        if (!l.empty()) v.push_back(l);
        a = b + 1;
    }
    return v;
}
std::vector<std::string> split(const std::string &s, const char *seps) {
    std::vector<std::string> v;
    size_t a = 0;
    while (a <= s.size()) {
        size_t b = s.find_first_of(seps, a);
        if (b == std::string::npos) b = s.size();
        v.push_back(s.substr(a, b - a));
        a = b + 1;
    }
    return v;
}
std::vector<std::string> split_ws(const std::string &s) {  // str.split(): runs of blanks, no empty fields
    std::vector<std::string> v;
    std::istringstream ss(s);
    std::string w;
    while (ss >> w) v.push_back(w);
    return v;
}
std::string gene_of(const std::string &allele) { return allele.substr(0, allele.find('*')); }
bool to_int(const std::string &s, int64_t *out) {
    if (s.empty()) return false;
    char *end = nullptr;
    const long long v = strtoll(s.c_str(), &end, 10);
    if (*end != '\0') return false;
    *out = v;
    return true;
}

}  // namespace

struct hgt_db {
    std::vector<Gene> genes;  // .locus order
    std::map<std::string, int> index;
    std::string scratch;  // blob handed out by hgt_db_text (valid until the next call on this handle)
};

static int fail(hgt_db *db, const char *what, const std::string &detail) {
    hgt_set_error("hgt_db_open: %s (%s)", what, detail.c_str());
    delete db;
    return HGT_ERR_PARSE;
}

extern "C" int hgt_db_open(const char *prefix, hgt_db **out) {
    if (!prefix || !out) {
        hgt_set_error("hgt_db_open: null argument");
        return HGT_ERR_ARG;
    }
    *out = nullptr;
    const std::string p(prefix);
    std::string t_locus, t_snp, t_link, t_bb, t_allele, t_partial, t_isnp, t_hap;
    const struct { const char *ext; std::string *dst; bool required; } files[] = {
        {"_backbone.fa", &t_bb, true}, {".locus", &t_locus, true},   {".snp", &t_snp, true},       {".link", &t_link, true},
        {".allele", &t_allele, true},  {".partial", &t_partial, true}, {".index.snp", &t_isnp, false}, {".haplotype", &t_hap, false}};
    for (const auto &f : files) {
        if (!slurp(p + f.ext, f.dst) && f.required) {
            hgt_set_error("Error: index files missing (%s%s)", prefix, f.ext);  // common:572-575
            return HGT_ERR_ARG;
        }
    }
    hgt_db *db = new hgt_db;
    // .locus: name chr left right length exon_str strand (common:279-309)
    for (const std::string &l : lines_of(t_locus)) {
        const std::vector<std::string> c = split_ws(l);
        if (c.size() != 7) return fail(db, ".locus line needs 7 columns", l);
        Gene g;
        g.name = c[0];
        g.gene = gene_of(c[0]);
        g.chrom = c[1];
        g.strand = c[6];
        if (!to_int(c[2], &g.left) || !to_int(c[3], &g.right) || !to_int(c[4], &g.length)) return fail(db, "bad integer in .locus", l);
        for (std::string ex : split(c[5], ",")) {
            if (ex.empty()) continue;
            int64_t primary = 0, a = 0, b = 0;
            if (ex.back() == 'p') {
                primary = 1;
                ex.pop_back();
            }
            const size_t dash = ex.find('-');
            if (dash == std::string::npos || !to_int(ex.substr(0, dash), &a) || !to_int(ex.substr(dash + 1), &b))
                return fail(db, "bad exon in .locus", l);
            g.exons.push_back(a);
            g.exons.push_back(b);
            g.exons.push_back(primary);
        }
        if (db->index.count(g.gene)) return fail(db, "gene listed twice in .locus", g.gene);
        db->index[g.gene] = (int)db->genes.size();
        db->genes.push_back(g);
    }
    // _backbone.fa (common:313-334): one sequence per gene
    {
        std::string name;
        for (const std::string &l : lines_of(t_bb)) {
            if (l[0] == '>') {
                name = l.substr(1);
                const auto it = db->index.find(gene_of(name));
                if (it == db->index.end()) {
                    name.clear();  // a backbone without a .locus line: ignored, as the reference only looks names up
                    continue;
                }
                if (!db->genes[it->second].seq.empty()) return fail(db, "Nonunique sequence name", name);
            } else if (!name.empty()) {
                db->genes[db->index[gene_of(name)]].seq += l;
            }
        }
    }
    // .index.snp ids (the variants that went into the graph index; HISAT2 names them in the Zs tag)
    std::map<std::string, bool> in_index;
    for (const std::string &l : lines_of(t_isnp)) in_index[l.substr(0, l.find('\t'))] = true;
    // .snp: id type backbone pos data (common:339-368)
    for (const std::string &l : lines_of(t_snp)) {
        const std::vector<std::string> c = split(l, "\t");
        if (c.size() != 5) return fail(db, ".snp line needs 5 columns", l);
        const auto it = db->index.find(gene_of(c[2]));
        if (it == db->index.end()) continue;
        Variant v;
        v.id = c[0];
        v.data = c[4];
        v.type = c[1] == "single" ? 0 : c[1] == "deletion" ? 1 : c[1] == "insertion" ? 2 : -1;
        if (v.type < 0 || !to_int(c[3], &v.pos)) return fail(db, "bad .snp line", l);
        v.in_index = in_index.count(v.id) > 0;
        db->genes[it->second].vars.push_back(v);
    }
    // .link: id <tab or blank> alleles (common:388-403)
    std::map<std::string, std::string> links;
    for (const std::string &l : lines_of(t_link)) {
        const std::vector<std::string> c = split_ws(l);
        if (c.empty()) continue;
        if (links.count(c[0])) return fail(db, "variant listed twice in .link", c[0]);
        std::string a;
        for (size_t k = 1; k < c.size(); k++) a += (k > 1 ? " " : "") + c[k];
        links[c[0]] = a;
    }
    for (Gene &g : db->genes) {
        std::stable_sort(g.vars.begin(), g.vars.end(), [](const Variant &x, const Variant &y) { return x.pos < y.pos; });
        g.links.resize(g.vars.size());
        for (size_t k = 0; k < g.vars.size(); k++) {
            const auto it = links.find(g.vars[k].id);
            if (it != links.end()) g.links[k] = it->second;
        }
    }
    // .allele / .partial: one name per line (core:2442-2467)
    for (const std::string &l : lines_of(t_allele)) {
        const std::vector<std::string> c = split_ws(l);
        if (c.empty()) continue;
        const auto it = db->index.find(gene_of(c[0]));
        if (it != db->index.end()) db->genes[it->second].alleles.push_back(c[0]);
    }
    for (const std::string &l : lines_of(t_partial)) {
        const std::vector<std::string> c = split_ws(l);
        if (c.empty()) continue;
        const auto it = db->index.find(gene_of(c[0]));
        if (it != db->index.end()) db->genes[it->second].partial.push_back(c[0]);
    }
    // .haplotype: htN backbone left right ids, (process.py:1215-1220)
    for (const std::string &l : lines_of(t_hap)) {
        const std::vector<std::string> c = split(l, "\t");
        if (c.size() != 5) return fail(db, ".haplotype line needs 5 columns", l);
        const auto it = db->index.find(gene_of(c[1]));
        if (it == db->index.end()) continue;
        Haplotype h;
        h.id = c[0];
        h.ids = c[4];
        while (!h.ids.empty() && h.ids.back() == ',') h.ids.pop_back();
        if (!to_int(c[2], &h.left) || !to_int(c[3], &h.right)) return fail(db, "bad integer in .haplotype", l);
        db->genes[it->second].haps.push_back(h);
    }
    *out = db;
    return HGT_OK;
}

extern "C" void hgt_db_close(hgt_db *db) { delete db; }

extern "C" int32_t hgt_db_n_genes(const hgt_db *db) { return db ? (int32_t)db->genes.size() : 0; }

extern "C" int hgt_db_sizes(const hgt_db *db, int32_t g, int64_t out[8]) {
    if (!db || !out || g < 0 || g >= (int32_t)db->genes.size()) {
        hgt_set_error("hgt_db_sizes: bad argument");
        return HGT_ERR_ARG;
    }
    const Gene &G = db->genes[g];
    out[0] = (int64_t)G.seq.size();
    out[1] = (int64_t)G.exons.size() / 3;
    out[2] = (int64_t)G.vars.size();
    out[3] = (int64_t)G.alleles.size();
    out[4] = (int64_t)G.partial.size();
    out[5] = (int64_t)G.haps.size();
    out[6] = G.left;
    out[7] = G.right;
    return HGT_OK;
}

extern "C" int hgt_db_ints(const hgt_db *db, int32_t g, int32_t what, int64_t *out, int64_t cap) {
    if (!db || !out || g < 0 || g >= (int32_t)db->genes.size()) {
        hgt_set_error("hgt_db_ints: bad argument");
        return HGT_ERR_ARG;
    }
    const Gene &G = db->genes[g];
    std::vector<int64_t> v;
    switch (what) {
        case HGT_DB_EXONS: v = G.exons; break;
        case HGT_DB_VAR_TYPE: for (const Variant &x : G.vars) v.push_back(x.type); break;
        case HGT_DB_VAR_POS: for (const Variant &x : G.vars) v.push_back(x.pos); break;
        case HGT_DB_VAR_IN_INDEX: for (const Variant &x : G.vars) v.push_back(x.in_index ? 1 : 0); break;
        case HGT_DB_HAP_RANGE:
            for (const Haplotype &h : G.haps) {
                v.push_back(h.left);
                v.push_back(h.right);
            }
            break;
        default: hgt_set_error("hgt_db_ints: unknown table %d", what); return HGT_ERR_ARG;
    }
    if ((int64_t)v.size() > cap) {
        hgt_set_error("hgt_db_ints: %lld values, room for %lld", (long long)v.size(), (long long)cap);
        return HGT_ERR_ARG;
    }
    if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(int64_t));
    return HGT_OK;
}

extern "C" int hgt_db_text(hgt_db *db, int32_t g, int32_t what, const char **text, size_t *n_bytes) {
    if (!db || !text || !n_bytes || g < 0 || g >= (int32_t)db->genes.size()) {
        hgt_set_error("hgt_db_text: bad argument");
        return HGT_ERR_ARG;
    }
    const Gene &G = db->genes[g];
    std::string &s = db->scratch;
    s.clear();
    auto add = [&s](const std::string &x) {
        s += x;
        s += '\n';
    };
    switch (what) {
        case HGT_DB_GENE: add(G.gene); add(G.name); add(G.chrom); add(G.strand); break;
        case HGT_DB_BACKBONE: s = G.seq; break;
        case HGT_DB_VAR_ID: for (const Variant &x : G.vars) add(x.id); break;
        case HGT_DB_VAR_DATA: for (const Variant &x : G.vars) add(x.data); break;
        case HGT_DB_VAR_LINKS: for (const std::string &x : G.links) add(x); break;
        case HGT_DB_ALLELES: for (const std::string &x : G.alleles) add(x); break;
        case HGT_DB_PARTIAL: for (const std::string &x : G.partial) add(x); break;
        case HGT_DB_HAP_ID: for (const Haplotype &h : G.haps) add(h.id); break;
        case HGT_DB_HAP_VARS: for (const Haplotype &h : G.haps) add(h.ids); break;
        default: hgt_set_error("hgt_db_text: unknown table %d", what); return HGT_ERR_ARG;
    }
    *text = s.data();
    *n_bytes = s.size();
    return HGT_OK;
}

// Writes the handle's tables back as database files (the formats of process.py:1055-1106, 1215-1220, 1242-1244): variants
// in Var_list order, one .link line per variant that has alleles, 60-column FASTA, .index.snp / .haplotype only when the
// handle holds such entries.  hgt_db_open(prefix_out) then gives the same tables.
extern "C" int hgt_db_write(const hgt_db *db, const char *prefix) {
    if (!db || !prefix) {
        hgt_set_error("hgt_db_write: null argument");
        return HGT_ERR_ARG;
    }
    static const char *type_name[3] = {"single", "deletion", "insertion"};
    std::ostringstream locus, snp, isnp, link, bb, allele, partial, hap;
    bool any_index = false, any_hap = false;
    for (const Gene &g : db->genes) {
        locus << g.name << '\t' << g.chrom << '\t' << g.left << '\t' << g.right << '\t' << (g.length ? g.length : (int64_t)g.seq.size())
              << '\t';
        for (size_t k = 0; k + 2 < g.exons.size(); k += 3) {
            if (k) locus << ',';
            locus << g.exons[k] << '-' << g.exons[k + 1] << (g.exons[k + 2] ? "p" : "");
        }
        locus << '\t' << g.strand << '\n';
        bb << '>' << g.name << '\n';
        for (size_t k = 0; k < g.seq.size(); k += 60) bb << g.seq.substr(k, 60) << '\n';
        for (size_t k = 0; k < g.vars.size(); k++) {
            const Variant &v = g.vars[k];
            std::ostringstream line;
            line << v.id << '\t' << type_name[v.type] << '\t' << g.name << '\t' << v.pos << '\t' << v.data << '\n';
            snp << line.str();
            if (v.in_index) {
                isnp << line.str();
                any_index = true;
            }
            if (!g.links[k].empty()) link << v.id << '\t' << g.links[k] << '\n';
        }
        for (const std::string &a : g.alleles) allele << a << '\n';
        for (const std::string &a : g.partial) partial << a << '\n';
        for (const Haplotype &h : g.haps) {
            hap << h.id << '\t' << g.name << '\t' << h.left << '\t' << h.right << '\t' << h.ids << '\n';
            any_hap = true;
        }
    }
    const std::string p(prefix);
    const struct { const char *ext; std::string text; bool write; } files[] = {
        {".locus", locus.str(), true},   {".snp", snp.str(), true},         {".index.snp", isnp.str(), any_index}, {".link", link.str(), true},
        {"_backbone.fa", bb.str(), true}, {".allele", allele.str(), true}, {".partial", partial.str(), true},     {".haplotype", hap.str(), any_hap}};
    for (const auto &f : files) {
        if (!f.write) continue;
        std::ofstream out(p + f.ext, std::ios::binary);
        if (!out || !(out << f.text)) {
            hgt_set_error("hgt_db_write: cannot write %s%s", prefix, f.ext);
            return HGT_ERR_ARG;
        }
    }
    return HGT_OK;
}
