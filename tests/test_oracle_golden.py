"""Pin the oracle (oracle/hgt_oracle.py) to vectors captured from the unmodified reference."""
import pytest

import hgt_oracle as O
from conftest import GOLDEN_NAMES, load_golden
from hisatgenotype_b200 import dbio


def build_oracle_locus(g, gene):
    db = dbio.load_database_text(g["db"])
    cap = next(l for l in g["loci"] if l["gene"] == gene)
    names = cap["Gene_names"]
    assert sorted(names) == sorted(db["Gene_names"][gene])
    locus = db["refGene_loci"][gene]
    return O.OracleLocus(g["params"]["base"], gene, db["refGenes"][gene], db["Genes"][gene][db["refGenes"][gene]],
                         db["Vars"][gene], db["Var_list"][gene], db["Links"], names, db["Gene_lengths"][gene],
                         locus[4], locus[5]), db


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_gene_names_order_matches_reference(name):
    g = load_golden(name)
    db = dbio.load_database_text(g["db"])
    for cap in g["loci"]:
        assert db["Gene_names"][cap["gene"]] == cap["Gene_names"]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_alternatives_and_rep_groups(name):
    g = load_golden(name)
    genes = []
    for cap in g["loci"]:
        if cap["gene"] not in genes:
            genes.append(cap["gene"])
    # get_alternatives is called once per locus run, in run order
    for cap, alts in zip(g["loci"], g["alts"]):
        loc, _ = build_oracle_locus(g, cap["gene"])
        assert {k: sorted(v) for k, v in loc.alts_left.items()} == alts["left"]
        assert {k: sorted(v) for k, v in loc.alts_right.items()} == alts["right"]
        assert {k: v for k, v in loc.allele_rep_groups.items()} == cap["allele_rep_groups"]
        assert {k: v for k, v in loc.primary_rep_groups.items()} == cap["primary_exon_allele_rep_groups"]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_stage_a_tables(name):
    g = load_golden(name)
    p = g["params"]
    for cap, mp in zip(g["loci"], g["mpileup"]):
        loc, _ = build_oracle_locus(g, cap["gene"])
        res = O.type_locus(loc, cap["sam"], simulation=p["simulation"], num_editdist=p["num_editdist"],
                           error_correction=p["error_correction"], allow_discordant=p["discordant"])
        assert ["".join(sorted(s)) for s in res["nt_sets"]] == mp["nt_set"]
        assert [dict(c) for c in res["counts"]] == mp["counts"]
        assert res["num_reads"] == cap["num_reads"]
        assert res["num_pairs"] == cap["num_pairs"]
        t = res["tables"]
        assert t["gene"].cmpt_items(loc) == cap["Gene_cmpt"]
        assert t["gene"].count_items(loc) == cap["Gene_counts"]
        assert t["exon"].cmpt_items(loc) == cap["Gene_exons_cmpt"]
        assert t["exon"].count_items(loc) == cap["Gene_exons_counts"]
        assert t["primary"].cmpt_items(loc) == cap["Gene_primary_exons_cmpt"]
        assert t["primary"].count_items(loc) == cap["Gene_primary_exons_counts"]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_identify_ambiguous_diffs_cases(name):
    g = load_golden(name)
    # all captured calls of a scenario belong to loci in run order; try each locus until ids resolve
    locs = {}
    for cap in g["loci"]:
        if cap["gene"] not in locs:
            locs[cap["gene"]] = build_oracle_locus(g, cap["gene"])[0]
    n = 0
    for case in g["iad"]:
        ids = [e[3] for e in case["cmp_list"] if len(e) > 3 and e[3].startswith("hv")]
        cand = [l for l in locs.values() if all(i in l.gene_vars for i in ids)]
        lo = case["cmp_list"][0][1]
        ok = False
        for loc in cand:
            try:
                l, r, la, ra = O.identify_ambiguous_diffs(loc, [list(e) for e in case["cmp_list"]])
            except (AssertionError, KeyError):
                continue
            if [l, r, sorted(la), sorted(ra)] == case["result"]:
                ok = True
                break
        assert ok, (case, lo)
        n += 1
    assert n == len(g["iad"])


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_em_python_restatement(name):
    g = load_golden(name)
    for call in g["em_calls"]:
        res, _ = O.single_abundance(dict((k, c) for k, c in call["cmpt"]), call["remove_low"], call["lengths"])
        assert [a for a, _ in res] == [a for a, _ in call["result"]]
        for (a, p), (b, q) in zip(res, call["result"]):
            assert p == pytest.approx(q, rel=1e-12, abs=1e-300)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_em_c_restatement_bit_exact(name):
    import em_oracle
    g = load_golden(name)
    for call in g["em_calls"]:
        res, _ = em_oracle.single_abundance(call["cmpt"], call["remove_low"], call["lengths"])
        assert res == [[a, p] for a, p in call["result"]]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_em_driver_reproduces_reference_reports(name):
    """oracle/hgt_oracle.py::locus_abundance (the two-level EM driver, core:1679-1789) on the oracle's own tables must
    reproduce the abundance lines the unmodified reference wrote (the report body of every captured run)."""
    from hisatgenotype_b200 import report
    g = load_golden(name)
    p = g["params"]
    n_loci = len(p["loci"])
    reports = [g["reports"][k] for k in sorted(g["reports"], key=lambda s: int(s.split("test-")[1].split(".")[0]))]
    for t_i, ref_text in enumerate(reports):
        body = [report.aligner_line("hisat2", "graph")]
        for cap in g["loci"][t_i * n_loci:(t_i + 1) * n_loci]:
            loc, _ = build_oracle_locus(g, cap["gene"])
            res = O.type_locus(loc, cap["sam"], simulation=p["simulation"], num_editdist=p["num_editdist"],
                               error_correction=p["error_correction"], allow_discordant=p["discordant"])
            prob = O.locus_abundance(loc, res, p["remove_low"])
            block, _ = report.locus_block(res["num_reads"], res["num_pairs"], res["tables"]["gene"].count_items(loc), prob,
                                          p["simulation"], cap["test_Gene_names"], p["output_allele_counts"],
                                          p["best_alleles"])
            body.append(block)
        marker = "\n\t\thisat2 graph\n"
        assert "".join(body) == ref_text[ref_text.index(marker):]
