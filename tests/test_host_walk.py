"""CPU-side parity of the record stage (csrc/walk_dev.cuh, the __host__ __device__ functions the kernels call, run in plain
loops by hgt_host_walk): alignment text -> haplotype jobs, checked against the reference goldens by finishing the
allele-set algebra in Python."""
import numpy as np
import pytest

import hgt_oracle as O
from conftest import GOLDEN_NAMES, load_golden
from helpers import golden_db, oracle_locus, pileup_arrays, product_locus, tables_from_jobs


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_locus_tables_match_reference(name):
    g = load_golden(name)
    db = golden_db(g)
    for cap, alts in zip(g["loci"], g["alts"]):
        t = product_locus(g, db, cap["gene"], cap["Gene_names"], host_only=True)
        assert {k: sorted(v) for k, v in t.alts_left.items()} == alts["left"]
        assert {k: sorted(v) for k, v in t.alts_right.items()} == alts["right"]
        assert t.allele_rep_groups == cap["allele_rep_groups"]
        assert t.primary_rep_groups == cap["primary_exon_allele_rep_groups"]
        t.close()


@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_host_walk_tables(name, general, monkeypatch):
    """general = True takes the paths for databases whose variant ids are not "hv<N>" (id hash table, character form of
    the substring rule of identify_ambigious_diffs) and for reads longer than 128 bases (error correction base by base
    instead of through the per-read mismatch bits); the result must not change."""
    from hisatgenotype_b200.typing_core import HostWalk, make_params
    if general:
        monkeypatch.setenv("HGT_IRREGULAR_IDS", "1")
        monkeypatch.setenv("HGT_EMU_NO_ECMASK", "1")
    chunk_bytes = 0
    g = load_golden(name)
    p = g["params"]
    db = golden_db(g)
    for cap in g["loci"]:
        loc = oracle_locus(g, db, cap["gene"], cap["Gene_names"])
        counts, nt_sets = O.get_mpileup(cap["sam"], len(loc.ref_seq), 0, p["discordant"])
        c, m = pileup_arrays(counts, nt_sets)
        t = product_locus(g, db, cap["gene"], cap["Gene_names"], host_only=True)
        walk = HostWalk(t, cap["sam"], make_params(p["num_editdist"], p["error_correction"], p["discordant"],
                                                   p["simulation"], chunk_bytes=chunk_bytes), c, m)
        assert walk.num_reads == cap["num_reads"]
        assert walk.num_pairs == cap["num_pairs"]
        res = tables_from_jobs(loc, walk, p["base"] == "hla")
        assert res[0][0] == cap["Gene_cmpt"]
        assert res[0][1] == cap["Gene_counts"]
        assert res[1][0] == cap["Gene_exons_cmpt"]
        assert res[1][1] == cap["Gene_exons_counts"]
        assert res[2][0] == cap["Gene_primary_exons_cmpt"]
        assert res[2][1] == cap["Gene_primary_exons_counts"]
        t.close()


def test_host_walk_rejects_malformed():
    from hisatgenotype_b200 import _lib
    from hisatgenotype_b200.typing_core import HostWalk, make_params
    g = load_golden(GOLDEN_NAMES[0])
    db = golden_db(g)
    cap = g["loci"][0]
    t = product_locus(g, db, cap["gene"], cap["Gene_names"], host_only=True)
    L = len(t.ref_seq)
    c, m = np.zeros((L, 6), np.uint32), np.zeros(L, np.uint8)
    good = cap["sam"][0]
    bad = "\t".join(good.split("\t")[:8])
    with pytest.raises(_lib.HgtError):
        HostWalk(t, [bad], make_params(simulation=True), c, m)
    no_md = "\t".join(x for x in good.split("\t") if not x.startswith("MD:Z"))
    with pytest.raises(_lib.HgtError):
        HostWalk(t, [no_md], make_params(simulation=True), c, m)
    # empty input is fine: zero reads
    w = HostWalk(t, [], make_params(simulation=True), c, m)
    assert w.num_reads == 0 and w.num_pairs == 0
    t.close()
