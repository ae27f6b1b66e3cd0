"""Assembly hand-off: the per-read haplotype strings read back from the GPU record stage
(Batch.unit_read_haplotypes -> hgt_batch_unit_reads) equal the ones the oracle's per-read loop builds (core:1386-1406),
record by record, including the reads with several alternative haplotypes."""
import pytest

import hgt_oracle as O
from conftest import GOLDEN_NAMES, load_golden
from helpers import golden_db, oracle_locus, product_locus

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_read_haplotypes_match_oracle(name):
    from hisatgenotype_b200 import typing_core as TC
    g = load_golden(name)
    p = g["params"]
    db = golden_db(g)
    multi = 0
    for cap in g["loci"]:
        ol = oracle_locus(g, db, cap["gene"], cap["Gene_names"])
        want = []
        O.type_locus(ol, cap["sam"], p["simulation"], p["num_editdist"], p["error_correction"], p["discordant"], collect_hts=want)
        t = product_locus(g, db, cap["gene"], cap["Gene_names"])
        batch = TC.Batch([t], TC.make_params(p["num_editdist"], p["error_correction"], p["discordant"], p["simulation"]), p["remove_low"])
        batch.add_unit(0, cap["sam"])
        batch.run()
        got = batch.unit_read_haplotypes(0, t.var_ids)
        assert len(got) == len(want) == cap["num_reads"]
        for (gl, gf, gh), (wl, wf, wh) in zip(got, want):
            assert (gl, gf) == (wl, wf)
            assert sorted(gh) == wh
            multi += len(wh) > 1
        raw = batch.unit_read_haplotypes(0)
        assert [len(h) for _, _, h in raw] == [len(h) for _, _, h in got]
        batch.close()
        t.close()
    if name == "hla_indel":
        assert multi > 0  # the scenario exists to exercise identify_ambigious_diffs
