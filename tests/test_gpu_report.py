"""`.report` body from the CUDA path == the report the unmodified reference wrote for the same alignments."""
import pytest

from conftest import GOLDEN_NAMES, load_golden
from helpers import golden_db, product_locus

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_report_body_matches_reference(name):
    from hisatgenotype_b200 import typing_core as TC
    g = load_golden(name)
    p = g["params"]
    db = golden_db(g)
    n_loci = len(p["loci"])
    reports = [g["reports"][k] for k in sorted(g["reports"], key=lambda s: int(s.split("test-")[1].split(".")[0]))]
    assert len(reports) * n_loci == len(g["loci"])
    tables = {}
    for t_i, ref_text in enumerate(reports):
        caps = g["loci"][t_i * n_loci:(t_i + 1) * n_loci]
        for cap in caps:
            if cap["gene"] not in tables:
                tables[cap["gene"]] = product_locus(g, db, cap["gene"], cap["Gene_names"])
        body, passed = TC.typing_from_alignments(
            p["base"], tables, [cap["test_Gene_names"] for cap in caps], {cap["gene"]: cap["sam"] for cap in caps},
            p["simulation"], p["num_editdist"], p["error_correction"], p["discordant"], p["remove_low"],
            p["best_alleles"], p["output_allele_counts"])
        marker = "\n\t\thisat2 graph\n"
        assert marker in ref_text
        assert body == ref_text[ref_text.index(marker):]
    for t in tables.values():
        t.close()
