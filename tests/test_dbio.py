"""Database file readers / writer (hisat-genotype_b200/dbio.py): round trip on the databases of the goldens and the
extra index files the synthetic writer produces (.haplotype, .index.snp)."""
import os

import pytest

from conftest import GOLDEN_NAMES, load_golden
from helpers import golden_db


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_round_trip(name):
    from hisatgenotype_b200 import dbio
    d = golden_db(load_golden(name))
    again = dbio.load_database_text(dbio.write_database_text(d))
    for k in ("refGenes", "refGene_loci", "Vars", "Var_list", "Links", "Genes", "Gene_lengths", "partial_alleles"):
        assert again[k] == d[k], k
    assert {g: sorted(v) for g, v in again["Gene_names"].items()} == {g: sorted(v) for g, v in d["Gene_names"].items()}


def test_haplotype_and_index_snp_files(tmp_path):
    from hisatgenotype_b200 import dbio, synth
    loc = synth.make_locus("A", 5, L=900, n_alleles=20, n_groups=4, core_vars=12, pool_private=20, del_frac=0.2)
    synth.write_database([loc], "hla", str(tmp_path))
    prefix = os.path.join(str(tmp_path), "hla")
    d = dbio.load_database(prefix)
    hts = dbio.read_haplotypes(open(prefix + ".haplotype").read())
    idx = dbio.read_index_variants(open(prefix + ".index.snp").read())
    assert set(hts) == {"A"} and len(hts["A"]) > 0
    for ht_id, left, right, ids in hts["A"]:
        assert ht_id.startswith("ht") and 0 <= left <= right < len(loc.backbone)
        for v in ids:
            assert v in d["Vars"]["A"]
            assert left <= d["Vars"]["A"][v][1] <= right
    assert idx["A"] <= set(d["Vars"]["A"]) and len(idx["A"]) > 0


def _write_files(texts, prefix):
    for ext, text in texts.items():
        with open(prefix + ext, "w") as f:
            f.write(text)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_native_reader_matches_python_on_goldens(name, tmp_path):
    """hgt_db_* (csrc/dbio.cpp) against the Python readers on the databases of the reference-captured goldens."""
    from hisatgenotype_b200 import dbio
    d = golden_db(load_golden(name))
    prefix = os.path.join(str(tmp_path), "db")
    _write_files(dbio.write_database_text(d), prefix)
    py = dbio.load_database(prefix)
    nat = dbio.load_database_native(prefix)
    for k in ("refGenes", "refGene_loci", "Vars", "Var_list", "Links", "Genes", "Gene_names", "Gene_lengths", "partial_alleles"):
        assert nat[k] == py[k], k


def test_native_reader_index_files_and_errors(tmp_path):
    from hisatgenotype_b200 import _lib, dbio, synth
    loc = synth.make_locus("A", 5, L=900, n_alleles=20, n_groups=4, core_vars=12, pool_private=20, del_frac=0.2)
    loc2 = synth.make_locus("DQA1", 6, L=700, n_alleles=8, n_groups=2, core_vars=8, pool_private=10, del_frac=0.1)
    synth.write_database([loc, loc2], "hla", str(tmp_path))
    prefix = os.path.join(str(tmp_path), "hla")
    py = dbio.load_database(prefix)
    nat = dbio.load_database_native(prefix)
    for k in ("refGenes", "refGene_loci", "Vars", "Var_list", "Links", "Genes", "Gene_names", "Gene_lengths", "partial_alleles"):
        assert nat[k] == py[k], k
    assert nat["haplotypes"] == dbio.read_haplotypes(open(prefix + ".haplotype").read())
    assert nat["index_vars"] == dbio.read_index_variants(open(prefix + ".index.snp").read())
    # a missing mandatory file: the reference's message and exit (common:572-575)
    os.remove(prefix + ".link")
    with pytest.raises(SystemExit) as e:
        dbio.load_database_native(prefix)
    assert "index files missing" in str(e.value)
    # a malformed .snp line is refused
    open(prefix + ".link", "w").write("")
    with open(prefix + ".snp", "a") as f:
        f.write("hv999999\tsingle\tA*BACKBONE\tnot_a_number\tC\n")
    with pytest.raises(_lib.HgtError):
        dbio.load_database_native(prefix)


def test_native_writer_round_trip(tmp_path):
    """hgt_db_write: the files it writes give the same tables back, through the native and through the Python readers."""
    from hisatgenotype_b200 import dbio, synth
    loc = synth.make_locus("A", 9, L=1200, n_alleles=30, n_groups=5, core_vars=15, pool_private=30, del_frac=0.2)
    loc2 = synth.make_locus("B", 10, L=800, n_alleles=12, n_groups=3, core_vars=8, pool_private=12, del_frac=0.1)
    src = os.path.join(str(tmp_path), "src")
    os.makedirs(src)
    synth.write_database([loc, loc2], "hla", src)
    a, b = os.path.join(src, "hla"), os.path.join(str(tmp_path), "copy")
    dbio.copy_database_native(a, b)
    assert dbio.load_database_native(b) == dbio.load_database_native(a)
    pa, pb = dbio.load_database(a), dbio.load_database(b)
    for k in pa:
        assert pb[k] == pa[k], k
    assert dbio.read_haplotypes(open(b + ".haplotype").read()) == dbio.read_haplotypes(open(a + ".haplotype").read())
    assert dbio.read_index_variants(open(b + ".index.snp").read()) == dbio.read_index_variants(open(a + ".index.snp").read())
    # a second copy of the copy is byte-identical: the written form is canonical
    c = os.path.join(str(tmp_path), "copy2")
    dbio.copy_database_native(b, c)
    for ext in (".locus", ".snp", ".index.snp", ".link", "_backbone.fa", ".allele", ".partial", ".haplotype"):
        assert open(b + ext, "rb").read() == open(c + ext, "rb").read(), ext
