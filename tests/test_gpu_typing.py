"""GPU parity of stage (a) (csrc/typing.cu + walk.hpp) through the C ABI: alignment lines in, Gene_cmpt /
Gene_counts / pileup out, bit-exact against the vectors captured from the unmodified reference; plus the EM
driver (core:1679-1789) on the device-resident tables."""
import numpy as np
import pytest

import hgt_oracle as O
from conftest import GOLDEN_NAMES, load_golden
from helpers import golden_db, oracle_locus, pileup_arrays, product_locus

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_tables_bit_exact_vs_reference(name):
    from hisatgenotype_b200 import typing_core as TC
    g = load_golden(name)
    p = g["params"]
    db = golden_db(g)
    em_i = 0
    for cap, mp in zip(g["loci"], g["mpileup"]):
        t = product_locus(g, db, cap["gene"], cap["Gene_names"])
        run = TC.type_locus(t, cap["sam"], p["num_editdist"], p["error_correction"], p["discordant"], p["simulation"])
        assert run.num_reads == cap["num_reads"]
        assert run.num_pairs == cap["num_pairs"]
        # pileup (common:1059-1134)
        counts, mask = run.pileup()
        code = {"A": 0, "C": 1, "G": 2, "T": 3, "D": 5}
        ref_counts = np.zeros_like(counts)
        for i, d in enumerate(mp["counts"]):
            for nt, k in d.items():
                ref_counts[i, code.get(nt, 4)] += k
        np.testing.assert_array_equal(counts, ref_counts)
        ref_mask = [sum(1 << "ACGT".index(c) for c in s) for s in mp["nt_set"]]
        np.testing.assert_array_equal(mask, np.asarray(ref_mask, np.uint8))
        # the three tables, dict order included
        assert list(map(list, run.gene_cmpt(TC.TABLE_GENE).items())) == cap["Gene_cmpt"]
        assert run.gene_counts(TC.TABLE_GENE) == cap["Gene_counts"]
        if p["base"] == "hla":
            assert list(map(list, run.gene_cmpt(TC.TABLE_EXON).items())) == cap["Gene_exons_cmpt"]
            assert run.gene_counts(TC.TABLE_EXON) == cap["Gene_exons_counts"]
            assert list(map(list, run.gene_cmpt(TC.TABLE_PRIMARY).items())) == cap["Gene_primary_exons_cmpt"]
            assert run.gene_counts(TC.TABLE_PRIMARY) == cap["Gene_primary_exons_counts"]
        # EM driver: same sequence of single_abundance calls with the same results
        calls = []
        orig = run.abundance

        def spy(table, keep=None, lengths=None, remove_low=False):
            res = orig(table, keep, lengths, remove_low)
            calls.append(res)
            return res

        run.abundance = spy
        TC.locus_abundance(run, p["remove_low"])
        for res in calls:
            ref = g["em_calls"][em_i]["result"]
            em_i += 1
            assert [a for a, _ in res] == [a for a, _ in ref]
            for (_, x), (_, y) in zip(res, ref):
                assert x == pytest.approx(y, rel=1e-6, abs=1e-12)
        run.close()
        t.close()
    assert em_i == len(g["em_calls"])


def _synthetic_case(seed, A=300, L=2500, n_reads=4000):
    """Bigger-than-golden case: database + reads from hisatgenotype_b200.synth, SAM synthesised from the true
    alignment; oracle on the same lines."""
    from hisatgenotype_b200 import synth
    loc = synth.make_locus("A", seed, L=L, n_alleles=A, n_groups=12, core_vars=60, pool_private=300, del_frac=0.1)
    rng = np.random.default_rng(seed)
    names = sorted(n for n in loc.alleles if loc.alleles[n])
    truth = [names[i] for i in rng.choice(len(names), 2, replace=False)]
    sam = synth.simulate_sam(loc, truth, n_pairs=n_reads // 2, rng=rng, err_rate=0.004)
    return loc, truth, sam


@pytest.mark.parametrize("seed", [3, 4])
def test_tables_bit_exact_vs_oracle_synthetic(seed):
    from hisatgenotype_b200 import synth
    from hisatgenotype_b200 import typing_core as TC
    if not hasattr(synth, "simulate_sam"):
        pytest.skip("synthetic SAM generator not built yet")
    loc, truth, sam = _synthetic_case(seed)
    cont = synth.reference_containers([loc], "hla")
    gene = "A"
    args = ("hla", gene, cont["refGenes"][gene], cont["Genes"][gene][cont["refGenes"][gene]], cont["Vars"][gene],
            cont["Var_list"][gene], cont["Links"], cont["Gene_names"][gene], cont["Gene_lengths"][gene],
            cont["refGene_loci"][gene][4], cont["refGene_loci"][gene][5])
    ol = O.OracleLocus(*args)
    from hisatgenotype_b200.locus import LocusTables
    t = LocusTables(*args)
    ref = O.type_locus(ol, sam, simulation=False)
    run = TC.type_locus(t, sam)
    assert run.num_reads == ref["num_reads"] and run.num_pairs == ref["num_pairs"]
    for tb, key in ((TC.TABLE_GENE, "gene"), (TC.TABLE_EXON, "exon"), (TC.TABLE_PRIMARY, "primary")):
        assert list(map(list, run.gene_cmpt(tb).items())) == ref["tables"][key].cmpt_items(ol)
        assert run.gene_counts(tb) == ref["tables"][key].count_items(ol)
    run.close()
    t.close()


@pytest.mark.parametrize("chunk_bytes", [0, 4000])
@pytest.mark.parametrize("name", ["hla_pair_err", "cyp_pair", "hla_indel"])
def test_batch_matches_reference(name, chunk_bytes):
    """All locus runs of a scenario as ONE batch (several loci, several units per locus): tables, counts and the
    two-level EM result must equal the per-run goldens."""
    from hisatgenotype_b200 import typing_core as TC
    g = load_golden(name)
    p = g["params"]
    db = golden_db(g)
    genes = []
    for cap in g["loci"]:
        if cap["gene"] not in genes:
            genes.append(cap["gene"])
    names = {cap["gene"]: cap["Gene_names"] for cap in g["loci"]}
    loci = [product_locus(g, db, gene, names[gene]) for gene in genes]
    batch = TC.Batch(loci, TC.make_params(p["num_editdist"], p["error_correction"], p["discordant"], p["simulation"],
                                          chunk_bytes=chunk_bytes), p["remove_low"])
    for cap in g["loci"]:
        batch.add_unit(genes.index(cap["gene"]), cap["sam"])
    batch.run()
    em_i = 0
    for u, cap in enumerate(g["loci"]):
        s = batch.unit_summary(u)
        assert s["num_reads"] == cap["num_reads"] and s["num_pairs"] == cap["num_pairs"]
        assert list(map(list, batch.unit_gene_cmpt(u, TC.TABLE_GENE).items())) == cap["Gene_cmpt"]
        assert batch.unit_gene_counts(u, TC.TABLE_GENE) == cap["Gene_counts"]
        if p["base"] == "hla":
            assert list(map(list, batch.unit_gene_cmpt(u, TC.TABLE_EXON).items())) == cap["Gene_exons_cmpt"]
            assert batch.unit_gene_counts(u, TC.TABLE_EXON) == cap["Gene_exons_counts"]
        for level in (0, 1):
            res = batch.unit_em(u, level)
            if res is None:
                continue
            ref = g["em_calls"][em_i]["result"]
            em_i += 1
            assert [a for a, _ in res] == [a for a, _ in ref]
            for (_, x), (_, y) in zip(res, ref):
                assert x == pytest.approx(y, rel=1e-6, abs=1e-12)
        ab = batch.unit_abundance(u)
        assert len(ab) > 0
        assert batch.unit_calls(u) == ab  # native ranking == Python combination rule (core:1771-1782)
        assert batch.unit_calls(u, 2) == ab[:2]
    assert em_i == len(g["em_calls"])
    # the tables an EM runs on come back in the reference's dict order (first pair of each class): class_sort_kernel
    for u in range(len(g["loci"])):
        for tb in ((TC.TABLE_EXON, 3) if p["base"] == "hla" else (TC.TABLE_GENE,)):
            first = batch.unit_table(u, tb)[2]
            assert (np.diff(first) > 0).all(), (u, tb)
    # every unit's calls through one library call (hgt_batch_abundances)
    assert batch.top_calls(3) == [batch.unit_calls(u, 3) for u in range(len(g["loci"]))]
    # repeat execute+finish on the prepared batch: identical tables (pools are reset)
    batch.execute()
    batch.finish()
    for u, cap in enumerate(g["loci"]):
        assert list(map(list, batch.unit_gene_cmpt(u, TC.TABLE_GENE).items())) == cap["Gene_cmpt"]
    # finish() without a new execute() would project into the same second-level tables again: refused
    from hisatgenotype_b200 import _lib
    with pytest.raises(_lib.HgtError):
        batch.finish()
    batch.close()
    for t in loci:
        t.close()


def test_key_list_overflow_falls_back_to_dense():
    """EM results normally come back as key lists (result_keys_kernel, <= 256 keys per unit); HGT_KEY_CAP=1 makes every
    unit with two keys overflow, so the same batch scenarios must pass through the dense fall-back of fetch_results.
    The capacity is read once per process, hence the child interpreter."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, HGT_KEY_CAP="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_typing.py"), "-x", "-q", "-m", "gpu",
                        "-k", "test_batch_matches_reference", "-p", "no:cacheprovider"], env=env, cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_batch_ranking_is_stable_over_runs():
    """The class rows of a batch land in the pool in scheduling order, so the EM sums run in a different row order every
    time; the ranked calls must not depend on it (alleles with identical membership columns tie EXACTLY in the reference,
    e.g. CYP2D6*06:01 / *06:06 of this scenario, and the tie decides their order)."""
    from hisatgenotype_b200 import typing_core as TC
    g = load_golden("cyp_pair")
    p = g["params"]
    db = golden_db(g)
    genes = []
    for cap in g["loci"]:
        if cap["gene"] not in genes:
            genes.append(cap["gene"])
    names = {cap["gene"]: cap["Gene_names"] for cap in g["loci"]}
    loci = [product_locus(g, db, gene, names[gene]) for gene in genes]
    for _ in range(60):
        batch = TC.Batch(loci, TC.make_params(p["num_editdist"], p["error_correction"], p["discordant"], p["simulation"],
                                              chunk_bytes=4000), p["remove_low"])
        for cap in g["loci"]:
            batch.add_unit(genes.index(cap["gene"]), cap["sam"])
        batch.run()
        em_i = 0
        for u in range(len(g["loci"])):
            for level in (0, 1):
                res = batch.unit_em(u, level)
                if res is None:
                    continue
                ref = g["em_calls"][em_i]["result"]
                em_i += 1
                assert [a for a, _ in res] == [a for a, _ in ref]
        batch.close()
    for t in loci:
        t.close()


def test_pipeline_matches_single_batches():
    """BatchPipeline (several batches in flight, one host thread + library context + CUDA stream each, loci shared):
    results come back in submission order and equal those of the batches run one at a time; pageable text and
    page-locked text (add_units_ptr) give the same."""
    from hisatgenotype_b200 import _lib
    from hisatgenotype_b200 import typing_core as TC
    g = load_golden("hla_pair_err")
    p = g["params"]
    db = golden_db(g)
    genes = []
    for cap in g["loci"]:
        if cap["gene"] not in genes:
            genes.append(cap["gene"])
    names = {cap["gene"]: cap["Gene_names"] for cap in g["loci"]}
    loci = [product_locus(g, db, gene, names[gene]) for gene in genes]
    params = TC.make_params(p["num_editdist"], p["error_correction"], p["discordant"], p["simulation"])
    caps = g["loci"]
    # batches of different composition: rotate and truncate the scenario's units
    batches = []
    for k in range(7):
        rot = caps[k % len(caps):] + caps[:k % len(caps)]
        batches.append([(genes.index(c["gene"]), c["sam"]) for c in rot[:1 + k % len(caps)]])

    def summarise(b):
        return [(b.unit_summary(u)["num_reads"], b.unit_summary(u)["num_pairs"],
                 list(map(list, b.unit_gene_cmpt(u, TC.TABLE_GENE).items()))) for u in range(len(b.unit_locus))], b.top_calls(4)

    want = []
    for units in batches:
        b = TC.Batch(loci, params, p["remove_low"])
        for li, sam in units:
            b.add_unit(li, sam)
        b.run()
        want.append(summarise(b))
        b.close()
    pipe = TC.BatchPipeline(loci, params, p["remove_low"], depth=3)
    got = list(pipe.map(batches, summarise))
    assert len(pipe.contexts()) >= 1
    # the same batches with the text in one page-locked allocation
    texts = [[(li, TC._sam_bytes(sam)) for li, sam in units] for units in batches]
    pinned = _lib.PinnedText(sum(len(t) + 16 for units in texts for _, t in units))
    ptr_batches = [[(li,) + pinned.add(t) for li, t in units] for units in texts]
    got_ptr = list(pipe.map(ptr_batches, summarise))
    pipe.close()
    pinned.close()
    for w, a, b in zip(want, got, got_ptr):
        assert w[0] == a[0] == b[0]
        for x, y, z in zip(w[1], a[1], b[1]):
            assert [n for n, _ in x] == [n for n, _ in y] == [n for n, _ in z]
            for (_, px), (_, py), (_, pz) in zip(x, y, z):
                assert px == pytest.approx(py, rel=1e-9) and px == pytest.approx(pz, rel=1e-9)
    for t in loci:
        t.close()


def test_batch_ragged_and_empty_units():
    """Units of very different sizes in one batch - an empty one, a single record, an odd cut through a pair, a full
    sample - and a malformed record: each unit's totals and tables equal the oracle's on the same lines; the malformed
    batch is refused with the reference's assertion (HGT_ERR_PARSE), not typed."""
    from hisatgenotype_b200 import _lib
    from hisatgenotype_b200 import typing_core as TC
    from hisatgenotype_b200.locus import LocusTables
    from helpers import assert_tables_equal_oracle, synthetic_case
    args, sam, truth = synthetic_case(41, 500, L=2500, n_pairs=400, base="hla", del_frac=0.1)
    ol = O.OracleLocus(*args)
    t = LocusTables(*args)
    cuts = [[], sam[:1], sam[:7], sam, sam[100:101], sam[200:460]]
    batch = TC.Batch([t], TC.make_params(), True)
    for lines in cuts:
        batch.add_unit(0, lines)
    batch.run()
    for u, lines in enumerate(cuts):
        ref = O.type_locus(ol, lines, simulation=False)
        s = batch.unit_summary(u)
        assert s["num_reads"] == ref["num_reads"] and s["num_pairs"] == ref["num_pairs"], (u, s, ref["num_reads"])
        assert_tables_equal_oracle(lambda tb: (list(map(list, batch.unit_gene_cmpt(u, tb).items())),
                                               batch.unit_gene_counts(u, tb)), ref, ol, True)
    batch.close()
    bad = "\t".join(sam[3].split("\t")[:8])
    batch = TC.Batch([t], TC.make_params(), True)
    batch.add_unit(0, sam[:3] + [bad] + sam[4:20])
    with pytest.raises((_lib.HgtError, AssertionError)):
        batch.run()
    batch.close()
    t.close()
