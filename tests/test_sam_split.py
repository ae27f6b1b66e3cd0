"""Native SAM intake (csrc/sam_split.cpp behind hisat-genotype_b200/sam_intake.py, host only): the per-locus, name-grouped
text it produces equals what the reference's pipe `samtools view <bam> <backbone> | sort -k1,1 -s` fed its per-read loop
(the `sam` lists captured from the unmodified reference in tests/golden)."""
import random

import pytest

from conftest import GOLDEN_NAMES, load_golden


def coordinate_sorted(caps):
    """The body of the coordinate-sorted alignment file the reference's pipe starts from: all loci, by (locus, POS), ties in
    the captured order (both sorts of the reference are stable)."""
    lines = []
    for cap in caps:
        body = list(cap["sam"])
        body.sort(key=lambda l: int(l.split("\t", 4)[3]))  # stable
        lines += body
    return lines


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_split_reproduces_the_reference_pipe(name):
    from hisatgenotype_b200.sam_intake import split_sam
    g = load_golden(name)
    n_loci = len(g["params"]["loci"])
    caps = g["loci"][:n_loci]  # the loci of the first test share one alignment file
    text = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:1\n" % c["ref_allele"] for c in caps)
    text += "\n".join(coordinate_sorted(caps)) + "\n"
    refs = [c["ref_allele"] for c in caps]
    out = split_sam(text.encode(), refs, n_threads=3)
    for c in caps:
        assert out[c["ref_allele"]].decode().splitlines() == c["sam"]


def test_split_order_rule_any_input_order_and_threads():
    """(name bytewise, position, input order) for shuffled input, several thread counts, a last line without newline, records
    of other references dropped."""
    from hisatgenotype_b200.sam_intake import split_sam
    rng = random.Random(3)
    recs = []
    for i in range(3000):
        nm = "r%d%s" % (rng.randrange(400), rng.choice(["", "|x", "A", "_2"]))
        ref = rng.choice(["L1*BACKBONE", "L2*BACKBONE", "other"])
        recs.append("%s\t%d\t%s\t%d\t60\t100M\t=\t1\t0\tACGT\tIIII\tNM:i:0\tMD:Z:100\tNH:i:1" % (nm, 99, ref, rng.randrange(1, 50)))
    text = "\n".join(recs)  # no trailing newline
    want = {}
    for ref in ("L1*BACKBONE", "L2*BACKBONE"):
        body = [r for r in recs if r.split("\t")[2] == ref]
        body.sort(key=lambda l: (l.split("\t")[0].encode(), int(l.split("\t")[3])))  # stable: ties keep the input order
        want[ref] = body
    for nt in (1, 2, 7):
        out = split_sam(text.encode(), ["L1*BACKBONE", "L2*BACKBONE"], n_threads=nt)
        for ref in want:
            assert out[ref].decode().splitlines() == want[ref]
            assert out[ref].endswith(b"\n") or not want[ref]


def test_split_rejects_malformed_and_handles_empty():
    from hisatgenotype_b200.sam_intake import split_sam
    assert split_sam(b"", ["X"]) == {"X": b""}
    assert split_sam(b"@HD\tVN:1.0\n", ["X"]) == {"X": b""}
    with pytest.raises(AssertionError):
        split_sam(b"read1\t99\tX\n", ["X"])
    with pytest.raises(AssertionError):
        split_sam(b"read1\t99\tX\tabc\t60\n", ["X"])


@pytest.mark.parametrize("name", GOLDEN_NAMES[:3])
def test_split_drop_qual(name):
    """drop_qual: every record keeps its columns except QUAL, which becomes '*'; records without a QUAL column or with '*'
    already are untouched; and the record stage does not care (same tables from the host emulation of the device code)."""
    import hgt_oracle as O
    from helpers import golden_db, pileup_arrays, product_locus, tables_from_jobs, oracle_locus
    from hisatgenotype_b200.sam_intake import split_sam
    from hisatgenotype_b200.typing_core import HostWalk, make_params
    g = load_golden(name)
    p = g["params"]
    n_loci = len(p["loci"])
    caps = g["loci"][:n_loci]
    text = "\n".join(coordinate_sorted(caps)) + "\n"
    refs = [c["ref_allele"] for c in caps]
    out = split_sam(text.encode(), refs, n_threads=2, drop_qual=True)
    db = golden_db(g)
    saved = 0
    for c in caps:
        got = out[c["ref_allele"]].decode().splitlines()
        assert len(got) == len(c["sam"])
        for a, b in zip(got, c["sam"]):
            ca, cb = a.split("\t"), b.split("\t")
            assert ca[:10] == cb[:10] and ca[11:] == cb[11:] and ca[10] == "*"
            saved += len(b) - len(a)
        # the device code's host emulation on the stripped lines: same haplotype jobs -> same tables as the reference
        loc = oracle_locus(g, db, c["gene"], c["Gene_names"])
        counts, nt_sets = O.get_mpileup(got, len(loc.ref_seq), 0, p["discordant"])
        cnt, mask = pileup_arrays(counts, nt_sets)
        t = product_locus(g, db, c["gene"], c["Gene_names"], host_only=True)
        walk = HostWalk(t, got, make_params(p["num_editdist"], p["error_correction"], p["discordant"], p["simulation"]), cnt, mask)
        res = tables_from_jobs(loc, walk, p["base"] == "hla")
        assert walk.num_reads == c["num_reads"] and walk.num_pairs == c["num_pairs"]
        assert res[0][0] == c["Gene_cmpt"] and res[0][1] == c["Gene_counts"]
        t.close()
    assert saved > 0
    # no QUAL column at all / already '*': nothing changes
    short = b"r1\t0\tX\t5\t60\t4M\t*\t0\t0\tACGT\n" b"r2\t0\tX\t7\t60\t4M\t*\t0\t0\tACGT\t*\tNM:i:0\n"
    assert split_sam(short, ["X"], drop_qual=True)["X"] == short
