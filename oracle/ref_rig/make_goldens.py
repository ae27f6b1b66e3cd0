#!/usr/bin/env python3
"""Generate tests/golden/*.json.gz by running the unmodified reference (run_reference.py) on small
synthetic databases, aligned with the reference's own HISAT2 (built from the submodule into a scratch
directory, see oracle/ref_rig/README.md).  TEST RIG ONLY, container-side; the fixtures are committed.

usage: python oracle/ref_rig/make_goldens.py [--hisat2 /tmp/hgt_rig/hisat2] [--only name]
"""
import argparse
import gzip
import importlib.util
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
RIG = os.path.join(ROOT, "oracle", "ref_rig")

spec = importlib.util.spec_from_file_location("hgt_synth", os.path.join(ROOT, "hisat-genotype_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
sys.modules["hgt_synth"] = synth
spec.loader.exec_module(synth)

# name -> (database spec, run_reference arguments)
SCENARIOS = {
    # hg_test1_basic-style: one allele, no errors (devel/pre-int_test.sh:22)
    "hla_basic": dict(
        base="hla",
        loci=[dict(gene="A", seed=7, L=1600, n_alleles=48, n_groups=6, core_vars=30, pool_private=60)],
        run=["--loci", "A", "--debug", "basic,test_size:2,set_seed:101", "--all-counts"]),
    # hg_test2_paired-style: two alleles, two loci, sequencing errors + novel SNPs
    "hla_pair_err": dict(
        base="hla",
        loci=[dict(gene="A", seed=8, L=1800, n_alleles=64, n_groups=8, core_vars=35, pool_private=80),
              dict(gene="B", seed=9, L=1500, n_alleles=40, n_groups=5, core_vars=30, pool_private=50, del_frac=0.12)],
        run=["--loci", "A,B", "--debug", "pair,test_size:2,set_seed:100", "--err", "0.5", "--snp", "0.1",
             "--all-counts"]),
    # deletion-rich + insertions: exercises get_alternatives / identify_ambigious_diffs and indel lookups
    "hla_indel": dict(
        base="hla",
        loci=[dict(gene="C", seed=21, L=1500, n_alleles=48, n_groups=6, core_vars=40, pool_private=60,
                   del_frac=0.30, ins_frac=0.08)],
        run=["--loci", "C", "--debug", "pair,test_size:2,set_seed:5", "--err", "0.3", "--interval", "7",
             "--all-counts"]),
    # non-"hla" database name: single table, plain EM (core:1783-1789)
    "cyp_pair": dict(
        base="cyp",
        loci=[dict(gene="CYP2D6", seed=31, L=1700, n_alleles=36, n_groups=6, core_vars=25, pool_private=40,
                   del_frac=0.15)],
        run=["--loci", "CYP2D6", "--debug", "pair,test_size:2,set_seed:11", "--err", "0.4", "--all-counts"]),
    # single-end, no error correction, keep low-abundance alleles, --discordant
    "hla_single_end": dict(
        base="hla",
        loci=[dict(gene="A", seed=41, L=1400, n_alleles=40, n_groups=5, core_vars=30, pool_private=50,
                   del_frac=0.10)],
        run=["--loci", "A", "--debug", "pair,test_size:1,set_seed:3,single-end", "--err", "0.5",
             "--no-error-correction", "--keep-low", "--discordant", "--all-counts"]),
    # wide locus: 2,200 alleles = 35 words per allele set, two words per lane in the allele-set kernels and multi-word
    # paths everywhere (the other scenarios stay below 66 alleles); pins the oracle and the CUDA path at a width the
    # benchmark shapes use, against the unmodified reference
    "hla_wide": dict(
        base="hla",
        loci=[dict(gene="A", seed=61, L=2200, n_alleles=2200, n_groups=30, core_vars=60, pool_private=700,
                   del_frac=0.10)],
        run=["--loci", "A", "--debug", "pair,test_size:1,set_seed:23", "--err", "0.4", "--all-counts"]),
    # deeper coverage so that the pileup thresholds (depth>=20) and error correction actually fire
    "hla_deep": dict(
        base="hla",
        loci=[dict(gene="A", seed=51, L=1200, n_alleles=32, n_groups=4, core_vars=30, pool_private=40,
                   del_frac=0.12)],
        run=["--loci", "A", "--debug", "pair,test_size:1,set_seed:17", "--err", "1.0", "--snp", "0.2",
             "--interval", "2", "--all-counts"]),
}


def run_scenario(name, sc, hisat2_dir, keep=False):
    work = tempfile.mkdtemp(prefix="hgt_golden_%s_" % name)
    ix = os.path.join(work, "ix")
    os.makedirs(ix)
    for d in ["hisatgenotype_db", "grch38"]:
        os.makedirs(os.path.join(ix, d))
    for fn in ["genome.fa", "genome.fa.fai"]:
        open(os.path.join(ix, fn), "w").close()
    loci = [synth.make_locus(**spec_) for spec_ in sc["loci"]]
    synth.write_database(loci, sc["base"], ix)
    env = dict(os.environ)
    env["PATH"] = RIG + ":" + hisat2_dir + ":" + env["PATH"]
    env["PYTHONHASHSEED"] = "0"
    env["LC_ALL"] = "C"
    cwd = os.path.join(work, "cwd")
    os.makedirs(cwd)
    out = os.path.join(work, "capture.json")
    cmd = [sys.executable, "-W", "ignore", os.path.join(RIG, "run_reference.py"), "--ix", ix, "--base", sc["base"],
           "--out", out, "--outdir", os.path.join(cwd, "out")] + sc["run"]
    r = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + "\n" + r.stderr[-6000:] + "\n")
        raise SystemExit("reference run failed for %s (workdir %s)" % (name, work))
    cap = json.load(open(out))
    db = {}
    for ext in ["_backbone.fa", ".locus", ".snp", ".link", ".allele", ".partial"]:
        db[ext] = open(os.path.join(ix, sc["base"] + ext)).read()
    cap["db"] = db
    cap["name"] = name
    cap["stderr_tail"] = r.stderr[-2000:]
    dst = os.path.join(ROOT, "tests", "golden", name + ".json.gz")
    with gzip.open(dst, "wt", compresslevel=9) as fo:
        json.dump(cap, fo, separators=(",", ":"))
    n_sam = sum(len(l["sam"]) for l in cap["loci"])
    print("%-16s loci-runs=%d sam_lines=%d em_calls=%d iad=%d  -> %s (%.1f KB)" % (
        name, len(cap["loci"]), n_sam, len(cap["em_calls"]), len(cap["iad"]), os.path.relpath(dst, ROOT),
        os.path.getsize(dst) / 1024.0))
    for l in cap["loci"]:
        print("    %s reads=%d pairs=%d classes=%d/%d/%d" % (
            l["gene"], l["num_reads"], l["num_pairs"], len(l["Gene_cmpt"]), len(l["Gene_exons_cmpt"]),
            len(l["Gene_primary_exons_cmpt"])))
    if not keep:
        shutil.rmtree(work)
    else:
        print("    kept", work)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hisat2", default="/tmp/hgt_rig/hisat2")
    ap.add_argument("--only", default="")
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    for name, sc in SCENARIOS.items():
        if args.only and name not in args.only.split(","):
            continue
        run_scenario(name, sc, args.hisat2, args.keep)


if __name__ == "__main__":
    main()
