#!/usr/bin/env python3
"""Run the UNMODIFIED reference typing path (imported from /root/reference) and capture what the
parity tests need.  TEST RIG ONLY: it runs in the build container (the GPU box has no /root/reference);
its outputs are committed as fixtures under tests/golden/ by make_goldens.py.

Captured per locus (hooks only observe; nothing in the reference is changed):
  * the alignment lines exactly as the per-read loop receives them
    (`samtools view <aln> <backbone> | sort -k1,1 -s`, reference hisatgenotype_typing_core.py:436-468)
  * Gene_cmpt / Gene_counts for the three tables just before ranking (core:1650) via a line tracer on
    typing()'s own frame, plus num_reads / num_pairs (core:1168, 1240, 1546)
  * every single_abundance call (arguments and result; common:1282-1410)
  * get_mpileup output (nt_set + counts; common:1059-1184)
  * get_alternatives output (common:1424-1657) and a sample of identify_ambigious_diffs calls
    (common:1663-1955)
  * the .report text
Environment expected: PYTHONHASHSEED=0, LC_ALL=C, PATH with hisat2 + the stub samtools, scratch CWD.
"""
import argparse
import copy
import json
import os
import subprocess
import sys

REF = os.environ.get("HGT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "hisatgenotype_modules"))

import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import hisatgenotype_typing_common as typing_common  # noqa: E402
import hisatgenotype_typing_core as typing_core  # noqa: E402

CAP = {"loci": [], "em_calls": [], "iad": [], "mpileup": [], "alts": []}
TYPING_CODE = typing_core.typing.__code__
_SNAPPED = set()
SNAP_LINE = 1650  # `Gene_counts = [[allele, count] ...` — tables are final here


def _view_sorted(alignment_fname, ref_allele):
    p1 = subprocess.Popen(["samtools", "view", alignment_fname, ref_allele], stdout=subprocess.PIPE,
                          universal_newlines=True)
    p2 = subprocess.Popen(["sort", "-k", "1,1", "-s"], stdin=p1.stdout, stdout=subprocess.PIPE,
                          universal_newlines=True)
    return p2.stdout.read().splitlines()


def _local_tracer(frame, event, arg):
    if event == "line" and frame.f_lineno == SNAP_LINE:
        loc = frame.f_locals
        key = (id(frame), loc.get("gene"), id(loc.get("Gene_cmpt")))
        if loc.get("index_type") == "graph" and isinstance(loc.get("Gene_counts"), dict) and key not in _SNAPPED:
            _SNAPPED.add(key)
            snap = {
                "gene": loc["gene"],
                "ref_allele": loc["ref_allele"],
                "num_reads": loc["num_reads"],
                "num_pairs": loc["num_pairs"],
                "Gene_cmpt": list(loc["Gene_cmpt"].items()),
                "Gene_counts": list(loc["Gene_counts"].items()),
                "Gene_exons_cmpt": list(loc["Gene_exons_cmpt"].items()),
                "Gene_exons_counts": list(loc["Gene_exons_counts"].items()),
                "Gene_primary_exons_cmpt": list(loc["Gene_primary_exons_cmpt"].items()),
                "Gene_primary_exons_counts": list(loc["Gene_primary_exons_counts"].items()),
                "allele_rep_groups": {k: list(v) for k, v in loc["allele_rep_groups"].items()},
                "primary_exon_allele_rep_groups": {k: list(v) for k, v in
                                                   loc["primary_exon_allele_rep_groups"].items()},
                "Gene_names": list(loc["Gene_names"][loc["gene"]]),
                "sam": _view_sorted(loc["alignment_fname"], loc["ref_allele"]),
                "test_Gene_names": loc["test_Gene_names"] if loc["simulation"] else None,
            }
            CAP["loci"].append(snap)
    return _local_tracer


def _tracer(frame, event, arg):
    if event == "call" and frame.f_code is TYPING_CODE:
        return _local_tracer
    return None


_orig_em = typing_common.single_abundance


def _em(Gene_cmpt, remove_low_abundance_allele=False, Gene_length={}):
    res = _orig_em(Gene_cmpt, remove_low_abundance_allele, Gene_length)
    used = set()
    for k in Gene_cmpt:
        used.update(k.split("-"))
    CAP["em_calls"].append({
        "cmpt": list(Gene_cmpt.items()),
        "remove_low": bool(remove_low_abundance_allele),
        "lengths": {a: Gene_length[a] for a in used} if len(Gene_length) > 0 else {},
        "result": [[a, p] for a, p in res],
    })
    return res


typing_common.single_abundance = _em

_orig_iad = typing_common.identify_ambigious_diffs
_iad_budget = {"plain": 40, "alt": 400}


def _iad(ref_seq, Vars, Alts_left, Alts_right, Alts_left_list, Alts_right_list, cmp_list, verbose, debug=False):
    inp = copy.deepcopy(cmp_list)
    res = _orig_iad(ref_seq, Vars, Alts_left, Alts_right, Alts_left_list, Alts_right_list, cmp_list, verbose, debug)
    l, r, la, ra = res
    trivial = (l == 0 and r == len(cmp_list) - 1 and len(la) == 1 and len(ra) == 1)
    kind = "plain" if trivial else "alt"
    if _iad_budget[kind] > 0:
        _iad_budget[kind] -= 1
        CAP["iad"].append({"cmp_list": inp, "result": [l, r, sorted(la), sorted(ra)]})
    return res


typing_common.identify_ambigious_diffs = _iad

_orig_mp = typing_common.get_mpileup


def _mp(alignview_cmd, ref_seq, base_locus, vars, allow_discordant):
    res = _orig_mp(alignview_cmd, ref_seq, base_locus, vars, allow_discordant)
    CAP["mpileup"].append({
        "nt_set": ["".join(sorted(e[0])) for e in res],
        "counts": [{nt: v[0] for nt, v in e[1].items()} for e in res],
    })
    return res


typing_common.get_mpileup = _mp

_orig_alts = typing_common.get_alternatives


def _alts(ref_seq, allele_vars, Vars, Var_list, verbose):
    l, r = _orig_alts(ref_seq, allele_vars, Vars, Var_list, verbose)
    CAP["alts"].append({"left": {k: sorted(v) for k, v in l.items()}, "right": {k: sorted(v) for k, v in r.items()}})
    return l, r


typing_common.get_alternatives = _alts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ix", required=True)
    ap.add_argument("--base", default="hla")
    ap.add_argument("--loci", default="A")
    ap.add_argument("--out", required=True)
    ap.add_argument("--outdir", default="out")
    ap.add_argument("--debug", default="basic,test_size:1,set_seed:101")
    ap.add_argument("--reads", default="")  # comma separated read files -> real-read mode
    ap.add_argument("--interval", type=int, default=10)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--frag-len", type=int, default=350)
    ap.add_argument("--editdist", type=int, default=2)
    ap.add_argument("--err", type=float, default=0.0)
    ap.add_argument("--snp", type=float, default=0.0)
    ap.add_argument("--no-error-correction", action="store_true")
    ap.add_argument("--discordant", action="store_true")
    ap.add_argument("--keep-low", action="store_true")
    ap.add_argument("--best-alleles", action="store_true")
    ap.add_argument("--all-counts", action="store_true")
    ap.add_argument("--threads", type=int, default=4)
    args = ap.parse_args()

    debug_instr = {}
    if args.debug:
        for item in args.debug.split(","):
            if ":" in item:
                k, v = item.split(":")
                debug_instr[k] = v
            else:
                debug_instr[item] = None
    read_fname = [r for r in args.reads.split(",") if r]
    os.makedirs(args.outdir, exist_ok=True)
    sys.argv = ["hisatgenotype", "--base", args.base, "--locus-list", args.loci]
    sys.settrace(_tracer)
    try:
        typing_core.genotyping_locus(
            args.base, args.loci.split(","), "", args.ix, [], True, [["hisat2", "graph"]], read_fname, False, "",
            args.threads, args.interval, args.read_len, args.frag_len, args.best_alleles, args.editdist, args.err,
            args.snp, [], False, "assembly_graph", not args.no_error_correction, True, args.discordant, False,
            not args.keep_low, [], 0, False, args.outdir, args.all_counts, debug_instr if not read_fname else {})
    finally:
        sys.settrace(None)
    reports = {}
    for fn in sorted(os.listdir(args.outdir)):
        if fn.endswith(".report"):
            reports[fn] = open(os.path.join(args.outdir, fn)).read()
    CAP["reports"] = reports
    CAP["params"] = {
        "base": args.base, "loci": args.loci.split(","), "simulation": not read_fname,
        "num_editdist": args.editdist, "error_correction": not args.no_error_correction,
        "discordant": args.discordant, "remove_low": not args.keep_low, "best_alleles": args.best_alleles,
        "output_allele_counts": args.all_counts, "debug": args.debug, "read_fname": read_fname,
        "interval": args.interval, "read_len": args.read_len, "frag_len": args.frag_len,
        "err": args.err, "snp": args.snp,
    }
    with open(args.out, "w") as fo:
        json.dump(CAP, fo)


if __name__ == "__main__":
    main()
