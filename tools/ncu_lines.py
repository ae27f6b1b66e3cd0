"""Per-source-line view of a kernel from an `ncu --set full --import-source on` report (kernels built with -lineinfo).

usage: tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N] [launch id]
Reads `ncu -i <rep> --page source --csv --print-source cuda,sass` for the kernels matching the regex, sums the warp
stall samples, executed warp instructions, thread instructions and L1 tag requests of the SASS lines behind every CUDA
source line and prints the top N lines by samples (used for profiles/*_lines.txt).
"""
import collections
import csv
import os
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_line = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
stall_cols = {}
stalls = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        si, ie, te = r.index("# Samples"), r.index("Instructions Executed"), r.index("Thread Instructions Executed")
        tg = r.index("L1 Tag Requests Global")
        stall_cols = {i: n for i, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n}
        continue
    if r[0] != "":
        cur_line = r[0]
        continue
    try:
        s, e, t = int(r[si]), int(r[ie]), int(r[te])
    except (ValueError, IndexError):
        continue
    try:
        g = int(r[tg])
    except ValueError:
        g = 0
    k = (cur_file, int(cur_line))
    a = agg[k]
    a[0] += s; a[1] += e; a[2] += t; a[3] += g
    for i, n in stall_cols.items():
        try:
            stalls[n] += int(r[i])
        except ValueError:
            pass
tots = sum(v[0] for v in agg.values()) or 1
tote = sum(v[1] for v in agg.values()) or 1
tott = sum(v[2] for v in agg.values())
totg = sum(v[3] for v in agg.values()) or 1
print("samples %d  warp instr %.1fM  thread instr %.1fM (%.1f active lanes / instr)  L1 tag requests %.1fM" % (
    tots, tote / 1e6, tott / 1e6, tott / tote, totg / 1e6))
print("stalls: " + ", ".join("%s %.0f%%" % (n[6:], 100 * c / tots) for n, c in stalls.most_common(7)))
cache = {}
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    if f not in cache:
        cache[f] = open(f).read().splitlines() if os.path.exists(f) else []
    src = cache[f][l - 1].strip()[:90] if l - 1 < len(cache[f]) else ""
    print("%-14s %5d  %5.1f%%smp %5.1f%%ins %5.1f%%tag  %4.1f lanes  %s" % (
        os.path.basename(f), l, 100 * v[0] / tots, 100 * v[1] / tote, 100 * v[3] / totg, v[2] / max(v[1], 1), src))
