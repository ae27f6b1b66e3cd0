#!/usr/bin/env python
"""Per-launch DRAM traffic of the typing kernels from `ncu --set full` reports -> profiles/traffic.json.

usage: tools/ncu_traffic.py <workload> <size> <report.ncu-rep> [...]   (size = samples per step, or reads for 'oversized')
Each report is read with `ncu -i <rep> --page raw --csv`; launches are grouped as em_kernel / stage_a (compat + class
kernels of one locus = one stage (a) "launch group") and dram__bytes_read.sum + dram__bytes_write.sum is averaged per
launch.  bench.py copies the figure into roofline.traffic; numbers printed by a run under ncu are never bench values.
"""
import csv
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    rd, wr, tm = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    chip = {"shared_memory_wavefronts_pct_of_peak": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
    for r in rows[2:]:
        extra = {k: float(r[hdr.index(m)]) for k, m in chip.items() if m in hdr}
        yield (r[kn], float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]], r[tm] + " " + units[tm], extra)


def main():
    workload, size, reps = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    groups = {"em_kernel": [], "stage_a": [], "records": []}
    for rep in reps:
        for name, b, t, extra in launches(rep):
            if "em_kernel" in name:
                groups["em_kernel"].append((name, b, t, extra))
            elif "compat_kernel" in name or "class_kernel" in name or "class_sort_kernel" in name:
                groups["stage_a"].append((name, b, t, extra))
            elif "hgtk::" in name or name.split("(")[0] in ("pileup_flags_kernel",):
                groups["records"].append((name, b, t, extra))
    # record stage: one execution ends with pair_fill_kernel (a capture may run on into the next step's first kernels)
    for k, x in enumerate(groups["records"]):
        if "pair_fill_kernel" in x[0]:
            groups["records"] = groups["records"][:k + 1]
            break
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    for g, ls in groups.items():
        if not ls:
            continue
        data["%s:%s:%d" % (workload, g, size)] = {
            "dram_bytes_per_launch": sum(x[1] for x in ls) / len(ls), "launches_captured": len(ls),
            # the captures are taken with `bench.py --steps 1 --warmup 0` and a launch count that covers exactly the step
            "dram_bytes_per_step": sum(x[1] for x in ls),
            "launch_times": [x[2] for x in ls][:12], "reports": [os.path.basename(r) for r in reps]}
        if g == "em_kernel":  # the problems live in shared memory: distance to the on-chip ceilings, per launch
            keys = sorted(ls[0][3])
            data["%s:%s:%d" % (workload, g, size)]["on_chip"] = dict(
                {k: [round(x[3].get(k, float("nan")), 1) for x in ls] for k in keys},
                launches=[x[0].split("(")[0].replace("void <unnamed>::", "") for x in ls],
                source="ncu --set full of " + ", ".join(os.path.basename(r) for r in reps))
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
