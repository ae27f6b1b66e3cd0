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
    for r in rows[2:]:
        yield (r[kn], float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]], r[tm] + " " + units[tm])


def main():
    workload, size, reps = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    groups = {"em_kernel": [], "stage_a": [], "records": []}
    for rep in reps:
        for name, b, t in launches(rep):
            if "em_kernel" in name:
                groups["em_kernel"].append((name, b, t))
            elif "compat_kernel" in name or "class_kernel" in name:
                groups["stage_a"].append((name, b, t))
            elif "hgtk::" in name or name.split("(")[0] in ("pileup_flags_kernel",):
                groups["records"].append((name, b, t))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    for g, ls in groups.items():
        if not ls:
            continue
        data["%s:%s:%d" % (workload, g, size)] = {
            "dram_bytes_per_launch": sum(b for _, b, _ in ls) / len(ls), "launches_captured": len(ls),
            # the captures are taken with `bench.py --steps 1 --warmup 0` and a launch count that covers exactly the step
            "dram_bytes_per_step": sum(b for _, b, _ in ls),
            "launch_times": [t for _, _, t in ls][:12], "reports": [os.path.basename(r) for r in reps]}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
