#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*.csv).

    python tools/summarize_launches.py profiles/r01_launches_final.csv [title] > profiles/r01_launch_summary_final.txt

ncu serialises kernels and replays them cold, so the absolute durations are not bench values; the SHARE of each
kernel in the window is what must agree with bench.py's `kernel_ms_per_step`.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(line for line in open(path) if line.startswith('"'))]
    hdr = rows[0]
    name_i, grid_i, val_i = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    unit_i = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    total, n = 0.0, 0
    for r in rows[1:]:
        if len(r) <= val_i:
            continue
        v = float(r[val_i].replace(",", ""))
        v = {"ns": v * 1e-3, "us": v, "ms": v * 1e3}.get(r[unit_i], v * 1e-3)  # -> microseconds
        name = re.sub(r"^void |\(.*$|sgpt::", "", r[name_i])
        key = (name, r[grid_i])
        agg[key][0] += 1
        agg[key][1] += v
        total += v
        n += 1
    print(f"# {title}")
    print("# window of consecutive launches (cold-cache, serialised replay): compare SHARES, not absolutes")
    print(f"# total {total / 1e3:.3f} ms over {n} launches")
    for (name, grid), (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:86]:86s} {grid:>14s} n={cnt:4d} avg={us / cnt:9.1f} us share={100 * us / total:5.1f}%")


if __name__ == "__main__":
    main()
