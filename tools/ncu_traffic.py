#!/usr/bin/env python
"""DRAM traffic per launch, by kernel class, from an ncu CSV that carries dram__bytes_read.sum / dram__bytes_write.sum
(and gpu__time_duration.sum) for a window of bench.py launches:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -s 500 -c 420 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 6 --warmup 2 --no-cpu-baseline
    python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r01_dram_traffic.json

bench.py reads the JSON (newest profiles/*dram_traffic.json) to fill `roofline.traffic` / `roofline_similarity.traffic`:
measured DRAM bytes per launch next to the algorithmic bytes.  ncu replays kernels one at a time with cold caches, so
these bytes are an upper bound of what a launch moves inside the pipelined step (where part of the previous kernel's
output is still in the 126 MB L2).
"""
import csv
import json
import re
import sys
from collections import defaultdict

CLASSES = [  # (class, regex on the demangled kernel name) — first match wins
    ("similarity_gemm", r"gemm_bf16_tn_kernel<.*Epi(FilterRows|Staged|Scores)"),
    ("linear_gemm", r"gemm_bf16_tn_kernel"),
    ("attention", r"attention_"),
    ("layernorm", r"layernorm_"),
    ("topk", r"topk_|tau_select|merge_sorted"),
    ("pool", r"pool_|row_stats_kernel|l2_scale_rows"),
    ("embed", r"embed_kernel"),
]


def unit_scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(line for line in open(src) if line.startswith('"'))]
    hdr = rows[0]
    idx = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = defaultdict(dict)
    names = {}
    for r in rows[1:]:
        if len(r) <= idx["Metric Value"]:
            continue
        v = float(r[idx["Metric Value"]].replace(",", "")) * unit_scale(r[idx["Metric Unit"]])
        launches[r[idx["ID"]]][r[idx["Metric Name"]]] = v
        names[r[idx["ID"]]] = r[idx["Kernel Name"]]
    agg = defaultdict(lambda: {"launches": 0, "read": 0.0, "write": 0.0, "us": 0.0, "kernels": defaultdict(int)})
    for lid, m in launches.items():
        name = re.sub(r"^void |sgpt::", "", names[lid])
        cls = next((c for c, pat in CLASSES if re.search(pat, name)), "misc")
        a = agg[cls]
        a["launches"] += 1
        a["read"] += m.get("dram__bytes_read.sum", 0.0)
        a["write"] += m.get("dram__bytes_write.sum", 0.0)
        a["us"] += m.get("gpu__time_duration.sum", 0.0)
        a["kernels"][re.sub(r"\(.*$", "", name)[:90]] += 1
    out = {"source": src, "note": "ncu replay (cold caches, serialised): bytes are per launch, averaged over the window"}
    for cls, a in agg.items():
        n = a["launches"]
        out[cls] = {"launches": n, "dram_read_bytes_per_launch": a["read"] / n, "dram_write_bytes_per_launch": a["write"] / n,
                    "dram_bytes_per_launch": (a["read"] + a["write"]) / n, "avg_us_under_ncu": a["us"] / n,
                    "kernels": dict(a["kernels"])}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    for cls in sorted(k for k in out if isinstance(out[k], dict)):
        o = out[cls]
        print(f"{cls:16s} n={o['launches']:4d}  {o['dram_bytes_per_launch'] / 1e6:9.1f} MB/launch "
              f"(r {o['dram_read_bytes_per_launch'] / 1e6:.1f} / w {o['dram_write_bytes_per_launch'] / 1e6:.1f})  "
              f"{o['avg_us_under_ncu']:.1f} us")


if __name__ == "__main__":
    main()
