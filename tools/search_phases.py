#!/usr/bin/env python
"""Per-phase device time of the fused exact search (scan kernels vs selection kernels) for cos_sim and dot scoring on the
shard shapes that matter for the bars of VERDICT r01 item 5: 1 M x 768 (whole search >= 0.60 of HBM) and 1.25 M x 768
(what one of 8 GPUs holds of the 10 M corpus).  Whole-search time with CUDA events (unprofiled loop), the per-category
split from the library's event profiler (second loop).  Supporting evidence for profiles/, not a bench line.

    python tools/search_phases.py [--steps 30] [--shapes 1000000x768,1250000x768]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sgpt_b200 import CorpusShard, _lib  # noqa: E402

CATS = ["embed", "layernorm", "linear_gemm", "attention", "pool", "similarity_gemm", "topk", "misc"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--shapes", default="1000000x768,1250000x768")
    ap.add_argument("--queries", type=int, default=128)
    ap.add_argument("--k", type=int, default=1001)
    ap.add_argument("--scores", default="cos_sim,dot")
    ap.add_argument("--pre", default="none", choices=["none", "dirty", "burn"],
                    help="what runs before every search: nothing, a 192 MB memset (dirty L2 lines to write back), or ~3 ms "
                         "of bf16 matmuls (the power / clock state an encode step leaves behind)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    hbm = 6561.6
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = json.load(open(p))["hbm_gbs"]
    for shape in a.shapes.split(","):
        n, D = (int(x) for x in shape.split("x"))
        g = torch.Generator(device=dev).manual_seed(7)
        shard = CorpusShard(D, n, device=dev)
        for s0 in range(0, n, 250_000):
            shard.add(torch.randn(min(250_000, n - s0), D, generator=g, device=dev))
        q = torch.randn(a.queries, D, generator=g, device=dev)
        dirty = torch.empty(192 << 20, dtype=torch.uint8, device=dev) if a.pre == "dirty" else None
        ma = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16) if a.pre == "burn" else None

        def pre():
            if dirty is not None:
                dirty.zero_()
            if ma is not None:
                for _ in range(4):
                    torch.matmul(ma, ma)
        for sf in a.scores.split(","):
            for _ in range(3):
                pre()
                shard.search(q, a.k, sf)
            torch.cuda.synchronize()
            ms = 0.0
            for _ in range(a.steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                pre()
                e0.record()
                shard.search(q, a.k, sf)
                e1.record()
                if a.pre != "none":
                    torch.cuda.synchronize()
                    ms += e0.elapsed_time(e1) / a.steps
            if a.pre == "none":  # back-to-back searches: one pair of events around the whole loop
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    shard.search(q, a.k, sf)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.steps
            ms_cat = (ctypes.c_double * len(CATS))()
            n_cat = (ctypes.c_int64 * len(CATS))()
            lib.sgpt_profile_read(None, None, None)
            lib.sgpt_profile_enable(1)
            for _ in range(a.steps):
                pre()
                shard.search(q, a.k, sf)
            torch.cuda.synchronize()
            lib.sgpt_profile_enable(0)
            lib.sgpt_profile_read(ms_cat, n_cat, None)
            tl = (ctypes.c_uint64 * 11)()
            shard.search(q, a.k, sf)
            lib.sgpt_debug_topk_timeline(tl, 11)
            timeline = [round((tl[i] - tl[0]) / 1e3, 2) for i in range(11)]  # us since the final selection's entry
            bytes_ = n * D * 2 + (n * 4 if sf == "cos_sim" else 0)
            scan_ms = ms_cat[5] / a.steps
            print(json.dumps({
                "docs": n, "dim": D, "queries": a.queries, "k": a.k, "score": sf, "before_each_search": a.pre, "ms_per_search": round(ms, 4),
                "whole_search_frac_of_hbm": round(bytes_ / (ms / 1e3) / 1e9 / hbm, 3),
                "scan_ms": round(scan_ms, 4), "scan_launches": n_cat[5] // a.steps,
                "scan_frac_of_hbm": round(bytes_ / (scan_ms / 1e3) / 1e9 / hbm, 3) if scan_ms > 0 else None,
                "select_ms": round(ms_cat[6] / a.steps, 4), "select_launches": n_cat[6] // a.steps,
                "final_select_timeline_us": timeline,
                "other_ms": round(sum(ms_cat[i] for i in range(len(CATS)) if i not in (5, 6)) / a.steps, 4),
                "env": {k: v for k, v in os.environ.items() if k.startswith("SGPT_")}}), flush=True)
        del shard
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
