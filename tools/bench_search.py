#!/usr/bin/env python
"""Exact top-1001 search throughput over the shard shapes of the other BASELINE.json configs and query-batch sizes
(Q in {1, 16, 128, 1024}; N x D per GPU = 1M x 768, 1M x 2048 (config 3), 125k x 4096 (config 4 on 8
GPUs), 1.25M x 4096 (config 5 on 8 GPUs)).  Device-timed with CUDA events around `CorpusShard.search`, corpus resident,
working set > L2 except the 125k case (noted).  Not a bench line; output goes to profiles/ as supporting evidence.

    python tools/bench_search.py [--steps 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sgpt_b200 import CorpusShard  # noqa: E402

SHAPES = [("config 1/2: 1M x 768", 1_000_000, 768), ("config 3: 1M x 2048", 1_000_000, 2048),
          ("config 4 per GPU (1M/8): 125k x 4096", 125_000, 4096), ("config 5 per GPU (10M/8): 1.25M x 4096", 1_250_000, 4096),
          ("north-star scaling anchor: 10M x 768 on one GPU", 10_000_000, 768)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--queries", default="1,16,128,1024")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    hbm = 6561.6
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = json.load(open(p))["hbm_gbs"]
    kk = 1001
    for label, n, D in SHAPES:
        g = torch.Generator(device=dev).manual_seed(7)
        shard = CorpusShard(D, n, device=dev)
        for s0 in range(0, n, 250_000):
            shard.add(torch.randn(min(250_000, n - s0), D, generator=g, device=dev))
        for nq in [int(x) for x in args.queries.split(",")]:
            q = torch.randn(nq, D, generator=g, device=dev)
            for _ in range(3):
                shard.search(q, kk, "cos_sim")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                s, i = shard.search(q, kk, "cos_sim")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            scans = (nq + 127) // 128  # the shard is streamed once per block of 128 queries
            gbs = scans * (n * D * 2 + n * 4) / (ms / 1e3) / 1e9
            print(json.dumps({"shard": label, "docs": n, "dim": D, "queries": nq, "top_k": kk - 1, "ms_per_search": round(ms, 4),
                              "queries_per_s": round(nq / (ms / 1e3), 1), "corpus_GBps": round(gbs, 1),
                              "frac_of_hbm_peak": round(gbs / hbm, 3), "tflops": round(2 * nq * n * D / (ms / 1e3) / 1e12, 1),
                              "fits_l2": n * D * 2 < 126e6}), flush=True)
        del shard
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
