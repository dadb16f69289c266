#!/usr/bin/env python
"""Where does the end-to-end search time go?  Wall-clock per step of CorpusShard.search through the public API with the
host<->device copies added one at a time (bench.py's e2e leg measured 2.7 ms per search step against 0.5 ms on the
device).  Run on a GPU box:  python tools/e2e_search_probe.py [n_docs]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgpt_b200 import CorpusShard  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D, NQ, KK, STEPS = 768, 128, 1001, 40
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
shard = CorpusShard(D, N, device=dev)
for s0 in range(0, N, 100_000):
    shard.add(torch.randn(min(100_000, N - s0), D, generator=g, device=dev))
q_dev = torch.randn(NQ, D, generator=g, device=dev)
q_host = q_dev.cpu().pin_memory()
s_host = [torch.empty((NQ, KK), dtype=torch.float32).pin_memory() for _ in range(2)]
i_host = [torch.empty((NQ, KK), dtype=torch.int64).pin_memory() for _ in range(2)]
evt = [torch.cuda.Event() for _ in range(2)]


def run(name, fn):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(STEPS):
        fn(k)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"{name:58s} {1e3 * t / STEPS:7.3f} ms/step  (host issue {1e3 * t_issue / STEPS:6.3f} ms/step)", flush=True)


def dev_only(k):
    shard.search(q_dev, KK, "cos_sim")


def with_h2d(k):
    shard.search(q_host.to(dev, non_blocking=True), KK, "cos_sim")


def with_d2h(k):
    slot = k % 2
    evt[slot].synchronize()
    s, i = shard.search(q_dev, KK, "cos_sim")
    s_host[slot].copy_(s, non_blocking=True)
    i_host[slot].copy_(i, non_blocking=True)
    evt[slot].record()


def both(k):
    slot = k % 2
    evt[slot].synchronize()
    s, i = shard.search(q_host.to(dev, non_blocking=True), KK, "cos_sim")
    s_host[slot].copy_(s, non_blocking=True)
    i_host[slot].copy_(i, non_blocking=True)
    evt[slot].record()


def both_blocking(k):
    s, i = shard.search(q_host.to(dev), KK, "cos_sim")
    s.cpu(), i.cpu()


def copies_only(k):
    slot = k % 2
    qd = q_host.to(dev, non_blocking=True)
    s_host[slot].copy_(qd[:, :KK - 233].contiguous().new_empty((NQ, KK)), non_blocking=True)
    i_host[slot].copy_(torch.empty((NQ, KK), dtype=torch.int64, device=dev), non_blocking=True)


run("search, queries resident, results stay on device", dev_only)
run("+ H2D queries (pinned, non_blocking)", with_h2d)
run("+ D2H scores+ids (pinned, non_blocking, 2 slots)", with_d2h)
run("+ both", both)
run("blocking .to / .cpu() (what a naive caller does)", both_blocking)
run("the copies alone (no search)", copies_only)
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
bd = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("H2D 64 MiB pinned", lambda: bd.copy_(big, non_blocking=True)), ("D2H 64 MiB pinned", lambda: big.copy_(bd, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"{name}: {(64 << 20) / t / 1e9:.1f} GB/s")
