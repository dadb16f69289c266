#!/usr/bin/env python
"""Localise the gap between bench.py's device-timed step (encode 6.9 ms + search 0.5 ms) and its end-to-end step with
search (9.5 ms): replays the e2e loop of bench.py with the pieces switched on one at a time, reporting wall time per
step, host issue time per step and the GPU time between the first and last kernel of a step (CUDA events)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sgpt_b200 import CorpusShard, Encoder, preset  # noqa: E402

B, S, NQ, D, KK, STEPS = 256, 128, 128, 768, 1001, 24
dev = torch.device("cuda:0")
enc = Encoder(preset("sgpt-125m"), bench.synthetic_weights(0), device=dev, max_tokens=B * S, max_batch=B)
batches = bench.token_batches(4, seed=1)
mask = np.ones((B, S), dtype=np.int8)
g = torch.Generator(device=dev).manual_seed(1)
N = 1_000_000
sampler = bench.ClockSampler(0) if "--sampler" in sys.argv else None  # the nvidia-smi -lms 100 poller bench.py runs
if sampler:
    time.sleep(1.0)
shard = CorpusShard(D, N, device=dev)
for s0 in range(0, N, 100_000):
    shard.add(torch.randn(100_000, D, generator=g, device=dev))
q_dev = torch.randn(NQ, D, generator=g, device=dev)
q_host = q_dev.cpu().pin_memory()
emb_host = [torch.empty((B, D), dtype=torch.float32).pin_memory() for _ in range(2)]
s_host = [torch.empty((NQ, KK), dtype=torch.float32).pin_memory() for _ in range(2)]
i_host = [torch.empty((NQ, KK), dtype=torch.int64).pin_memory() for _ in range(2)]
slot_evt = [torch.cuda.Event() for _ in range(2)]


def loop(name, search, h2d_q, d2h_res, d2h_emb=True, slots=True):
    def step(k, ev=None):
        slot = k % 2
        if slots:
            slot_evt[slot].synchronize()
        if ev:
            ev[0].record()
        emb = enc.encode_tokens(batches[k % 4].numpy(), mask)
        if d2h_emb:
            emb_host[slot].copy_(emb, non_blocking=True)
        if search:
            qd = q_host.to(dev, non_blocking=True) if h2d_q else q_dev
            s, i = shard.search(qd, KK, "cos_sim")
            if d2h_res:
                s_host[slot].copy_(s, non_blocking=True)
                i_host[slot].copy_(i, non_blocking=True)
        if ev:
            ev[1].record()
        slot_evt[slot].record()

    for k in range(4):
        step(k)
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(STEPS)]
    t0 = time.perf_counter()
    for k in range(STEPS):
        step(k, evs[k])
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    gpu = sum(a.elapsed_time(b) for a, b in evs) / STEPS
    print(f"{name:64s} wall {1e3 * t / STEPS:7.3f} ms/step | host issue {1e3 * t_issue / STEPS:7.3f} | GPU first->last {gpu:7.3f}",
          flush=True)


print("nvidia-smi sampler:", "ON" if sampler else "off", flush=True)
loop("encode e2e (H2D ids, D2H emb)", False, False, False)
loop("encode e2e + search (device queries, results stay)", True, False, False)
loop("encode e2e + search + H2D queries", True, True, False)
loop("encode e2e + search + D2H results", True, False, True)
loop("encode e2e + search + both (= bench e2e step)", True, True, True)
loop("same, without the 2-slot event wait", True, True, True, slots=False)
loop("encode (no D2H emb) + search + both", True, True, True, d2h_emb=False)
if sampler:
    print(sampler.stop())
