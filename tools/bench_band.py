#!/usr/bin/env python
"""Encoder GEMM shapes with the tile rasterisation forced to N-fastest (SGPT_GEMM_BAND=0) or banded M-fastest with band
heights 4 / 8 / 16 (gemm.cuh TileMap::band; read per call) against the automatic choice; isolated launches with the L2
flushed.  Usage: python tools/bench_band.py [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgpt_b200 import _lib  # noqa: E402

# (label, M, N, K, epilogue)
E_BF16, E_GELU, E_RESID_BF16 = 0, 1, 3
SHAPES = [("125m c_fc", 32768, 3072, 768, E_GELU), ("1.3b c_fc", 16384, 8192, 2048, E_GELU),
          ("1.3b c_proj", 16384, 2048, 8192, E_RESID_BF16),
          ("5.8b out_proj", 9600, 4096, 4096, E_RESID_BF16), ("5.8b c_proj", 9600, 4096, 16384, E_RESID_BF16),
          ("5.8b qkv", 9600, 12288, 4096, E_BF16), ("5.8b c_fc", 9600, 16384, 4096, E_GELU)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    L, lib = _lib, _lib.lib()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    bf = torch.bfloat16
    st = L.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.iters

    for label, M, N, K, epi in SHAPES:
        x = torch.randn(M, K, generator=g, device=dev).to(bf)
        w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(bf)
        bias = torch.randn(N, generator=g, device=dev) * 0.1
        out = torch.zeros(M, N, device=dev, dtype=bf)
        res = {"gemm": label, "M": M, "N": N, "K": K, "epilogue": epi}
        for setting in ("0", "4", "8", "16", None):
            if setting is None:
                os.environ.pop("SGPT_GEMM_BAND", None)
            else:
                os.environ["SGPT_GEMM_BAND"] = setting
            t = timeit(lambda: L.check(lib.sgpt_linear(x.data_ptr(), K, w.data_ptr(), K, bias.data_ptr(), out.data_ptr(), N,
                                                       out.data_ptr() if epi == E_RESID_BF16 else None, M, N, K, epi, st)))
            res["auto" if setting is None else "band" + setting] = {"us": round(1e3 * t, 1),
                                                                "tflops": round(2.0 * M * N * K / t / 1e9, 1)}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
