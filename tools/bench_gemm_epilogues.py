#!/usr/bin/env python
"""Per-shape timing of the encoder GEMMs with each epilogue (C ABI called directly): the stand-alone-LayerNorm variants
(sgpt_linear bf16 / gelu / reduce-add residual) against the LayerNorm-folded ones (sgpt_linear_lnfold,
sgpt_linear_resid_ln).  Usage: python tools/bench_gemm_epilogues.py [--model 125m|1.3b|6b] [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgpt_b200 import _lib  # noqa: E402

SHAPES = {"125m": (256 * 128, 768, 3072), "1.3b": (64 * 256, 2048, 8192), "6b": (32 * 300, 4096, 16384)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="125m")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    L, lib = _lib, _lib.lib()
    M, d, ff = SHAPES[a.model]
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    bf = torch.bfloat16
    x_d = torch.randn(M, d, generator=g, device=dev).to(bf)
    x_ff = torch.randn(M, ff, generator=g, device=dev).to(bf)
    resid = torch.randn(M, d, generator=g, device=dev)
    P = (d + 127) // 128
    stats = torch.zeros(M, P, 2, device=dev)
    xb = torch.empty(M, d, dtype=bf, device=dev)
    L.check(lib.sgpt_resid_stats(resid.data_ptr(), xb.data_ptr(), stats.data_ptr(), M, d, L.current_stream()))
    st = L.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.iters):
            flush.zero_()  # evict L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.iters

    for name, N, K, xin in (("qkv", 3 * d, d, x_d), ("fc", ff, d, x_d), ("out_proj", d, d, x_d), ("c_proj", d, ff, x_ff)):
        w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(bf)
        bias = torch.randn(N, generator=g, device=dev) * 0.1
        cs = torch.randn(N, generator=g, device=dev)
        flops = 2.0 * M * N * K
        res = {"gemm": name, "M": M, "N": N, "K": K}
        if N != d:
            out = torch.empty(M, N, dtype=bf, device=dev)
            for epi, code in (("bf16", 0), ("gelu_bf16", 1)):
                t = timeit(lambda: L.check(lib.sgpt_linear(xin.data_ptr(), K, w.data_ptr(), K, bias.data_ptr(), out.data_ptr(), N,
                                                           None, M, N, K, code, st)))
                res[epi] = {"us": round(1e3 * t, 1), "tflops": round(flops / t / 1e9, 1)}
            for gelu in (0, 1):
                t = timeit(lambda: L.check(lib.sgpt_linear_lnfold(xin.data_ptr(), K, w.data_ptr(), K, bias.data_ptr(), cs.data_ptr(),
                                                                  stats.data_ptr(), P, 1e-5, out.data_ptr(), N, M, N, K, gelu, st)))
                res["lnfold" + ("_gelu" if gelu else "")] = {"us": round(1e3 * t, 1), "tflops": round(flops / t / 1e9, 1)}
        else:
            r2 = resid.clone()
            t = timeit(lambda: L.check(lib.sgpt_linear(xin.data_ptr(), K, w.data_ptr(), K, bias.data_ptr(), r2.data_ptr(), N,
                                                       r2.data_ptr(), M, N, K, 2, st)))
            res["resid_reduce_add"] = {"us": round(1e3 * t, 1), "tflops": round(flops / t / 1e9, 1)}
            t = timeit(lambda: L.check(lib.sgpt_linear_resid_ln(xin.data_ptr(), K, w.data_ptr(), K, bias.data_ptr(), r2.data_ptr(),
                                                                xb.data_ptr(), stats.data_ptr(), M, N, K, st)))
            res["resid_ln"] = {"us": round(1e3 * t, 1), "tflops": round(flops / t / 1e9, 1)}
        print(json.dumps(res), flush=True)
    # the stand-alone LayerNorm this replaces
    y = torch.empty(M, d, dtype=bf, device=dev)
    gam = torch.ones(d, device=dev)
    t = timeit(lambda: L.check(lib.sgpt_layernorm(resid.data_ptr(), gam.data_ptr(), gam.data_ptr(), y.data_ptr(), M, d, 1e-5, st)))
    print(json.dumps({"kernel": "layernorm (stand-alone pass)", "us": round(1e3 * t, 1)}), flush=True)


if __name__ == "__main__":
    main()
