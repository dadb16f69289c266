#!/usr/bin/env python
"""Calibration only (not a product path): what does the vendor library (cuBLASLt through torch) reach on the four linear
shapes of SGPT-125M at batch 256 x 128, timed the same way as `native_tests linperf` (CUDA events, back to back)?

    python tools/cublas_shapes.py
"""
import json

import torch

dev = torch.device("cuda:0")
M = 32768
SHAPES = [("qkv", 2304, 768, "bf16"), ("out_proj+resid", 768, 768, "resid"), ("c_fc+gelu", 3072, 768, "gelu"),
          ("c_proj+resid", 768, 3072, "resid"), ("big", 8192, 2048, "bf16")]
for name, N, K, epi in SHAPES:
    m = 16384 if name == "big" else M
    x = torch.randn(m, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
    b = torch.randn(N, device=dev, dtype=torch.bfloat16)
    resid = torch.zeros(m, N, device=dev, dtype=torch.float32)

    def run():
        y = torch.nn.functional.linear(x, w, b)
        if epi == "gelu":
            y = torch.nn.functional.gelu(y, approximate="tanh")
        elif epi == "resid":
            resid.add_(y)
        return y

    def run_mm_only():
        return torch.nn.functional.linear(x, w, b)

    out = {}
    for label, fn in (("with_epilogue_ops", run), ("linear_only", run_mm_only)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[label] = {"ms": round(ms, 4), "tflops": round(2.0 * m * N * K / (ms * 1e9), 1)}
    print(json.dumps({"shape": name, "M": m, "N": N, "K": K, **out}), flush=True)
