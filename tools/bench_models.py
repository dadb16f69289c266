#!/usr/bin/env python
"""Throughput of the encoder on the other BASELINE.json configs (random-init bf16 weights generated on the GPU):
SGPT-1.3B (batch 64 x 256), SGPT-5.8B / GPT-J (batch 32 x 300) and sgpt-bloom-7b1 (batch 32 x 300).
Not a bench line (bench.py measures configs[1]); output goes to profiles/ as supporting evidence.

    python tools/bench_models.py [--models sgpt-1.3b,sgpt-5.8b,sgpt-bloom-7b1] [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sgpt_b200 import Encoder, preset  # noqa: E402

SHAPES = {"sgpt-125m": (256, 128), "sgpt-1.3b": (64, 256), "sgpt-2.7b": (64, 256), "sgpt-5.8b": (32, 300),
          "sgpt-bloom-7b1": (32, 300)}


def rand_weights(cfg, dev):
    g = torch.Generator(device=dev).manual_seed(0)

    def w(*shape, sd=0.02, mean=0.0, dtype=torch.bfloat16):
        return (torch.randn(*shape, generator=g, device=dev) * sd + mean).to(dtype)

    d, ff, L = cfg.d_model, cfg.d_ff, cfg.n_layer
    sd = {}
    f32 = torch.float32
    if cfg.arch == "gpt_neo":
        sd["wte.weight"], sd["wpe.weight"] = w(cfg.vocab, d), w(cfg.max_pos, d, sd=0.01)
        for i in range(L):
            p = f"h.{i}."
            sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = w(d, sd=0.1, mean=1, dtype=f32), w(d, sd=0.05, dtype=f32)
            for n in ("q_proj", "k_proj", "v_proj"):
                sd[p + f"attn.attention.{n}.weight"] = w(d, d, sd=0.02 * (768 / d) ** 0.5)
            sd[p + "attn.attention.out_proj.weight"], sd[p + "attn.attention.out_proj.bias"] = w(d, d), w(d, dtype=f32)
            sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = w(d, sd=0.1, mean=1, dtype=f32), w(d, sd=0.05, dtype=f32)
            sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"] = w(ff, d), w(ff, dtype=f32)
            sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"] = w(d, ff, sd=0.01), w(d, dtype=f32)
    elif cfg.arch == "gptj":
        sd["wte.weight"] = w(cfg.vocab, d)
        for i in range(L):
            p = f"h.{i}."
            sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = w(d, sd=0.1, mean=1, dtype=f32), w(d, sd=0.05, dtype=f32)
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                sd[p + f"attn.{n}.weight"] = w(d, d, sd=0.01)
            sd[p + "mlp.fc_in.weight"], sd[p + "mlp.fc_in.bias"] = w(ff, d, sd=0.01), w(ff, dtype=f32)
            sd[p + "mlp.fc_out.weight"], sd[p + "mlp.fc_out.bias"] = w(d, ff, sd=0.005), w(d, dtype=f32)
    else:
        sd["word_embeddings.weight"] = w(cfg.vocab, d)
        sd["word_embeddings_layernorm.weight"], sd["word_embeddings_layernorm.bias"] = w(d, sd=0.1, mean=1, dtype=f32), w(d, sd=0.05, dtype=f32)
        for i in range(L):
            p = f"h.{i}."
            sd[p + "input_layernorm.weight"], sd[p + "input_layernorm.bias"] = w(d, sd=0.1, mean=1, dtype=f32), w(d, sd=0.05, dtype=f32)
            sd[p + "self_attention.query_key_value.weight"], sd[p + "self_attention.query_key_value.bias"] = w(3 * d, d, sd=0.01), w(3 * d, dtype=f32)
            sd[p + "self_attention.dense.weight"], sd[p + "self_attention.dense.bias"] = w(d, d, sd=0.01), w(d, dtype=f32)
            sd[p + "post_attention_layernorm.weight"], sd[p + "post_attention_layernorm.bias"] = w(d, sd=0.1, mean=1, dtype=f32), w(d, sd=0.05, dtype=f32)
            sd[p + "mlp.dense_h_to_4h.weight"], sd[p + "mlp.dense_h_to_4h.bias"] = w(ff, d, sd=0.01), w(ff, dtype=f32)
            sd[p + "mlp.dense_4h_to_h.weight"], sd[p + "mlp.dense_4h_to_h.bias"] = w(d, ff, sd=0.005), w(d, dtype=f32)
    sd["ln_f.weight"], sd["ln_f.bias"] = w(d, sd=0.1, mean=1, dtype=f32), w(d, sd=0.05, dtype=f32)
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="sgpt-1.3b,sgpt-5.8b,sgpt-bloom-7b1")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="override the batch size (sequences) of every model")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-launch event timing")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = 1431.3
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["bf16_tflops_sustained"]
    for name in args.models.split(","):
        cfg = preset(name)
        B, S = SHAPES[name]
        if args.batch:
            B = args.batch
        sd = rand_weights(cfg, dev)
        enc = Encoder(cfg, sd, device=dev, max_tokens=B * S, max_batch=B)
        del sd
        g = torch.Generator().manual_seed(1)
        ids = torch.randint(0, cfg.vocab, (B, S), generator=g).numpy()
        mask = np.ones((B, S), dtype=np.int8)
        for _ in range(2):
            out = enc.encode_tokens(ids, mask)
        torch.cuda.synchronize()
        import ctypes

        from sgpt_b200 import _lib
        lib = _lib.lib()
        ms_cat, n_cat = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
        lib.sgpt_profile_read(None, None, None)
        lib.sgpt_profile_enable(0 if args.no_profile else 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = enc.encode_tokens(ids, mask)
        e1.record()
        torch.cuda.synchronize()
        lib.sgpt_profile_enable(0)
        lib.sgpt_profile_read(ms_cat, n_cat, None)
        ms = e0.elapsed_time(e1) / args.steps
        cats = ["embed", "layernorm", "linear_gemm", "attention", "pool", "similarity_gemm", "topk", "misc"]
        per_kernel = {c: round(ms_cat[i] / args.steps, 3) for i, c in enumerate(cats) if ms_cat[i] > 0}
        L, d, ff = cfg.n_layer, cfg.d_model, cfg.d_ff
        flops = B * (S * 2 * L * (4 * d * d + 2 * d * ff) + L * 2 * S * (S + 1) * d)
        line = {"model": name, "arch": cfg.arch, "batch": B, "seq_len": S, "ms_per_batch": ms,
                "embeddings_per_s": B / (ms / 1e3), "model_tflops": flops / (ms / 1e3) / 1e12,
                "frac_of_sustained_bf16_peak": flops / (ms / 1e3) / 1e12 / peak, "finite": bool(torch.isfinite(out).all()), "kernel_ms_per_batch": per_kernel}
        print(json.dumps(line), flush=True)
        enc.close()
        del enc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
