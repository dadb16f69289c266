"""Shared test helpers: a tiny deterministic tokenizer (no pretrained files are reachable offline) and builders."""
import re
import zlib

import numpy as np
import torch


class ToyTokenizer:
    """Whitespace/bracket tokenizer with the HF methods the reference calls (tokenize, convert_tokens_to_ids, encode)."""

    SPECIAL = {"[": 58, "]": 60, "{": 90, "}": 92, "[SOS]": 997, "{SOS}": 998}  # GPT-2 BPE ids of the brackets

    def __init__(self, vocab=1000, pad_token_id=999):
        self.vocab = vocab
        self.pad_token_id = pad_token_id
        self.eos_token = "<|endoftext|>"
        self.pad_token = None

    def tokenize(self, text):
        return re.findall(r"\[SOS\]|\{SOS\}|[\[\]{}]|[^\s\[\]{}]+", text)

    def convert_tokens_to_ids(self, tokens):
        return [self.SPECIAL[t] if t in self.SPECIAL else 100 + zlib.crc32(t.encode()) % (self.vocab - 200) for t in tokens]

    def encode(self, text, add_special_tokens=False):
        return self.convert_tokens_to_ids(self.tokenize(text))


def ragged_batch(B, S, vocab, seed, pad_id=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, S + 1, (B,), generator=g)
    lens[0] = 1
    if B > 1:
        lens[1] = S
    ids = torch.randint(0, vocab, (B, S), generator=g)
    mask = (torch.arange(S).unsqueeze(0) < lens.unsqueeze(1)).long()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, pad_id))
    return ids, mask


def planted_corpus(n, D, nq, seed):
    """Random corpus with ~1% planted near-duplicates of the queries so the top of the ranking is meaningful."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(nq, D, generator=g)
    c = torch.randn(n, D, generator=g)
    idx = torch.arange(0, n, 97)
    c[idx] = q[idx // 97 % nq] + 0.5 * c[idx]
    return q, c


def min_row_cosine(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return torch.nn.functional.cosine_similarity(a, b, dim=1).min().item()


class LazyF32Weights(dict):
    """bf16/fp32 CPU tensors that are widened to fp32 only while an oracle expression uses them: the functional oracles
    index the state dict once per use, so a 6-7 B-parameter model never exists in fp32 as a whole (12-14 GB of bf16 on
    the host instead of 24-28 GB of fp32)."""

    def __getitem__(self, key):
        return dict.__getitem__(self, key).float()


def full_size_weights(arch, spec, seed, device="cuda:0"):
    """Random weights of a FULL-SIZE model (BASELINE.json configs[2..4]) drawn on the GPU (torch.randn on the CPU needs
    a minute per billion parameters), linear weights in bf16 — exactly the values the CUDA path stores — LayerNorm
    parameters and biases in fp32.  Key names follow the HF state_dict of the architecture (oracle/*.py:init_weights
    use the same construction at small sizes).  Returns (device dict for sgpt_b200.Encoder, LazyF32Weights for the CPU
    oracle holding the same values)."""
    import math

    g = torch.Generator(device=device).manual_seed(seed)
    bf, f32 = torch.bfloat16, torch.float32

    def rnd(*shape, sd=0.02, mean=0.0, dtype=bf):
        return (torch.randn(*shape, generator=g, device=device) * sd + mean).to(dtype)

    d, ff, L = spec.d_model, spec.d_ff, spec.n_layer
    hd = d // spec.n_head
    w = {}
    if arch == "gpt_neo":
        qk_sd = math.sqrt(2.5 / (d * math.sqrt(hd)))  # un-scaled q.k logits with a spread of ~2.5 (neither flat nor one-hot)
        w["wte.weight"], w["wpe.weight"] = rnd(spec.vocab, d), rnd(spec.max_pos, d, sd=0.01)
        for i in range(L):
            p = f"h.{i}."
            w[p + "ln_1.weight"], w[p + "ln_1.bias"] = rnd(d, sd=0.1, mean=1.0, dtype=f32), rnd(d, sd=0.05, dtype=f32)
            w[p + "attn.attention.q_proj.weight"], w[p + "attn.attention.k_proj.weight"] = rnd(d, d, sd=qk_sd), rnd(d, d, sd=qk_sd)
            w[p + "attn.attention.v_proj.weight"] = rnd(d, d)
            w[p + "attn.attention.out_proj.weight"], w[p + "attn.attention.out_proj.bias"] = rnd(d, d), rnd(d, dtype=f32)
            w[p + "ln_2.weight"], w[p + "ln_2.bias"] = rnd(d, sd=0.1, mean=1.0, dtype=f32), rnd(d, sd=0.05, dtype=f32)
            w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"] = rnd(ff, d), rnd(ff, dtype=f32)
            w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"] = rnd(d, ff), rnd(d, dtype=f32)
    elif arch == "gptj":
        qk_sd = math.sqrt(2.5 / (d * math.sqrt(hd))) * math.sqrt(math.sqrt(hd))  # logits / sqrt(hd) with spread ~2.5
        w["wte.weight"] = rnd(spec.vocab, d)
        for i in range(L):
            p = f"h.{i}."
            w[p + "ln_1.weight"], w[p + "ln_1.bias"] = rnd(d, sd=0.1, mean=1.0, dtype=f32), rnd(d, sd=0.05, dtype=f32)
            w[p + "attn.q_proj.weight"], w[p + "attn.k_proj.weight"] = rnd(d, d, sd=qk_sd), rnd(d, d, sd=qk_sd)
            w[p + "attn.v_proj.weight"], w[p + "attn.out_proj.weight"] = rnd(d, d), rnd(d, d)
            w[p + "mlp.fc_in.weight"], w[p + "mlp.fc_in.bias"] = rnd(ff, d), rnd(ff, dtype=f32)
            w[p + "mlp.fc_out.weight"], w[p + "mlp.fc_out.bias"] = rnd(d, ff), rnd(d, dtype=f32)
    elif arch == "bloom":
        qk_sd = math.sqrt(2.5 / (d * math.sqrt(hd))) * math.sqrt(math.sqrt(hd))
        w["word_embeddings.weight"] = rnd(spec.vocab, d)
        w["word_embeddings_layernorm.weight"] = rnd(d, sd=0.1, mean=1.0, dtype=f32)
        w["word_embeddings_layernorm.bias"] = rnd(d, sd=0.05, dtype=f32)
        for i in range(L):
            p = f"h.{i}."
            w[p + "input_layernorm.weight"], w[p + "input_layernorm.bias"] = rnd(d, sd=0.1, mean=1.0, dtype=f32), rnd(d, sd=0.05, dtype=f32)
            qkv = rnd(3 * d, d, sd=qk_sd).view(spec.n_head, 3, hd, d)  # per head [q | k | v] rows (HF:bloom:211-215)
            qkv[:, 2] = rnd(spec.n_head, hd, d)                         # v rows at the ordinary 0.02
            w[p + "self_attention.query_key_value.weight"] = qkv.view(3 * d, d)
            w[p + "self_attention.query_key_value.bias"] = rnd(3 * d, dtype=f32)
            w[p + "self_attention.dense.weight"], w[p + "self_attention.dense.bias"] = rnd(d, d), rnd(d, dtype=f32)
            w[p + "post_attention_layernorm.weight"] = rnd(d, sd=0.1, mean=1.0, dtype=f32)
            w[p + "post_attention_layernorm.bias"] = rnd(d, sd=0.05, dtype=f32)
            w[p + "mlp.dense_h_to_4h.weight"], w[p + "mlp.dense_h_to_4h.bias"] = rnd(ff, d), rnd(ff, dtype=f32)
            w[p + "mlp.dense_4h_to_h.weight"], w[p + "mlp.dense_4h_to_h.bias"] = rnd(d, ff), rnd(d, dtype=f32)
    else:
        raise ValueError(arch)
    w["ln_f.weight"], w["ln_f.bias"] = rnd(d, sd=0.1, mean=1.0, dtype=f32), rnd(d, sd=0.05, dtype=f32)
    return w, LazyF32Weights({k: v.cpu() for k, v in w.items()})
