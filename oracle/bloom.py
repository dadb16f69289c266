"""Oracle: BLOOM forward pass (sgpt-bloom-7b1-msmarco family), restated functionally on CPU tensors (test infra only).

Follows HuggingFace ``transformers/models/bloom/modeling_bloom.py`` (installed 5.5.0; cited as ``HF:bloom:<lines>``).
Weights: flat dict keyed like ``BloomModel.state_dict()``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from .gpt_neo import layer_norm


@dataclass
class BloomSpec:
    n_layer: int = 30          # n_layer
    d_model: int = 4096        # hidden_size
    n_head: int = 32           # n_head
    vocab: int = 250880        # vocab_size
    ln_eps: float = 1e-5       # layer_norm_epsilon

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_head

    @property
    def d_ff(self) -> int:
        return 4 * self.d_model


def init_weights(spec: BloomSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, sd=0.02, mean=0.0):
        return (torch.randn(*shape, generator=g) * sd + mean).to(torch.bfloat16).float()

    d, ff = spec.d_model, spec.d_ff
    w = {"word_embeddings.weight": rnd(spec.vocab, d),
         "word_embeddings_layernorm.weight": rnd(d, sd=0.1, mean=1.0), "word_embeddings_layernorm.bias": rnd(d, sd=0.05)}
    for i in range(spec.n_layer):
        p = f"h.{i}."
        w[p + "input_layernorm.weight"], w[p + "input_layernorm.bias"] = rnd(d, sd=0.1, mean=1.0), rnd(d, sd=0.05)
        w[p + "self_attention.query_key_value.weight"] = rnd(3 * d, d, sd=math.sqrt(2.5 / (d * math.sqrt(spec.head_dim))))
        w[p + "self_attention.query_key_value.bias"] = rnd(3 * d)
        w[p + "self_attention.dense.weight"], w[p + "self_attention.dense.bias"] = rnd(d, d), rnd(d)
        w[p + "post_attention_layernorm.weight"] = rnd(d, sd=0.1, mean=1.0)
        w[p + "post_attention_layernorm.bias"] = rnd(d, sd=0.05)
        w[p + "mlp.dense_h_to_4h.weight"], w[p + "mlp.dense_h_to_4h.bias"] = rnd(ff, d), rnd(ff)
        w[p + "mlp.dense_4h_to_h.weight"], w[p + "mlp.dense_4h_to_h.bias"] = rnd(d, ff), rnd(d)
    w["ln_f.weight"], w["ln_f.bias"] = rnd(d, sd=0.1, mean=1.0), rnd(d, sd=0.05)
    return w


def alibi_slopes(n_head: int) -> torch.Tensor:
    """build_alibi_tensor slopes (HF:bloom:62-78)."""
    cp2 = 2 ** math.floor(math.log2(n_head))
    base = torch.tensor(2 ** (-(2 ** -(math.log2(cp2) - 3))), dtype=torch.float32)
    slopes = torch.pow(base, torch.arange(1, 1 + cp2, dtype=torch.int32))
    if cp2 != n_head:
        extra_base = torch.tensor(2 ** (-(2 ** -(math.log2(2 * cp2) - 3))), dtype=torch.float32)
        n_rem = min(cp2, n_head - cp2)
        slopes = torch.cat([slopes, torch.pow(extra_base, torch.arange(1, 1 + 2 * n_rem, 2, dtype=torch.int32))], dim=0)
    return slopes


def bloom_gelu(x: torch.Tensor) -> torch.Tensor:
    """bloom_gelu_forward (HF:bloom:108-116)."""
    return x * 0.5 * (1.0 + torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x)))


def forward(spec: BloomSpec, w: Dict[str, torch.Tensor], input_ids: torch.Tensor,
            attention_mask: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """BloomModel.forward(output_hidden_states=True) (HF:bloom:440-560)."""
    B, S = input_ids.shape
    H, hd = spec.n_head, spec.head_dim
    mask = attention_mask if attention_mask is not None else torch.ones(B, S, dtype=torch.long)
    h = layer_norm(w["word_embeddings.weight"][input_ids], w["word_embeddings_layernorm.weight"],
                   w["word_embeddings_layernorm.bias"], spec.ln_eps)  # HF:bloom:496
    # alibi[b,h,0,j] = slope_h * ((cumsum(mask)-1) * mask)[b,j]   (HF:bloom:84-86)
    arange = ((mask.cumsum(dim=-1) - 1) * mask)[:, None, :].float()
    alibi = (alibi_slopes(H)[None, :, None] * arange)[:, :, None, :]  # [B,H,1,S]
    i = torch.arange(S).unsqueeze(1)
    j = torch.arange(S).unsqueeze(0)
    neg = torch.finfo(torch.float32).min
    bias = torch.where(j <= i, 0.0, neg)[None, None] + (1.0 - mask[:, None, None, :].float()) * neg
    hidden = []
    for li in range(spec.n_layer):
        hidden.append(h)
        p = f"h.{li}."
        x = layer_norm(h, w[p + "input_layernorm.weight"], w[p + "input_layernorm.bias"], spec.ln_eps)
        fused = x @ w[p + "self_attention.query_key_value.weight"].T + w[p + "self_attention.query_key_value.bias"]
        fused = fused.view(B, S, H, 3, hd)  # HF:bloom:211-215: per head [q | k | v]
        q, k, v = (fused[..., t, :].transpose(1, 2) for t in range(3))
        scores = alibi + torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)  # baddbmm, HF:bloom:267-273
        scores = scores + bias
        a = torch.matmul(torch.softmax(scores, dim=-1, dtype=torch.float32), v)  # HF:bloom:280-289
        a = a.transpose(1, 2).reshape(B, S, H * hd)
        h = h + (a @ w[p + "self_attention.dense.weight"].T + w[p + "self_attention.dense.bias"])  # HF:bloom:304-306
        x = layer_norm(h, w[p + "post_attention_layernorm.weight"], w[p + "post_attention_layernorm.bias"], spec.ln_eps)
        m = bloom_gelu(x @ w[p + "mlp.dense_h_to_4h.weight"].T + w[p + "mlp.dense_h_to_4h.bias"])
        h = h + (m @ w[p + "mlp.dense_4h_to_h.weight"].T + w[p + "mlp.dense_4h_to_h.bias"])  # HF:bloom:330-339
    h = layer_norm(h, w["ln_f.weight"], w["ln_f.bias"], spec.ln_eps)
    hidden.append(h)
    return hidden
