"""Oracle: GPT-J forward pass (SGPT-5.8B family), restated functionally on CPU tensors (test infrastructure only).

Follows HuggingFace ``transformers/models/gptj/modeling_gptj.py`` (installed 5.5.0; cited as ``HF:gptj:<lines>``).
Weights: flat dict keyed like ``GPTJModel.state_dict()``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from .gpt_neo import gelu_new, layer_norm


@dataclass
class GPTJSpec:
    n_layer: int = 28          # n_layer
    d_model: int = 4096        # n_embd
    n_head: int = 16           # n_head
    d_ff: int = 16384          # n_inner (4 * n_embd)
    vocab: int = 50400         # vocab_size
    max_pos: int = 2048        # n_positions
    rotary_dim: int = 64       # rotary_dim
    ln_eps: float = 1e-5       # layer_norm_epsilon

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_head


def init_weights(spec: GPTJSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, sd=0.02, mean=0.0):
        return (torch.randn(*shape, generator=g) * sd + mean).to(torch.bfloat16).float()

    d, ff = spec.d_model, spec.d_ff
    w = {"wte.weight": rnd(spec.vocab, d)}
    qk_sd = math.sqrt(2.5 / (d * math.sqrt(spec.head_dim)))  # logits / sqrt(hd) with a spread of ~2.5
    for i in range(spec.n_layer):
        p = f"h.{i}."
        w[p + "ln_1.weight"], w[p + "ln_1.bias"] = rnd(d, sd=0.1, mean=1.0), rnd(d, sd=0.05)
        w[p + "attn.q_proj.weight"], w[p + "attn.k_proj.weight"] = rnd(d, d, sd=qk_sd), rnd(d, d, sd=qk_sd)
        w[p + "attn.v_proj.weight"], w[p + "attn.out_proj.weight"] = rnd(d, d), rnd(d, d)
        w[p + "mlp.fc_in.weight"], w[p + "mlp.fc_in.bias"] = rnd(ff, d), rnd(ff)
        w[p + "mlp.fc_out.weight"], w[p + "mlp.fc_out.bias"] = rnd(d, ff), rnd(d)
    w["ln_f.weight"], w["ln_f.bias"] = rnd(d, sd=0.1, mean=1.0), rnd(d, sd=0.05)
    return w


def rotary_tables(max_pos: int, rotary_dim: int):
    """create_sinusoidal_positions (HF:gptj:45-48): theta_{p,i} = p * 10000^(-2i/rotary_dim); returns (sin, cos) [P, rd/2]."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, rotary_dim, 2, dtype=torch.int64) / rotary_dim))
    ang = torch.einsum("i,j->ij", torch.arange(max_pos, dtype=torch.int64).float(), inv_freq).float()
    return torch.sin(ang), torch.cos(ang)


def apply_rotary(x: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor) -> torch.Tensor:
    """apply_rotary_pos_emb + rotate_every_two (HF:gptj:55-67) on x [B,S,H,rd]; sin/cos [B,S,rd/2].
    Pairs are INTERLEAVED: (x[2i], x[2i+1]) -> (x[2i] c - x[2i+1] s, x[2i+1] c + x[2i] s)."""
    s = torch.repeat_interleave(sin[:, :, None, :], 2, 3)
    c = torch.repeat_interleave(cos[:, :, None, :], 2, 3)
    x1, x2 = x[..., ::2], x[..., 1::2]
    rot = torch.stack((-x2, x1), dim=-1).flatten(-2)
    return x * c + rot * s


def forward(spec: GPTJSpec, w: Dict[str, torch.Tensor], input_ids: torch.Tensor,
            attention_mask: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """GPTJModel.forward(output_hidden_states=True) (HF:gptj:440-560): hidden_states list of n_layer+1 tensors."""
    B, S = input_ids.shape
    H, hd, rd = spec.n_head, spec.head_dim, spec.rotary_dim
    pos = torch.arange(S).unsqueeze(0).expand(B, S)  # position_ids default: arange (HF:gptj:500-503)
    sin_t, cos_t = rotary_tables(spec.max_pos, rd)
    sin, cos = sin_t[pos], cos_t[pos]
    h = w["wte.weight"][input_ids]  # no position embedding (HF:gptj:494)
    i = torch.arange(S).unsqueeze(1)
    j = torch.arange(S).unsqueeze(0)
    neg = torch.finfo(torch.float32).min
    bias = torch.where(j <= i, 0.0, neg)[None, None]
    if attention_mask is not None:
        bias = bias + (1.0 - attention_mask[:, None, None, :].float()) * neg
    hidden = []
    for li in range(spec.n_layer):
        hidden.append(h)
        p = f"h.{li}."
        x = layer_norm(h, w[p + "ln_1.weight"], w[p + "ln_1.bias"], spec.ln_eps)  # HF:gptj:400
        q = (x @ w[p + "attn.q_proj.weight"].T).view(B, S, H, hd)  # no biases (HF:gptj:98-101)
        k = (x @ w[p + "attn.k_proj.weight"].T).view(B, S, H, hd)
        v = (x @ w[p + "attn.v_proj.weight"].T).view(B, S, H, hd)
        q = torch.cat([apply_rotary(q[..., :rd], sin, cos), q[..., rd:]], dim=-1)  # HF:gptj:196-207
        k = torch.cat([apply_rotary(k[..., :rd], sin, cos), k[..., rd:]], dim=-1)
        q, k, v = q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)
        scores = torch.matmul(q.float(), k.float().transpose(-1, -2)) / math.sqrt(hd)  # HF:gptj:145-148
        scores = scores + bias
        a = torch.matmul(torch.softmax(scores, dim=-1), v)  # HF:gptj:153-157
        a = a.permute(0, 2, 1, 3).reshape(B, S, H * hd) @ w[p + "attn.out_proj.weight"].T
        m = gelu_new(x @ w[p + "mlp.fc_in.weight"].T + w[p + "mlp.fc_in.bias"])  # MLP reads the SAME ln_1 output
        m = m @ w[p + "mlp.fc_out.weight"].T + w[p + "mlp.fc_out.bias"]
        h = a + m + h  # parallel residual (HF:gptj:411)
    h = layer_norm(h, w["ln_f.weight"], w["ln_f.bias"], spec.ln_eps)
    hidden.append(h)
    return hidden
