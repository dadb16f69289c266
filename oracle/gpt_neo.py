"""Oracle: GPT-Neo forward pass, restated functionally on CPU tensors (test infrastructure only).

Follows HuggingFace ``transformers/models/gpt_neo/modeling_gpt_neo.py`` (installed 5.5.0; the reference pins
>=4.6,<5 — the eager-attention arithmetic is unchanged across those versions).  Every function cites the lines it
restates as ``HF:gpt_neo:<lines>``.  Weights are passed as a flat ``dict[str, Tensor]`` keyed exactly like
``GPTNeoModel.state_dict()`` so the same dict can be loaded into the HF module (see tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch


@dataclass
class NeoSpec:
    """Hyper-parameters of a GPT-Neo encoder (HF GPTNeoConfig field names in comments)."""

    n_layer: int = 12              # num_layers
    d_model: int = 768             # hidden_size
    n_head: int = 12               # num_heads
    d_ff: int = 3072               # intermediate_size (4 * hidden_size)
    vocab: int = 50257             # vocab_size
    max_pos: int = 2048            # max_position_embeddings
    window: int = 256              # window_size
    ln_eps: float = 1e-5           # layer_norm_epsilon
    attention_layers: List[str] = field(default_factory=list)  # "global" / "local" per layer

    def __post_init__(self):
        if not self.attention_layers:
            # GPTNeoConfig default attention_types=[[["global","local"], n/2]] -> alternating, global first
            self.attention_layers = ["global" if i % 2 == 0 else "local" for i in range(self.n_layer)]

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_head


SGPT_125M = dict(n_layer=12, d_model=768, n_head=12, d_ff=3072)
SGPT_1_3B = dict(n_layer=24, d_model=2048, n_head=16, d_ff=8192)


def init_weights(spec: NeoSpec, seed: int = 0, dtype=torch.float32, round_bf16: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (no hub access offline).  sigma=0.02 like HF's initializer_range, but with
    non-trivial LayerNorm scales/offsets and biases so every term of the forward pass is exercised.  Values are
    rounded to bf16-representable numbers (the CUDA path stores weights in bf16) and returned in `dtype`."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, sd=0.02, mean=0.0):
        t = torch.randn(*shape, generator=g, dtype=torch.float32) * sd + mean
        if round_bf16:
            t = t.to(torch.bfloat16).to(torch.float32)
        return t.to(dtype)

    d, ff = spec.d_model, spec.d_ff
    w: Dict[str, torch.Tensor] = {}
    w["wte.weight"] = rnd(spec.vocab, d)
    w["wpe.weight"] = rnd(spec.max_pos, d, sd=0.01)
    for i in range(spec.n_layer):
        p = f"h.{i}."
        w[p + "ln_1.weight"] = rnd(d, sd=0.1, mean=1.0)
        w[p + "ln_1.bias"] = rnd(d, sd=0.05)
        for n in ("q_proj", "k_proj", "v_proj"):
            # sd 0.02: with LN-scale inputs the (unscaled!) GPT-Neo logits q.k get a spread of ~sqrt(hd)*d*sd^2 ~ 2.5,
            # i.e. a softmax that is neither flat nor one-hot
            w[p + f"attn.attention.{n}.weight"] = rnd(d, d)
        w[p + "attn.attention.out_proj.weight"] = rnd(d, d)
        w[p + "attn.attention.out_proj.bias"] = rnd(d, sd=0.02)
        w[p + "ln_2.weight"] = rnd(d, sd=0.1, mean=1.0)
        w[p + "ln_2.bias"] = rnd(d, sd=0.05)
        w[p + "mlp.c_fc.weight"] = rnd(ff, d)
        w[p + "mlp.c_fc.bias"] = rnd(ff, sd=0.02)
        w[p + "mlp.c_proj.weight"] = rnd(d, ff)
        w[p + "mlp.c_proj.bias"] = rnd(d, sd=0.02)
    w["ln_f.weight"] = rnd(d, sd=0.1, mean=1.0)
    w["ln_f.bias"] = rnd(d, sd=0.05)
    return w


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
    """nn.LayerNorm over the last dim (HF:gpt_neo:332,345,492): biased variance, eps inside the sqrt."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """NewGELUActivation (HF activations.py): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attention_mask: Optional[torch.Tensor], local: bool,
              window: int) -> torch.Tensor:
    """GPTNeoSelfAttention._attn (HF:gpt_neo:105-130).  q,k,v: [B,H,S,hd].

    * scores = q @ k^T in fp32, **no 1/sqrt(hd) scaling** (:110)
    * causal mask, and for "local" layers key j is visible to query i iff i - window < j <= i (:63-66, :114-118);
      masked scores are set to finfo.min
    * + additive padding mask (0 / finfo.min over key positions) (:120-122)
    * softmax in fp32 (:124), then P @ V
    """
    B, H, S, hd = q.shape
    scores = torch.matmul(q.float(), k.float().transpose(-1, -2))
    i = torch.arange(S).unsqueeze(1)
    j = torch.arange(S).unsqueeze(0)
    visible = j <= i
    if local:
        visible = visible & (j > i - window)
    neg = torch.finfo(scores.dtype).min
    scores = torch.where(visible, scores, torch.tensor(neg, dtype=scores.dtype))
    if attention_mask is not None:
        pad = (1.0 - attention_mask[:, None, None, :].to(scores.dtype)) * neg
        scores = scores + pad  # may overflow to -inf for doubly masked entries, exactly as in HF
    p = torch.softmax(scores, dim=-1).to(v.dtype)
    return torch.matmul(p, v)


def forward(spec: NeoSpec, w: Dict[str, torch.Tensor], input_ids: torch.Tensor,
            attention_mask: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """GPTNeoModel.forward(..., output_hidden_states=True) (HF:gpt_neo:400-500).

    Returns the tuple ``hidden_states`` as a list of n_layer+1 tensors [B,S,d]: entry i < n_layer is the input of
    block i, entry n_layer is ln_f(output of the last block) (== last_hidden_state, HF:gpt_neo:492-497).
    """
    B, S = input_ids.shape
    H, hd = spec.n_head, spec.head_dim
    pos = torch.arange(S)
    h = w["wte.weight"][input_ids] + w["wpe.weight"][pos].unsqueeze(0)  # :462-463
    hidden = []
    for li in range(spec.n_layer):
        hidden.append(h)
        p = f"h.{li}."
        x = layer_norm(h, w[p + "ln_1.weight"], w[p + "ln_1.bias"], spec.ln_eps)  # :332

        def split(t):  # _split_heads :90-96
            return t.view(B, S, H, hd).permute(0, 2, 1, 3)

        q = split(x @ w[p + "attn.attention.q_proj.weight"].T)  # no bias :84-86
        k = split(x @ w[p + "attn.attention.k_proj.weight"].T)
        v = split(x @ w[p + "attn.attention.v_proj.weight"].T)
        a = attention(q, k, v, attention_mask, spec.attention_layers[li] == "local", spec.window)
        a = a.permute(0, 2, 1, 3).reshape(B, S, H * hd)  # _merge_heads :98-103
        a = a @ w[p + "attn.attention.out_proj.weight"].T + w[p + "attn.attention.out_proj.bias"]  # :153
        h = a + h  # :342
        x = layer_norm(h, w[p + "ln_2.weight"], w[p + "ln_2.bias"], spec.ln_eps)  # :345
        m = gelu_new(x @ w[p + "mlp.c_fc.weight"].T + w[p + "mlp.c_fc.bias"])  # :304-305
        m = m @ w[p + "mlp.c_proj.weight"].T + w[p + "mlp.c_proj.bias"]  # :306
        h = h + m  # :348
    h = layer_norm(h, w["ln_f.weight"], w["ln_f.bias"], spec.ln_eps)  # :492
    hidden.append(h)
    return hidden
