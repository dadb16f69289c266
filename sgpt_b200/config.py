"""Encoder hyper-parameters for the model families the reference evaluates."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List


@dataclass
class ModelConfig:
    arch: str = "gpt_neo"          # "gpt_neo" | "gptj" | "bloom"
    n_layer: int = 12
    d_model: int = 768
    n_head: int = 12
    d_ff: int = 3072
    vocab: int = 50257
    max_pos: int = 2048            # learned positions (GPT-Neo) / rotary table length (GPT-J); unused by BLOOM
    window: int = 256              # GPT-Neo local attention window
    rotary_dim: int = 0            # GPT-J
    ln_eps: float = 1e-5
    attention_layers: List[str] = field(default_factory=list)  # GPT-Neo: "global"/"local" per layer

    def __post_init__(self):
        if self.arch not in ("gpt_neo", "gptj", "bloom"):
            raise ValueError(f"unknown arch {self.arch!r}")
        if self.arch == "gpt_neo" and not self.attention_layers:
            self.attention_layers = ["global" if i % 2 == 0 else "local" for i in range(self.n_layer)]

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_head

    @classmethod
    def from_hf(cls, hf_config) -> "ModelConfig":
        """Build from a HuggingFace config object (GPTNeoConfig / GPTJConfig / BloomConfig)."""
        mt = getattr(hf_config, "model_type", "")
        if mt == "gpt_neo":
            inter = hf_config.intermediate_size or 4 * hf_config.hidden_size
            return cls(arch="gpt_neo", n_layer=hf_config.num_layers, d_model=hf_config.hidden_size,
                       n_head=hf_config.num_heads, d_ff=inter, vocab=hf_config.vocab_size,
                       max_pos=hf_config.max_position_embeddings, window=hf_config.window_size,
                       ln_eps=hf_config.layer_norm_epsilon, attention_layers=list(hf_config.attention_layers))
        if mt == "gptj":
            inter = hf_config.n_inner or 4 * hf_config.n_embd
            return cls(arch="gptj", n_layer=hf_config.n_layer, d_model=hf_config.n_embd, n_head=hf_config.n_head,
                       d_ff=inter, vocab=hf_config.vocab_size, max_pos=hf_config.n_positions,
                       rotary_dim=hf_config.rotary_dim or (hf_config.n_embd // hf_config.n_head),
                       ln_eps=hf_config.layer_norm_epsilon)
        if mt == "bloom":
            return cls(arch="bloom", n_layer=hf_config.n_layer, d_model=hf_config.hidden_size, n_head=hf_config.n_head,
                       d_ff=4 * hf_config.hidden_size, vocab=hf_config.vocab_size, max_pos=1 << 20,
                       ln_eps=hf_config.layer_norm_epsilon)
        raise NotImplementedError(f"model_type {mt!r} is not supported (gpt_neo, gptj, bloom are)")


PRESETS = {
    # SGPT-125M-weightedmean-* (EleutherAI/gpt-neo-125m)
    "sgpt-125m": dict(arch="gpt_neo", n_layer=12, d_model=768, n_head=12, d_ff=3072),
    # SGPT-1.3B-weightedmean-* (EleutherAI/gpt-neo-1.3B)
    "sgpt-1.3b": dict(arch="gpt_neo", n_layer=24, d_model=2048, n_head=16, d_ff=8192),
    # SGPT-2.7B-weightedmean-* (EleutherAI/gpt-neo-2.7B)
    "sgpt-2.7b": dict(arch="gpt_neo", n_layer=32, d_model=2560, n_head=20, d_ff=10240),
    # SGPT-5.8B-weightedmean-* (EleutherAI/gpt-j-6B)
    "sgpt-5.8b": dict(arch="gptj", n_layer=28, d_model=4096, n_head=16, d_ff=16384, vocab=50400, rotary_dim=64),
    # sgpt-bloom-7b1-msmarco (bigscience/bloom-7b1)
    "sgpt-bloom-7b1": dict(arch="bloom", n_layer=30, d_model=4096, n_head=32, d_ff=16384, vocab=250880, max_pos=1 << 20),
}


def preset(name: str, **overrides) -> ModelConfig:
    kw = dict(PRESETS[name])
    kw.update(overrides)
    return ModelConfig(**kw)
