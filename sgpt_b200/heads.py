"""Sentence-embedding heads that follow the pooling step: ``Dense`` and ``Normalize``.

``DenseHead`` is ``sentence_transformers/models/Dense.py:11-52`` applied to the pooled embedding (``key_name ==
"sentence_embedding"``): ``y = activation(x @ W^T + b)`` in fp32, run by ``sgpt_dense`` (include/sgpt_b200.h).  The
``Asym`` wrapper of the reference (models/Asym.py, used by the `asym` checkpoints) routes queries and documents to
different Dense stacks: ``AsymHeads`` holds one list of heads per text key.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib

# activation_function strings written by Dense.get_config_dict (fullname of the nn.Module class), Dense.py:49
ACTIVATIONS = {
    "torch.nn.modules.linear.Identity": _lib.ACT_IDENTITY,
    "torch.nn.modules.activation.Tanh": _lib.ACT_TANH,
    "torch.nn.modules.activation.ReLU": _lib.ACT_RELU,
    "torch.nn.modules.activation.Sigmoid": _lib.ACT_SIGMOID,
    "identity": _lib.ACT_IDENTITY, "tanh": _lib.ACT_TANH, "relu": _lib.ACT_RELU, "sigmoid": _lib.ACT_SIGMOID,
}


def activation_id(name) -> int:
    """Dense.activation_function (class path string, short name, or an nn.Module instance) -> SGPT_ACT_* code."""
    if not isinstance(name, str):
        name = type(name).__module__ + "." + type(name).__name__
    if name not in ACTIVATIONS:
        raise NotImplementedError(f"Dense activation {name!r}: built: Identity, Tanh, ReLU, Sigmoid")
    return ACTIVATIONS[name]


class DenseHead:
    """fp32 ``[out, in]`` weight (+ optional bias) resident on the device; call with fp32 ``[B, in]`` embeddings."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, activation="tanh", device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("sgpt_b200.DenseHead needs a CUDA device; there is no CPU path")
        if weight.dim() != 2:
            raise ValueError(f"Dense weight must be [out_features, in_features], got {tuple(weight.shape)}")
        self.out_features, self.in_features = int(weight.shape[0]), int(weight.shape[1])
        if bias is not None and tuple(bias.shape) != (self.out_features,):
            raise ValueError(f"Dense bias shape {tuple(bias.shape)} does not match out_features {self.out_features}")
        self.activation = activation_id(activation)
        self.weight = weight.detach().to(self.device, torch.float32).contiguous()
        self.bias = None if bias is None else bias.detach().to(self.device, torch.float32).contiguous()

    def __call__(self, emb: torch.Tensor) -> torch.Tensor:
        if emb.dim() != 2 or emb.shape[1] != self.in_features:
            raise ValueError(f"Dense expects [B, {self.in_features}] embeddings, got {tuple(emb.shape)}")
        x = emb.to(self.device, torch.float32).contiguous()
        y = torch.empty((x.shape[0], self.out_features), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().sgpt_dense(x.data_ptr(), self.weight.data_ptr(), _lib.ptr(self.bias), y.data_ptr(),
                                       x.shape[0], self.in_features, self.out_features, self.activation,
                                       _lib.current_stream())
        _lib.check(rc, "sgpt_dense")
        return y


class AsymHeads:
    """models/Asym.py:8-60: a dict ``{text_key: [heads]}``; ``apply(emb, key)`` runs the stack registered for `key`
    (the reference picks the stack from the ``{"QRY": text}`` / ``{"DOCPOS": text}`` input dict key)."""

    def __init__(self, heads: Dict[str, Sequence[DenseHead]]):
        self.heads = {k: list(v) for k, v in heads.items()}

    def apply(self, emb: torch.Tensor, key: str) -> torch.Tensor:
        if key not in self.heads:
            raise KeyError(f"no Asym head registered for text key {key!r} (have {sorted(self.heads)})")
        for h in self.heads[key]:
            emb = h(emb)
        return emb


def normalize_rows_(emb: torch.Tensor) -> torch.Tensor:
    """In-place L2 normalisation of fp32 [B, d] rows on the device (models/Normalize.py) with the library's kernel."""
    if emb.dtype != torch.float32 or not emb.is_cuda or not emb.is_contiguous():
        raise ValueError("normalize_rows_ expects a contiguous fp32 CUDA tensor")
    with torch.cuda.device(emb.device):
        _lib.check(_lib.lib().sgpt_normalize_rows(emb.data_ptr(), emb.shape[0], emb.shape[1], _lib.current_stream()),
                   "sgpt_normalize_rows")
    return emb


def apply_heads(emb: torch.Tensor, heads: Optional[List[DenseHead]]) -> torch.Tensor:
    for h in heads or ():
        emb = h(emb)
    return emb
