"""Oracle: SGPT pooling over a padded [B,S,d] hidden state (test infrastructure only).

Restates biencoder/beir/beir_dense_retriever.py:238-282 (script path, "BDR") and
biencoder/nli_msmarco/sentence-transformers/sentence_transformers/models/Pooling.py:85-168 (ST path).
The two differ only in the denominator clamp (ST: clamp(min=1e-9), Pooling.py:122; script: none) and in how the
last token is located; both variants are exposed.
"""
from __future__ import annotations

import torch


def weighted_mean(hidden: torch.Tensor, attention_mask: torch.Tensor, clamp: bool = False) -> torch.Tensor:
    """weightedmean: w_i = i + 1 for i = index in the PADDED row; sum_i h_i m_i w_i / sum_i m_i w_i.

    BDR:258-270 (no clamp) / Pooling.py:99-125 (clamp).  The mask is .float() so the result is fp32 (BDR:210-215).
    """
    B, S, D = hidden.shape
    mask = attention_mask.unsqueeze(-1).expand(hidden.size()).float()  # BDR:210-215
    weights = torch.arange(start=1, end=S + 1).unsqueeze(0).unsqueeze(-1).expand(hidden.size()).float()  # BDR:259-265
    num = torch.sum(hidden * mask * weights, dim=1)  # BDR:267
    den = torch.sum(mask * weights, dim=1)  # BDR:268
    if clamp:
        den = torch.clamp(den, min=1e-9)  # Pooling.py:122
    return num / den  # BDR:270


def mean(hidden: torch.Tensor, attention_mask: torch.Tensor, clamp: bool = False) -> torch.Tensor:
    """mean: BDR:238-242 / Pooling.py:98,114-125."""
    mask = attention_mask.unsqueeze(-1).expand(hidden.size()).float()
    num = torch.sum(hidden * mask, dim=1)
    den = mask.sum(dim=1)
    if clamp:
        den = torch.clamp(den, min=1e-9)
    return num / den


def last_token(hidden: torch.Tensor, attention_mask: torch.Tensor, st_variant: bool = False) -> torch.Tensor:
    """lasttoken: hidden state of the last attended token of each row.

    Script path (default), BDR:271-282: gathers index len-1, recorded at tokenisation time (BDR:198).
    ST path (st_variant=True), Pooling.py:129-164: index = argmin(mask) - 1 (first 0 of a right-padded mask, minus one),
    clamped to >= 0, and the gathered vector is multiplied by its mask value.  NOTE the ST code mis-handles rows WITHOUT
    padding: argmin of an all-ones mask is 0, so it returns token 0 instead of the last token.  The CUDA path follows
    the script semantics; the ST variant is restated only so the oracle can be pinned against Pooling.py's own output.
    """
    B, S, D = hidden.shape
    if st_variant:
        idx = torch.argmin(attention_mask.long(), dim=1) - 1  # Pooling.py:134
        idx = torch.clamp(idx, min=0)  # Pooling.py:147
        m = attention_mask.unsqueeze(-1).expand(hidden.size()).float()
        return (hidden * m)[torch.arange(B), idx]  # Pooling.py:158-159
    idx = attention_mask.long().sum(dim=1) - 1  # right padding: last 1 of the mask == len - 1 (BDR:198)
    idx = torch.clamp(idx, min=0)
    return hidden[torch.arange(B), idx]


def mean_mean(all_hidden, attention_mask: torch.Tensor) -> torch.Tensor:
    """meanmean, BDR:243-257: sum over ALL L+1 hidden states and all tokens of h * mask, divided by the equally expanded
    mask sum — i.e. the average over the hidden states of the per-state masked mean."""
    hs = torch.stack(list(all_hidden))  # BDR:246  [L+1, B, S, d]
    m = attention_mask.unsqueeze(-1).expand(hs.shape[1:]).float().unsqueeze(0).expand(hs.size())  # BDR:248
    num = torch.sum(torch.sum(hs * m, dim=2), dim=0)  # BDR:252-254
    den = m.sum(dim=2).sum(dim=0)  # BDR:255
    return num / den  # BDR:257


def last_token_mean(all_hidden, attention_mask: torch.Tensor) -> torch.Tensor:
    """lasttokenmean, BDR:284-301: the last attended token (index len-1, BDR:198) gathered from every hidden state,
    averaged over the L+1 states."""
    hs = torch.stack(list(all_hidden))  # BDR:288
    B = hs.shape[1]
    idx = torch.clamp(attention_mask.long().sum(dim=1) - 1, min=0)
    emb = hs[:, torch.arange(B), idx]  # BDR:291-298  [L+1, B, d]
    return torch.mean(emb, 0)  # BDR:301


def normalize(x: torch.Tensor) -> torch.Tensor:
    """models/Normalize.py:13-14: F.normalize(p=2, dim=1) == x / max(||x||, 1e-12)."""
    return x / torch.clamp(torch.linalg.vector_norm(x, dim=1, keepdim=True), min=1e-12)
