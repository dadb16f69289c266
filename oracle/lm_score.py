"""Oracle: continuation log-likelihood of the SGPT cross-encoder (test infrastructure only).

Restates crossencoder/beir/sgptce.py:150-262 (``_loglikelihood_tokens``) for one request at a time, on top of the
oracle's own GPT-Neo forward (``oracle.gpt_neo.forward``): input = (instruction + left-truncated (rest of context +
continuation))[:-1] (:186-193), logits = hidden_states[-1] @ wte^T (tied LM head of HF ``GPTNeoForCausalLM``),
``log_softmax`` in fp32 (:221), the rows that predict the continuation tokens (:226), gather (:247), sum (:250).
Pinned against the reference functions themselves, executed from the reference file's syntax tree by
tests/golden/make_ce.py (fixture tests/golden/ce_tiny.npz).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def model_input(context_enc: Sequence[int], continuation_enc: Sequence[int], max_length: int, instruction_len: int = 0
                ) -> List[int]:
    """sgptce.py:186-193: keep the instruction, truncate the rest FROM THE LEFT to max_length+1 tokens, drop the last."""
    ctx, cont = list(context_enc), list(continuation_enc)
    return (ctx[:instruction_len] + (ctx[instruction_len:] + cont)[-(max_length + 1 - instruction_len):])[:-1]


def loglikelihood(spec, weights, requests: Sequence[Tuple[object, Sequence[int], Sequence[int]]], max_length: int,
                  instruction_len: int = 0, arch: str = "gpt_neo", lm_head=None, lm_bias=None) -> List[float]:
    """One float per request (cache_key, context_enc, continuation_enc): sum of log p(cont_i | prefix).
    arch "gpt_neo" (LM head tied to wte, HF GPTNeoForCausalLM) or "gptj" (untied ``lm_head`` weight [vocab, d] + bias,
    HF GPTJForCausalLM)."""
    from . import gpt_neo, gptj

    fwd = {"gpt_neo": gpt_neo.forward, "gptj": gptj.forward}[arch]
    wte = (weights["wte.weight"] if lm_head is None else lm_head).float()
    out = []
    for _, ctx, cont in requests:
        inp = model_input(ctx, cont, max_length, instruction_len)
        ids = torch.tensor([inp], dtype=torch.long)
        mask = torch.ones_like(ids)
        with torch.no_grad():
            hidden = fwd(spec, weights, ids, mask)[-1][0]                       # [S, d], after ln_f
            logits = hidden.float() @ wte.t()
            if lm_bias is not None:
                logits = logits + lm_bias.float()
            logp = torch.log_softmax(logits, dim=-1)                            # :221
        rows = logp[len(inp) - len(cont):len(inp)]                              # :226
        out.append(float(rows.gather(1, torch.tensor(list(cont)).unsqueeze(1)).sum()))  # :247-250
    return out
