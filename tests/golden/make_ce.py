"""Generate the cross-encoder (continuation log-likelihood) fixture by EXECUTING the reference's own scoring functions.

    python tests/golden/make_ce.py     # needs /root/reference and `transformers`; writes ce_tiny.npz here

crossencoder/beir/sgptce.py cannot be imported (it parses argv and loads a hub model at import time), so the
definitions this fixture needs — ``group``, ``Reorderer``, ``chunks``, ``_model_call`` and ``_loglikelihood_tokens``
(sgptce.py:76-262) — are taken from its syntax tree and executed as they are, against HuggingFace
``GPTNeoForCausalLM`` built from the st_tiny fixture weights (LM head tied to wte, like the GPT-Neo checkpoints the
reference scores with).  Nothing of that file is copied into the repository.
"""
import ast
import collections
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/crossencoder/beir/sgptce.py"
WANTED = ("group", "Reorderer", "chunks", "_model_call", "_loglikelihood_tokens")


def reference_functions():
    tree = ast.parse(open(REF).read())
    nodes = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in WANTED]
    assert sorted(n.name for n in nodes) == sorted(WANTED)
    ns = {"torch": torch, "F": F, "collections": collections, "tqdm": lambda it, disable=False: it}
    exec(compile(ast.Module(body=nodes, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    from transformers import GPTNeoForCausalLM

    ns = reference_functions()
    model = GPTNeoForCausalLM.from_pretrained(os.path.join(HERE, "st_tiny")).float().eval()
    assert model.lm_head.weight.data_ptr() == model.transformer.wte.weight.data_ptr()  # tied
    g = torch.Generator().manual_seed(11)

    def toks(n):
        return torch.randint(0, 299, (n,), generator=g).tolist()

    max_length, instruction_len = 32, 3
    requests = []
    for i, (nc, nq) in enumerate([(10, 4), (1, 6), (25, 7), (40, 5), (60, 9), (3, 1), (31, 2), (12, 12)]):
        requests.append((("ctx%d" % i, "cont%d" % i), toks(nc), toks(nq)))
    requests.append((("dup", "dup"), list(requests[0][1]), list(requests[0][2])))  # identical tokens: grouped by Reorderer
    out = {}
    for bs in (1, 4):
        res = ns["_loglikelihood_tokens"](requests, model, max_length, torch.device("cpu"), disable_tqdm=True,
                                          batch_size=bs, instruction_len=instruction_len)
        out[bs] = np.array(res, dtype=np.float64)
    assert np.abs(out[1] - out[4]).max() < 1e-4, (out[1], out[4])  # right padding does not change causal logits
    ctx = [r[1] for r in requests]
    cont = [r[2] for r in requests]
    np.savez_compressed(os.path.join(HERE, "ce_tiny.npz"), max_length=max_length, instruction_len=instruction_len,
                        ctx_flat=np.concatenate([np.array(c, dtype=np.int64) for c in ctx]),
                        ctx_off=np.cumsum([0] + [len(c) for c in ctx]),
                        cont_flat=np.concatenate([np.array(c, dtype=np.int64) for c in cont]),
                        cont_off=np.cumsum([0] + [len(c) for c in cont]), loglik=out[1])
    print("loglik", np.round(out[1], 3))


if __name__ == "__main__":
    main()
