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


def gptj_lm_head(vocab, d, seed=123):
    """Untied GPT-J LM head (lm_head.weight / lm_head.bias, HF:gptj GPTJForCausalLM): bf16-representable, seeded so the
    GPU test can rebuild it without storing a megabyte of weights."""
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(vocab, d, generator=g) * 0.05).to(torch.bfloat16).float()
    b = (torch.randn(vocab, generator=g) * 0.5).float()
    return w, b


def gptj_case(ns):
    """GPT-J (the architecture of SGPT-5.8B / SGPT-CE-6.1B): rotary attention, parallel residual, untied LM head + bias."""
    from transformers import GPTJConfig, GPTJForCausalLM

    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import gptj as ogptj

    js = ogptj.GPTJSpec(n_layer=3, d_model=256, n_head=2, d_ff=1024, vocab=1000, max_pos=128, rotary_dim=32)
    w = ogptj.init_weights(js, 0)
    cfg = GPTJConfig(vocab_size=js.vocab, n_positions=js.max_pos, n_embd=js.d_model, n_layer=js.n_layer, n_head=js.n_head,
                     rotary_dim=js.rotary_dim, n_inner=js.d_ff, activation_function="gelu_new", resid_pdrop=0.0,
                     embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=js.ln_eps, tie_word_embeddings=False)
    model = GPTJForCausalLM(cfg)
    hw, hb = gptj_lm_head(js.vocab, js.d_model)
    sd = {"transformer." + k: v for k, v in w.items()}
    sd["lm_head.weight"], sd["lm_head.bias"] = hw, hb
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("embed_positions" in k or "attn.bias" in k or "masked_bias" in k for k in missing), (missing, unexpected)
    model = model.float().eval()
    g = torch.Generator().manual_seed(21)

    def toks(n):
        return torch.randint(0, js.vocab, (n,), generator=g).tolist()

    max_length, instruction_len = 48, 5
    requests = [((f"c{i}", f"q{i}"), toks(nc), toks(nq)) for i, (nc, nq) in
                enumerate([(20, 6), (2, 3), (47, 1), (70, 8), (100, 11), (9, 9)])]
    res = ns["_loglikelihood_tokens"](requests, model, max_length, torch.device("cpu"), disable_tqdm=True, batch_size=3,
                                      instruction_len=instruction_len)
    ctx, cont = [r[1] for r in requests], [r[2] for r in requests]
    np.savez_compressed(os.path.join(HERE, "ce_gptj_tiny.npz"), max_length=max_length, instruction_len=instruction_len,
                        spec=np.array([js.n_layer, js.d_model, js.n_head, js.d_ff, js.vocab, js.max_pos, js.rotary_dim]),
                        weight_seed=0, head_seed=123,
                        ctx_flat=np.concatenate([np.array(c, dtype=np.int64) for c in ctx]),
                        ctx_off=np.cumsum([0] + [len(c) for c in ctx]),
                        cont_flat=np.concatenate([np.array(c, dtype=np.int64) for c in cont]),
                        cont_off=np.cumsum([0] + [len(c) for c in cont]), loglik=np.array(res, dtype=np.float64))
    print("gptj loglik", np.round(res, 3))


def main():
    from transformers import GPTNeoForCausalLM

    ns = reference_functions()
    gptj_case(ns)
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
