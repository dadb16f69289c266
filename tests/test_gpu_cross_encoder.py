"""GPU parity tests of the cross-encoder scoring path (pytest -m gpu): continuation
log-likelihoods through sgpt_forward + sgpt_lm_logprobs vs the values the REFERENCE's own functions produced
(tests/golden/make_ce.py executes crossencoder/beir/sgptce.py:76-262 from its syntax tree), the log-softmax-gather kernel
alone vs torch in fp64, and the GPTRanker / Rerank surface vs the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import gpt_neo, lm_score
from tests.conftest import GOLDEN
from tests.helpers import ToyTokenizer

pytestmark = pytest.mark.gpu

# bf16 hidden states and LM-head weights against an fp32 reference: measured per-token log-prob error is ~1e-3 on this
# model; the bar per request is 0.02 nats per continuation token
TOL_PER_TOKEN = 0.02


def _requests(z):
    co, qo = z["ctx_off"], z["cont_off"]
    return [(i, z["ctx_flat"][co[i]:co[i + 1]].tolist(), z["cont_flat"][qo[i]:qo[i + 1]].tolist())
            for i in range(len(co) - 1)]


@pytest.fixture(scope="module")
def scorer():
    from sgpt_b200.cross_encoder import LogLikelihoodScorer
    from sgpt_b200.st_loader import load_st_directory

    spec = load_st_directory(os.path.join(GOLDEN, "st_tiny"))
    s = LogLikelihoodScorer(spec.config, spec.state_dict, max_tokens=2048, max_batch=16, rows_per_chunk=8)
    yield s, spec
    s.close()


def test_loglikelihood_tokens_vs_executed_reference(scorer):
    s, _ = scorer
    z = np.load(os.path.join(GOLDEN, "ce_tiny.npz"))
    reqs = _requests(z)
    want = z["loglik"]
    for bs in (1, 4, 64):  # rows_per_chunk=8 < rows of a batch: the LM head runs in several chunks
        got = s.loglikelihood_tokens(reqs, int(z["max_length"]), batch_size=bs, instruction_len=int(z["instruction_len"]))
        assert len(got) == len(reqs) and got[0] == got[-1]
        for g, w, (_, _, cont) in zip(got, want, reqs):
            assert abs(g - w) < TOL_PER_TOKEN * len(cont), (bs, g, w, len(cont))
    assert np.argsort(got).tolist() == np.argsort(want).tolist() or np.abs(np.sort(got) - np.sort(want)).max() < 0.05
    with pytest.raises(AssertionError):
        s.loglikelihood_tokens([(0, [], [1])], 32)
    with pytest.raises(AssertionError):
        s.loglikelihood_tokens([(0, [1], list(range(40)))], 32)
    with pytest.raises(ValueError, match="vocabulary"):
        s.score_batch([[1, 2, 3]], [[5, 300]])


def test_score_batch_greedy_flags_and_ragged_equivalence(scorer):
    s, spec = scorer
    g = torch.Generator().manual_seed(5)
    inputs = [torch.randint(0, 299, (n,), generator=g).tolist() for n in (7, 30, 1, 19)]
    conts = [torch.randint(0, 299, (n,), generator=g).tolist() for n in (3, 30, 1, 5)]
    sums, is_greedy = s.score_batch(inputs, conts, return_greedy=True)
    one_by_one = [s.score_batch([i], [c]).item() for i, c in zip(inputs, conts)]
    assert np.abs(sums.cpu().numpy() - np.array(one_by_one)).max() < 1e-3  # batch composition does not matter
    assert is_greedy == [False] * 4
    # make request 0's continuation the greedy one: feed back the argmax of every predicting position
    nspec = gpt_neo.NeoSpec(n_layer=2, d_model=128, n_head=2, d_ff=256, vocab=300, max_pos=64, window=8)
    w = {k: v.float() for k, v in spec.state_dict.items()}
    with torch.no_grad():
        h = gpt_neo.forward(nspec, w, torch.tensor([inputs[0]]), torch.ones(1, 7, dtype=torch.long))[-1][0]
        logits = h @ w["wte.weight"].t()
    top2 = logits[-3:].topk(2, dim=-1).values
    if (top2[:, 0] - top2[:, 1]).min() > 0.05:  # unambiguous argmax (bf16 noise is ~1e-3)
        greedy_cont = logits[-3:].argmax(-1).tolist()
        _, flags = s.score_batch([inputs[0]], [greedy_cont], return_greedy=True)
        assert flags == [True]


def test_token_logprob_kernel_vs_torch_fp64():
    from sgpt_b200 import _lib

    g = torch.Generator().manual_seed(9)
    for M, V in ((1, 5), (7, 301), (33, 50257)):
        lds = (V + 3) // 4 * 4
        logits = torch.randn(M, lds, generator=g) * 4
        logits[:, V:] = 1e30  # padding columns must be ignored
        bias = torch.randn(V, generator=g)
        tgt = torch.randint(0, V, (M,), generator=g, dtype=torch.int32)
        logits[0, V - 1] = logits[0, :V].max() + 1  # argmax in the last valid column
        if M > 1:
            logits[1, 3 % V] = logits[1, 1 % V] = logits[1, :V].max() + 2  # tie: lowest index wins
        for b in (None, bias):
            lp = torch.empty(M, dtype=torch.float32, device="cuda")
            gr = torch.empty(M, dtype=torch.int32, device="cuda")
            ld, bd, td = logits.cuda(), (None if b is None else b.cuda()), tgt.cuda()
            rc = _lib.lib().sgpt_token_logprobs(ld.data_ptr(), lds, M, V, _lib.ptr(bd), td.data_ptr(), lp.data_ptr(),
                                                gr.data_ptr(), _lib.current_stream())
            _lib.check(rc, "sgpt_token_logprobs")
            z = logits[:, :V].double() + (0 if b is None else b.double())
            want = torch.log_softmax(z, -1).gather(1, tgt.long().unsqueeze(1)).squeeze(1)
            assert (lp.cpu().double() - want).abs().max() < 2e-5, (M, V, b is None)
            assert torch.equal(gr.cpu().long(), z.argmax(-1)), (M, V)
    x = torch.arange(10, dtype=torch.float32, device="cuda")
    off = torch.tensor([0, 3, 3, 10], dtype=torch.int32, device="cuda")
    out = torch.empty(3, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().sgpt_segment_sum(x.data_ptr(), off.data_ptr(), 3, out.data_ptr(), _lib.current_stream()))
    assert out.cpu().tolist() == [3.0, 0.0, 42.0]


def test_gpt_ranker_and_rerank_vs_oracle(scorer):
    from sgpt_b200.cross_encoder import GPTRanker, Rerank, encode

    s, spec = scorer
    tok = ToyTokenizer(vocab=300, pad_token_id=299)
    tok.eos_token_id = 298
    prompt = 'Documents are searched to find matches with the same content.\nThe document "{}" is a good search result for "'
    ranker = GPTRanker(s, tok, max_length=32, prompt_doc=prompt, batch_size=3)
    assert ranker.instruction_len == len(tok.tokenize(prompt[:prompt.index("{")]))
    corpus = {f"d{i}": {"title": f"title {i}", "text": " ".join(f"w{i}x{j}" for j in range(3 + 7 * i))} for i in range(5)}
    queries = {"q0": "alpha beta gamma", "q1": "delta"}
    first_stage = {"q0": {"d0": 3.0, "d1": 2.0, "d2": 1.0, "d3": 0.5}, "q1": {"d4": 1.0, "d0": 0.1}}
    res = Rerank(ranker, batch_size=8).rerank(corpus, queries, first_stage, top_k=3)
    assert sorted(res["q0"]) == ["d0", "d1", "d2"] and sorted(res["q1"]) == ["d0", "d4"]
    # the same pairs through the CPU oracle
    nspec = gpt_neo.NeoSpec(n_layer=2, d_model=128, n_head=2, d_ff=256, vocab=300, max_pos=64, window=8)
    w = {k: v.float() for k, v in spec.state_dict.items()}
    for qid, docs in res.items():
        pairs = [(queries[qid], prompt.format((corpus[d]["title"] + " " + corpus[d]["text"]).strip())) for d in docs]
        reqs = encode(pairs, tok)
        want = lm_score.loglikelihood(nspec, w, reqs, 32, ranker.instruction_len)
        for d, wv, (_, _, cont) in zip(docs, want, reqs):
            assert abs(res[qid][d] - wv) < TOL_PER_TOKEN * len(cont), (qid, d, res[qid][d], wv)
    assert ranker.predict([("q", "")], batch_size=1)[0] < 0  # empty document text still has the prompt as context
