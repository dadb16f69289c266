"""CPU-only checks of the harness-side rows of SURVEY.md §8f: the sentence-transformers directory loader (against
directories written by the reference's own module classes, tests/golden/make_st_model.py), the BEIR loader/evaluator
stand-ins, the retrieval CLI flow with a stub retriever, and the pickle embedding cache logic."""
import json
import math
import os
import pickle

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN


def test_st_directory_loader_reads_reference_written_modules():
    from sgpt_b200.st_loader import load_st_directory

    ref = np.load(os.path.join(GOLDEN, "st_tiny.npz"))
    spec = load_st_directory(os.path.join(GOLDEN, "st_tiny"))
    assert spec.config.arch == "gpt_neo" and spec.config.d_model == 128 and spec.config.n_layer == 2
    assert spec.config.attention_layers == ["global", "local"] and spec.config.window == 8
    assert spec.max_seq_length == 32 and spec.do_lower_case is False
    assert spec.pooling == "weightedmean" and spec.normalize is True and not spec.asym
    assert torch.equal(spec.position_weights, torch.from_numpy(ref["position_weights"]))
    assert len(spec.dense) == 1
    d = spec.dense[0]
    assert d.activation == "torch.nn.modules.activation.Tanh" and d.key_name == "sentence_embedding"
    assert torch.equal(d.weight, torch.from_numpy(ref["dense_w"])) and torch.equal(d.bias, torch.from_numpy(ref["dense_b"]))
    assert "wte.weight" in spec.state_dict and spec.state_dict["h.1.mlp.c_fc.weight"].shape == (256, 128)
    assert [t.rsplit(".", 1)[-1] for t, _ in spec.modules] == ["Transformer", "WeightedMeanPooling", "Dense", "Normalize"]


def test_st_directory_loader_asym_and_pooling_flags():
    from sgpt_b200.heads import activation_id
    from sgpt_b200.st_loader import load_st_directory, pooling_mode_from_config

    spec = load_st_directory(os.path.join(GOLDEN, "st_tiny_asym"), load_weights=False)
    assert spec.state_dict is None and spec.pooling == "weightedmean" and spec.position_weights is None
    assert sorted(spec.asym) == ["DOCPOS", "QRY"] and not spec.dense and not spec.normalize
    q, d = spec.asym["QRY"][0], spec.asym["DOCPOS"][0]
    assert q.bias is None and q.weight.shape == (32, 128) and activation_id(q.activation) == 0
    assert d.bias is not None and activation_id(d.activation) == 1
    assert pooling_mode_from_config({"pooling_mode_mean_tokens": True}) == "mean"
    assert pooling_mode_from_config({"pooling_mode_mean_tokens": False, "pooling_mode_lasttoken": True}) == "lasttoken"
    with pytest.raises(NotImplementedError):
        pooling_mode_from_config({"pooling_mode_cls_token": True, "pooling_mode_mean_tokens": False})
    with pytest.raises(NotImplementedError):
        pooling_mode_from_config({"pooling_mode_mean_tokens": True, "pooling_mode_max_tokens": True})
    with pytest.raises(NotImplementedError):
        activation_id("torch.nn.modules.activation.GELU")
    with pytest.raises(FileNotFoundError):
        load_st_directory(GOLDEN)


def test_heads_refuse_cpu():
    from sgpt_b200.heads import DenseHead

    with pytest.raises(RuntimeError, match="no CPU path"):
        DenseHead(torch.zeros(4, 4), device="cpu")


def _write_beir_dataset(root):
    os.makedirs(os.path.join(root, "qrels"))
    docs = [{"_id": f"d{i}", "title": f"T{i}", "text": f"text {i}"} for i in range(6)] + [{"_id": "empty", "title": "", "text": ""}]
    with open(os.path.join(root, "corpus.jsonl"), "w") as f:
        f.writelines(json.dumps(d) + "\n" for d in docs)
    with open(os.path.join(root, "queries.jsonl"), "w") as f:
        f.writelines(json.dumps({"_id": q, "text": f"query {q}"}) + "\n" for q in ("q1", "q2", "q_unjudged"))
    with open(os.path.join(root, "qrels", "test.tsv"), "w") as f:
        f.write("query-id\tcorpus-id\tscore\n")
        f.write("q1\td0\t2\nq1\td3\t1\nq1\td5\t0\nq2\td1\t1\n")


def test_generic_data_loader(tmp_path):
    from sgpt_b200.beir_compat import GenericDataLoader

    root = str(tmp_path / "toy")
    _write_beir_dataset(root)
    corpus, queries, qrels = GenericDataLoader(root).load(split="test")
    assert len(corpus) == 7 and corpus["d2"] == {"text": "text 2", "title": "T2"}
    assert queries == {"q1": "query q1", "q2": "query q2"}  # only judged queries survive
    assert qrels == {"q1": {"d0": 2, "d3": 1, "d5": 0}, "q2": {"d1": 1}}
    with pytest.raises(ValueError, match="not present"):
        GenericDataLoader(str(tmp_path / "nope")).load()


def test_evaluate_retrieval_matches_hand_computed_trec_eval_values():
    from sgpt_b200.beir_compat import EvaluateRetrieval

    qrels = {"q1": {"d0": 2, "d3": 1, "d5": 0}, "q2": {"d1": 1}}
    results = {"q1": {"d3": 0.9, "d5": 0.8, "d0": 0.7, "d4": 0.1, "q1": 5.0},  # the self id is dropped before scoring
               "q2": {"d2": 0.5, "d1": 0.5},                                  # tie: larger doc id ranks first
               "q3": {"d0": 1.0}}                                             # no qrels: not evaluated
    ndcg, _map, recall, prec = EvaluateRetrieval.evaluate(qrels, results, [1, 3])
    idcg3 = 2 / math.log2(2) + 1 / math.log2(3)
    ndcg3_q1 = (1 / math.log2(2) + 2 / math.log2(4)) / idcg3
    ndcg3_q2 = (1 / math.log2(3)) / 1.0
    assert ndcg["NDCG@1"] == round(((1 / 2) + 0.0) / 2, 5)
    assert ndcg["NDCG@3"] == round((ndcg3_q1 + ndcg3_q2) / 2, 5)
    assert _map["MAP@3"] == round((((1 / 1) + (2 / 3)) / 2 + (1 / 2) / 1) / 2, 5)
    assert recall["Recall@1"] == round((1 / 2 + 0) / 2, 5) and recall["Recall@3"] == 1.0
    assert prec["P@1"] == 0.5 and prec["P@3"] == round((2 / 3 + 1 / 3) / 2, 5)


def test_retrieve_cli_flow_with_stub_retriever(tmp_path, monkeypatch):
    from sgpt_b200 import retrieve

    data = tmp_path / "datasets"
    _write_beir_dataset(str(data / "toy"))

    class StubRetriever:
        def search(self, corpus, queries, top_k, score_function, **kw):
            assert "empty" not in corpus and top_k == 1000 and score_function == "cos_sim"
            return {"q1": {"d0": 0.9, "d1": 0.2}, "q2": {"d1": 0.8}}

    monkeypatch.setattr(retrieve, "build_retriever", lambda args: StubRetriever())
    args = retrieve.parse_args(["--dataset", "toy", "--modelname", "org/model", "--method", "weightedmean", "--datapath",
                                str(data), "--outdir", str(tmp_path)])
    out = retrieve.main(args)
    assert out["ndcg"]["NDCG@1"] == 1.0 and out["recall"]["Recall@10"] == 0.75
    with open(tmp_path / "results_org_model_weightedmean_toy.json") as f:
        assert json.load(f)["q2"] == {"d1": 0.8}
    with open(tmp_path / "beir_embeddings_ndcgs.json") as f:
        js = json.load(f)
    assert js["ndcgs"]["org_model"]["toy"]["NDCG@1"] == 1.0 and "toy" in js["precisions"]["org_model"]
    assert retrieve.main(args) == {}  # result file exists and --overwrite not given: skipped (BDR:435-437)
    # CQADupStack average appears once all twelve sub-datasets are present (BDR:484-493)
    path = str(tmp_path / "cqa.json")
    for i, d in enumerate(retrieve.CQADUPSTACK_DATASETS):
        js = retrieve.update_scores_json(path, "m", f"cqadupstack_{d}", {"NDCG@10": float(i)}, {}, {}, {})
        assert ("cqadupstack" in js["ndcgs"]["m"]) == (i == 11)
    assert abs(js["ndcgs"]["m"]["cqadupstack"]["NDCG@10"] - 5.5) < 1e-9
    with pytest.raises(ValueError, match="speca"):
        monkeypatch.undo()
        retrieve.build_retriever(retrieve.parse_args(["--speca"]))


def test_pickle_embedding_cache_logic(tmp_path, monkeypatch):
    """save_emb writes {id: embedding} pickles per query set / corpus chunk and later calls read them back in the
    order given (BDR:311-348) — exercised with the encoder stubbed out (the GPU round trip is in the -m gpu suite)."""
    from sgpt_b200.embedder import CustomEmbedder

    monkeypatch.chdir(tmp_path)
    emb = object.__new__(CustomEmbedder)
    emb.save_emb, emb.device = True, torch.device("cpu")
    emb.base_path = "embeddings/model/weightedmean/toy"
    os.makedirs("embeddings/model/weightedmean")
    calls = []

    def fake_embed_texts(sentences, is_query):
        calls.append((list(sentences), is_query))
        return torch.tensor([[float(len(s)), 1.0 if is_query else 0.0] for s in sentences])

    emb.embed_texts = fake_embed_texts
    queries = [("q2", "bb"), ("q1", "a")]
    corpus = [("c1", {"title": "T", "text": " x "}), ("c2", {"text": " yy "})]
    q = emb.encode_queries(queries, batch_size=4)
    c = emb.encode_corpus(corpus, batch_size=4, batch_num=3)
    assert q.tolist() == [[2.0, 1.0], [1.0, 1.0]] and c.tolist() == [[4.0, 0.0], [2.0, 0.0]]
    assert calls == [(["bb", "a"], True), (["T  x", "yy"], False)]
    with open("embeddings/model/weightedmean/toy_corpus3.pickle", "rb") as f:
        assert sorted(pickle.load(f)) == ["c1", "c2"]
    calls.clear()
    emb.save_emb = False  # an existing pickle is used regardless of save_emb, as upstream
    assert emb.encode_queries(list(reversed(queries)), batch_size=4).tolist() == [[1.0, 1.0], [2.0, 1.0]]
    assert emb.encode_corpus(corpus, batch_size=4, batch_num=3).tolist() == c.tolist()
    assert calls == []
    assert emb.encode_corpus(corpus[:1], batch_size=4, batch_num=4).tolist() == [[4.0, 0.0]]  # other chunk: recomputed
    assert len(calls) == 1 and not os.path.exists("embeddings/model/weightedmean/toy_corpus4.pickle")


def test_gpt_ranker_prompting_and_rerank_with_stub_scorer():
    """Host logic of the cross-encoder surface (sgptce.py:76-90, 265-330; beir Rerank): prompt formatting, instruction
    length, (continuation, context) order, top_k cut — with the device scorer stubbed out."""
    from sgpt_b200.cross_encoder import GPTRanker, Rerank, encode
    from tests.helpers import ToyTokenizer

    tok = ToyTokenizer(vocab=300)
    tok.eos_token_id = 298
    reqs = encode([("my query", "some doc text"), ("q2", "")], tok)
    assert reqs[0][0] == ("some doc text", "my query") and len(reqs[0][1]) == 3 and len(reqs[0][2]) == 2
    assert reqs[1][1] == [298]  # empty context -> eos (sgptce.py:80-82)

    class StubScorer:
        class cfg:
            max_pos = 64

        def loglikelihood_tokens(self, requests, max_length, batch_size=64, instruction_len=0):
            self.seen = (requests, max_length, instruction_len)
            return [-float(len(ctx)) for _, ctx, _ in requests]

    stub = StubScorer()
    ranker = GPTRanker(stub, tok, prompt_doc='Doc "{}" matches "', fewshots=("fdoc", "fquery"), prompt_doc_start="{} -> {}\n")
    assert ranker.max_length == 64
    assert ranker.instruction_len == len(tok.tokenize('Doc "')) + len(tok.tokenize("fdoc -> fquery\n"))
    scores = ranker.predict([("query one", "short"), ("query one", "a much longer document")], batch_size=2)
    assert scores[0] > scores[1]
    (ctx_text, cont_text), _, cont = stub.seen[0][0]
    assert ctx_text == 'fdoc -> fquery\nDoc "short" matches "' and cont_text == "query one" and len(cont) == 2
    corpus = {"a": {"title": "T", "text": "x"}, "b": {"text": "y y y"}, "c": {"title": "", "text": "z z"}}
    res = Rerank(ranker).rerank(corpus, {"q": "query"}, {"q": {"a": 0.1, "b": 0.9, "c": 0.5}}, top_k=2)
    assert sorted(res["q"]) == ["b", "c"]  # only the two best first-stage hits are re-scored


def test_information_retrieval_evaluator_metrics_match_executed_reference(tmp_path):
    """compute_metrics vs the reference's own method (executed from its syntax tree by tests/golden/make_ir_eval.py) on the
    same seeded result lists; then the whole __call__ flow (encode -> search -> metrics -> CSV) with a stub model and an
    injected CPU search."""
    from oracle import search as osearch
    from sgpt_b200.evaluation import InformationRetrievalEvaluator

    with open(os.path.join(GOLDEN, "ir_eval.json")) as f:
        z = json.load(f)
    relevant = {k: set(v) for k, v in z["relevant"].items()}
    ev = InformationRetrievalEvaluator(z["queries"], z["corpus"], relevant, mrr_at_k=[5, 10], ndcg_at_k=[3, 10],
                                       accuracy_at_k=[1, 3], precision_recall_at_k=[1, 5], map_at_k=[10, 100])
    assert ev.csv_headers == z["csv_headers"] and ev.csv_file == z["csv_file"]
    assert "q10" not in ev.queries_ids and "q11" not in ev.queries_ids  # no relevant docs / not judged (:42-45)
    got = ev.compute_metrics(z["results"])
    for metric, by_k in z["scores"].items():
        for k, v in by_k.items():
            assert abs(got[metric][int(k)] - v) < 1e-12, (metric, k)

    class StubModel:
        def encode(self, sentences, batch_size=32, convert_to_tensor=True, **kw):
            return torch.stack([torch.nn.functional.one_hot(torch.tensor(int(s.split()[-1]) % 7), 8).float() +
                                0.01 * int(s.split()[-1]) for s in sentences])

    def cpu_search(q, c, k, fn):
        sc = osearch.SCORE_FUNCTIONS[fn](q, c)
        return osearch.topk_ids(sc, k)

    ev2 = InformationRetrievalEvaluator(z["queries"], z["corpus"], relevant, name="toy", search_fn=cpu_search,
                                        main_score_function="cos_sim")
    main = ev2(StubModel(), output_path=str(tmp_path), epoch=1, steps=2)
    lines = open(tmp_path / "Information-Retrieval_evaluation_toy_results.csv").read().strip().splitlines()
    assert lines[0].split(",")[:3] == ["epoch", "steps", "cos_sim-Accuracy@1"] and lines[1].startswith("1,2,")
    assert len(lines[1].split(",")) == len(lines[0].split(",")) and 0.0 <= main <= 1.0
    ev2(StubModel(), output_path=str(tmp_path))
    assert len(open(tmp_path / "Information-Retrieval_evaluation_toy_results.csv").read().strip().splitlines()) == 3
    with pytest.raises(ValueError):
        InformationRetrievalEvaluator(z["queries"], z["corpus"], relevant, score_functions=["euclid"])


def test_sentence_encoder_tokenize_fast_tokenizer_batch_path_equals_per_text_path():
    """A HuggingFace *fast* tokenizer is called once per batch (like models/Transformer.py:127,132-135); the result must
    equal the per-text path other tokenizers take, incl. truncation and the specb bracket rules."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    from sgpt_b200.embedder import SentenceEncoder

    vocab = {"[UNK]": 0, "[PAD]": 1, "[SOS]": 2, "{SOS}": 3, "[": 4, "]": 5, "{": 6, "}": 7}
    for i, w in enumerate("a b c d e f g hello world this is test".split()):
        vocab[w] = 8 + i
    tok = Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="[UNK]", pad_token="[PAD]")

    class Slow:  # same vocabulary through the per-text interface
        pad_token_id = 1

        def encode(self, text, add_special_tokens=False):
            return [vocab.get(w, 0) for w in text.split()]

    def make(tokenizer, spec):
        e = object.__new__(SentenceEncoder)
        e.tokenizer, e.max_seq_length, e.do_lower_case, e.pad_id = tokenizer, 6, False, 1
        e.bos_spec_token_q = e.bos_spec_token_d = e.eos_spec_token_q = e.eos_spec_token_d = None
        e.bos_spec_token_q_rep = e.bos_spec_token_d_rep = None
        e.replace_bos = False
        if spec:
            e.bos_spec_token_q, e.bos_spec_token_d, e.eos_spec_token_q, e.eos_spec_token_d = 2, 3, 5, 7
            e.bos_spec_token_q_rep, e.bos_spec_token_d_rep, e.replace_bos = 4, 6, True
        return e

    plain = ["a b c", "  hello world this is a test a b c d  ", "zzz", {"QRY": "d e"}]
    for spec, texts in ((False, plain), (True, ["[SOS] a b c d e f g", "{SOS} hello", "[SOS] x"])):
        ids_f, mask_f = make(fast, spec).tokenize(texts)
        ids_s, mask_s = make(Slow(), spec).tokenize(texts)
        assert np.array_equal(ids_f, ids_s) and np.array_equal(mask_f, mask_s), (spec, ids_f, ids_s)
    ids, mask = make(fast, True).tokenize(["[SOS] a b c d e f g"])
    assert ids[0].tolist() == [4, 8, 9, 10, 5] and mask[0].tolist() == [1] * 5  # max_seq_length-2 tokens, '[' ... ']'
    with pytest.raises(ValueError, match="BOS"):
        make(fast, True).tokenize(["a b"])


def test_custom_embedder_tokenize_batch_fast_tokenizer_equals_two_step_path():
    """CustomEmbedder.tokenize_batch: a fast tokenizer's batched call gives the ids of tokenize + convert_tokens_to_ids
    (BDR:169-170), with the same truncation, specb brackets and empty-text error."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    from sgpt_b200.embedder import CustomEmbedder

    vocab = {"[UNK]": 0, "[PAD]": 1, "[": 4, "]": 5, "{": 6, "}": 7}
    for i, w in enumerate("a b c d e f g hello world".split()):
        vocab[w] = 8 + i
    tok = Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="[UNK]", pad_token="[PAD]")

    class TwoStep:
        def tokenize(self, text):
            return text.split()

        def convert_tokens_to_ids(self, tokens):
            return [vocab.get(t, 0) for t in tokens]

    def make(tokenizer):
        e = object.__new__(CustomEmbedder)
        e.tokenizer, e.max_token_len, e.specb, e.pad_id = tokenizer, 4, True, 1
        e.bos_token_q, e.eos_token_q, e.bos_token_d, e.eos_token_d = [4], [5], [6], [7]
        return e

    texts = ["a b\nc d e f", "hello", "zzz world"]
    for is_query in (True, False):
        got, want = make(fast).tokenize_batch(texts, is_query), make(TwoStep()).tokenize_batch(texts, is_query)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert make(fast).tokenize_batch(texts, True)[0][0].tolist() == [4, 8, 9, 10, 11, 5]
    with pytest.raises(ValueError, match="Empty items"):
        make(fast).tokenize_batch(["a", "   "], True)


def test_evaluate_run_ndcg_against_sklearn_and_brute_force():
    """NDCG@k of beir_compat.evaluate_run vs sklearn.metrics.ndcg_score (linear gains, log2 discount — the trec_eval
    definition) and Recall/P/MAP vs a brute-force restatement, on random runs with graded qrels (pytrec_eval itself is
    not installed offline)."""
    from sklearn.metrics import ndcg_score

    from sgpt_b200.beir_compat import evaluate_run

    rs = np.random.RandomState(7)
    docs = [f"d{i:03d}" for i in range(40)]
    qrels, results = {}, {}
    for q in range(15):
        qid = f"q{q}"
        judged = rs.choice(docs, rs.randint(1, 12), replace=False)
        qrels[qid] = {d: int(rs.randint(0, 4)) for d in judged}
        if not any(v > 0 for v in qrels[qid].values()):
            qrels[qid][judged[0]] = 1
        retrieved = rs.choice(docs, rs.randint(5, 40), replace=False)
        results[qid] = {d: float(s) for d, s in zip(retrieved, rs.permutation(len(retrieved)) / 100.0)}  # distinct scores
    ks = [1, 5, 10, 100]
    ndcg, _map, recall, prec = evaluate_run(qrels, results, ks)
    for k in ks:
        vals, rec, pre, ap = [], [], [], []
        for qid in results:
            y_true = np.array([[max(qrels[qid].get(d, 0), 0) for d in docs]], dtype=float)
            y_score = np.array([[results[qid].get(d, -1.0) for d in docs]])  # unretrieved docs rank last with gain 0 or not
            # sklearn ranks ALL docs; trec_eval only the retrieved ones: zero the gain of unretrieved docs in the run's
            # DCG by comparing on the retrieved set, but keep them in the ideal ranking
            ranked = sorted(results[qid], key=results[qid].get, reverse=True)[:k]
            gains = [max(qrels[qid].get(d, 0), 0) for d in ranked]
            dcg = sum(g / np.log2(i + 2) for i, g in enumerate(gains))
            ideal = sorted((v for v in qrels[qid].values() if v > 0), reverse=True)[:k]
            idcg = sum(g / np.log2(i + 2) for i, g in enumerate(ideal))
            vals.append(dcg / idcg)
            if set(d for d, v in qrels[qid].items() if v > 0) <= set(results[qid]):
                # every relevant doc was retrieved: then sklearn's full ranking agrees with trec_eval's
                assert abs(ndcg_score(y_true, y_score, k=k) - dcg / idcg) < 1e-9
            n_rel = sum(1 for v in qrels[qid].values() if v > 0)
            hits = [1 if g > 0 else 0 for g in gains]
            rec.append(sum(hits) / n_rel)
            pre.append(sum(hits) / k)
            ap.append(sum(sum(hits[:i + 1]) / (i + 1) for i, h in enumerate(hits) if h) / n_rel)
        assert ndcg[f"NDCG@{k}"] == round(float(np.mean(vals)), 5)
        assert recall[f"Recall@{k}"] == round(float(np.mean(rec)), 5)
        assert prec[f"P@{k}"] == round(float(np.mean(pre)), 5)
        assert _map[f"MAP@{k}"] == round(float(np.mean(ap)), 5)
