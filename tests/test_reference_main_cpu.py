"""SURVEY.md §8f row 1 / VERDICT r01 item 9: the reference's own ``main`` (biencoder/beir/beir_dense_retriever.py:352-498)
runs UNMODIFIED against the stand-in ``beir`` package and ``custommodels`` of sgpt_b200/compat.  This container has the
reference but no GPU, so the two classes that launch CUDA kernels are replaced by CPU stand-ins that keep their
interfaces (a toy embedder with CustomEmbedder's constructor keywords; the oracle's restatement of the search loop behind
DenseRetrievalExactSearch's ``search``); everything else — argument parsing, GenericDataLoader, empty-text filtering,
EvaluateRetrieval.retrieve/evaluate, the result and metric files — is the reference's code driving the stand-in
package.  (On the GPU box the reference is absent; the same flow with the real classes is covered by
tests/test_gpu_heads.py and `python -m sgpt_b200.compat.run_reference`.)"""
import json
import os
import sys
import zlib

import numpy as np
import pytest
import torch

REF = "/root/reference/biencoder/beir/beir_dense_retriever.py"


class _ToyEmbedder:
    """CustomEmbedder's protocol (BDR:107-120, 316-348) with a deterministic bag-of-words embedding on the CPU."""

    def __init__(self, model_name, batch_size=250, device="cpu", save_emb=False, reinit=False, layeridx=-1, method="mean",
                 dataset="scifact", specb=False, maxseqlen=None, **kwargs):
        self.kw = dict(model_name=model_name, batch_size=batch_size, device=device, method=method, specb=specb,
                       layeridx=layeridx, maxseqlen=maxseqlen, save_emb=save_emb)
        self.device = torch.device("cpu")

    @staticmethod
    def _vec(text, dim=32):
        v = torch.zeros(dim)
        for w in text.lower().split():
            g = torch.Generator().manual_seed(zlib.crc32(w.encode()))
            v += torch.randn(dim, generator=g)
        return v

    def encode_queries(self, queries, batch_size=None, **kwargs):
        return torch.stack([self._vec(t) for _, t in queries])

    def encode_corpus(self, corpus, batch_size=None, **kwargs):
        return torch.stack([self._vec((d["title"] + " " + d["text"]).strip() if "title" in d else d["text"].strip())
                            for _, d in corpus])


class _CpuDRES:
    """DenseRetrievalExactSearch's interface (XS:22-42) on the oracle's restatement of XS:44-134."""

    def __init__(self, model, batch_size=128, corpus_chunk_size=50000, **kwargs):
        self.model, self.batch_size, self.corpus_chunk_size = model, batch_size, corpus_chunk_size
        self.results = {}

    def search(self, corpus, queries, top_k, score_function, return_sorted=False, **kwargs):
        from oracle import search as osearch

        if score_function not in ("cos_sim", "dot"):
            raise ValueError("score function: {} must be either (cos_sim) for cosine similarity or (dot) for dot product"
                             .format(score_function))
        qids = list(queries)
        cids = sorted(corpus, key=lambda k: len(corpus[k].get("title", "") + corpus[k].get("text", "")), reverse=True)
        q = self.model.encode_queries([(i, queries[i]) for i in qids], batch_size=self.batch_size)
        c = self.model.encode_corpus([(i, corpus[i]) for i in cids], batch_size=self.batch_size, batch_num=0)
        self.results = osearch.search_embeddings(qids, q, cids, c, top_k, score_function,
                                                 corpus_chunk_size=self.corpus_chunk_size)
        return self.results


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree is only present in the build container")
def test_reference_main_runs_unmodified_on_the_stand_in_packages(tmp_path, monkeypatch):
    from sgpt_b200 import compat
    from sgpt_b200.compat.run_reference import load_reference_script

    # toy BEIR directory: the relevant document of every query repeats the query's words
    rs = np.random.RandomState(0)
    words = [f"w{i}" for i in range(200)]
    root = tmp_path / "datasets" / "toyset"
    os.makedirs(root / "qrels")
    corpus = {f"d{i}": {"title": " ".join(rs.choice(words, 2)), "text": " ".join(rs.choice(words, rs.randint(3, 30)))}
              for i in range(60)}
    corpus["d_empty"] = {"title": "t", "text": ""}  # removed by the script (BDR:382-388)
    queries = {f"q{i}": corpus[f"d{i}"]["text"] for i in range(8)}
    with open(root / "corpus.jsonl", "w") as f:
        for k, v in corpus.items():
            f.write(json.dumps({"_id": k, **v}) + "\n")
    with open(root / "queries.jsonl", "w") as f:
        for k, v in queries.items():
            f.write(json.dumps({"_id": k, "text": v}) + "\n")
    with open(root / "qrels" / "test.tsv", "w") as f:
        f.write("query-id\tcorpus-id\tscore\n")
        for i in range(8):
            f.write(f"q{i}\td{i}\t1\n")

    monkeypatch.chdir(tmp_path)  # the script writes its result files into the working directory
    monkeypatch.setattr(sys, "path", list(sys.path))
    mod = load_reference_script(REF, embedder_cls=_ToyEmbedder, module_name="ref_bdr_under_test")
    import beir
    import custommodels

    assert os.path.dirname(os.path.dirname(beir.__file__)) == compat.COMPAT_DIR  # the stand-in, not an installed beir
    monkeypatch.setattr(mod, "DenseRetrievalExactSearch", _CpuDRES)  # (bound at import: `from custommodels import ...`)
    assert custommodels.DenseRetrievalExactSearch.__module__ == "sgpt_b200.exact_search"
    monkeypatch.setattr(sys, "argv", ["beir_dense_retriever.py", "--dataset", "toyset", "--datapath", str(tmp_path / "datasets"),
                                      "--modelname", "toy/model", "--method", "weightedmean", "--device", "cpu",
                                      "--batchsize", "16", "--specb"])
    args = mod.parse_args()
    mod.main(args)

    results = json.load(open(tmp_path / "results_toy_model_weightedmean_toyset.json"))
    assert set(results) == set(queries)
    for i in range(8):
        assert max(results[f"q{i}"], key=results[f"q{i}"].get) == f"d{i}"  # the planted document ranks first
        assert "d_empty" not in results[f"q{i}"]
    nd = json.load(open(tmp_path / "beir_embeddings_ndcgs.json"))
    assert nd["ndcgs"]["toy_model"]["toyset"]["NDCG@1"] == 1.0
    assert nd["recalls"]["toy_model"]["toyset"]["Recall@10"] == 1.0
    assert set(nd) >= {"ndcgs", "maps", "recalls", "precisions"}
    # a second run finds the result file and skips (BDR:433-436)
    mod.main(args)
