"""The two `beir` pieces the reference's retrieval script needs around the hot path, so that
`python -m sgpt_b200.retrieve` can run the flow of biencoder/beir/beir_dense_retriever.py:352-498 with no `beir` /
`pytrec_eval` installed (both are absent offline):

* ``GenericDataLoader`` — reader of the BEIR on-disk format the script loads at BDR:379
  (``corpus.jsonl`` {"_id","title","text"}, ``queries.jsonl`` {"_id","text"}, ``qrels/<split>.tsv`` with a header row and
  ``query-id<TAB>corpus-id<TAB>score``); queries are restricted to those that have qrels, as beir 0.2.3 does.
* ``EvaluateRetrieval`` — ``retrieve`` (BDR:440-442: ``retriever.search(corpus, queries, top_k, score_function)``) and
  ``evaluate`` (BDR:446): NDCG@k, MAP@k, Recall@k, P@k with trec_eval's definitions (what beir computes through
  pytrec_eval): linear gains, log2(rank+1) discount, ideal ranking from the judged documents, ties in score broken by
  descending document id, means over the queries present in both qrels and results, rounded to 5 decimals.

Host-side Python only (no CUDA); none of it is on the measured path.
"""
from __future__ import annotations

import csv
import json
import logging
import math
import os
from typing import Dict, List, Tuple

logger = logging.getLogger(__name__)


class GenericDataLoader:
    def __init__(self, data_folder: str = None, prefix: str = None, corpus_file: str = "corpus.jsonl",
                 query_file: str = "queries.jsonl", qrels_folder: str = "qrels", qrels_file: str = ""):
        self.corpus: Dict[str, Dict[str, str]] = {}
        self.queries: Dict[str, str] = {}
        self.qrels: Dict[str, Dict[str, int]] = {}
        if prefix:
            query_file = prefix + "-" + query_file
            qrels_folder = prefix + "-" + qrels_folder
        self.corpus_file = os.path.join(data_folder, corpus_file) if data_folder else corpus_file
        self.query_file = os.path.join(data_folder, query_file) if data_folder else query_file
        self.qrels_folder = os.path.join(data_folder, qrels_folder) if data_folder else None
        self.qrels_file = qrels_file

    @staticmethod
    def check(fIn: str, ext: str):
        if not os.path.exists(fIn):
            raise ValueError("File {} not present! Please provide accurate file.".format(fIn))
        if not fIn.endswith(ext):
            raise ValueError("File {} must be present with extension {}".format(fIn, ext))

    def load(self, split: str = "test") -> Tuple[Dict[str, Dict[str, str]], Dict[str, str], Dict[str, Dict[str, int]]]:
        self.qrels_file = os.path.join(self.qrels_folder, split + ".tsv")
        self.check(self.corpus_file, "jsonl")
        self.check(self.query_file, "jsonl")
        self.check(self.qrels_file, "tsv")
        if not self.corpus:
            with open(self.corpus_file, encoding="utf8") as f:
                for line in f:
                    row = json.loads(line)
                    self.corpus[row.get("_id")] = {"text": row.get("text"), "title": row.get("title")}
            logger.info("Loaded %d %s Documents.", len(self.corpus), split.upper())
        if not self.queries:
            with open(self.query_file, encoding="utf8") as f:
                for line in f:
                    row = json.loads(line)
                    self.queries[row.get("_id")] = row.get("text")
        if os.path.exists(self.qrels_file):
            with open(self.qrels_file, encoding="utf-8") as f:
                reader = csv.reader(f, delimiter="\t", quoting=csv.QUOTE_MINIMAL)
                next(reader)  # header
                for row in reader:
                    self.qrels.setdefault(row[0], {})[row[1]] = int(row[2])
            self.queries = {qid: self.queries[qid] for qid in self.qrels}
            logger.info("Loaded %d %s Queries.", len(self.queries), split.upper())
        return self.corpus, self.queries, self.qrels


def _ranked(run: Dict[str, float]) -> List[str]:
    """trec_eval ranking: by score descending, ties by document id descending."""
    return [d for d, _ in sorted(run.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)]


def evaluate_run(qrels: Dict[str, Dict[str, int]], results: Dict[str, Dict[str, float]], k_values: List[int]
                 ) -> Tuple[Dict[str, float], Dict[str, float], Dict[str, float], Dict[str, float]]:
    """ndcg_cut / map_cut / recall / P at every k of `k_values`, averaged over the queries in qrels ∩ results."""
    ndcg = {f"NDCG@{k}": 0.0 for k in k_values}
    _map = {f"MAP@{k}": 0.0 for k in k_values}
    recall = {f"Recall@{k}": 0.0 for k in k_values}
    precision = {f"P@{k}": 0.0 for k in k_values}
    qids = [q for q in results if q in qrels]
    for qid in qids:
        rel = qrels[qid]
        ranking = _ranked(results[qid])
        gains = [max(rel.get(d, 0), 0) for d in ranking]
        ideal = sorted((g for g in rel.values() if g > 0), reverse=True)
        n_rel = len(ideal)
        for k in k_values:
            top = gains[:k]
            dcg = sum(g / math.log2(r + 2) for r, g in enumerate(top))
            idcg = sum(g / math.log2(r + 2) for r, g in enumerate(ideal[:k]))
            ndcg[f"NDCG@{k}"] += dcg / idcg if idcg > 0 else 0.0
            hits, ap = 0, 0.0
            for r, g in enumerate(top):
                if g > 0:
                    hits += 1
                    ap += hits / (r + 1)
            _map[f"MAP@{k}"] += ap / n_rel if n_rel else 0.0
            recall[f"Recall@{k}"] += hits / n_rel if n_rel else 0.0
            precision[f"P@{k}"] += hits / k
    n = max(len(qids), 1)
    for d in (ndcg, _map, recall, precision):
        for key in d:
            d[key] = round(d[key] / n, 5)
    return ndcg, _map, recall, precision


class EvaluateRetrieval:
    def __init__(self, retriever=None, k_values: List[int] = (1, 3, 5, 10, 100, 1000), score_function: str = "cos_sim"):
        self.k_values = list(k_values)
        self.top_k = max(self.k_values)
        self.retriever = retriever
        self.score_function = score_function

    def retrieve(self, corpus: Dict[str, Dict[str, str]], queries: Dict[str, str], **kwargs) -> Dict[str, Dict[str, float]]:
        if not self.retriever:
            raise ValueError("Model/Technique has not been provided!")
        return self.retriever.search(corpus, queries, self.top_k, self.score_function, **kwargs)

    @staticmethod
    def evaluate(qrels: Dict[str, Dict[str, int]], results: Dict[str, Dict[str, float]], k_values: List[int],
                 ignore_identical_ids: bool = True):
        if ignore_identical_ids:
            logger.info("For evaluation, we ignore identical query and document ids (default).")
            results = {qid: {pid: s for pid, s in rels.items() if pid != qid} for qid, rels in results.items()}
        ndcg, _map, recall, precision = evaluate_run(qrels, results, list(k_values))
        for metric in (ndcg, _map, recall, precision):
            logger.info("\n")
            for k, v in metric.items():
                logger.info("{}: {:.4f}".format(k, v))
        return ndcg, _map, recall, precision
