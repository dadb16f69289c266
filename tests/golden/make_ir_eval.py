"""Generate the InformationRetrievalEvaluator fixture by EXECUTING the reference's metric code.

    python tests/golden/make_ir_eval.py      # needs /root/reference; writes ir_eval.json here

The class cannot be imported as a package member offline (sentence_transformers/__init__ needs hub helpers), so
``compute_metrics`` / ``compute_dcg_at_k`` / ``__init__`` (evaluation/InformationRetrievalEvaluator.py:22-88, 177-299)
are taken from the file's syntax tree and executed as they are on a seeded synthetic result list.
"""
import ast
import json
import os
import random

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = ("/root/reference/biencoder/nli_msmarco/sentence-transformers/sentence_transformers/evaluation/"
       "InformationRetrievalEvaluator.py")


def reference_class():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "InformationRetrievalEvaluator")
    cls.bases = []  # SentenceEvaluator is an empty interface
    import typing

    ns = {"np": np, "List": typing.List, "Dict": typing.Dict, "Set": typing.Set, "Tuple": typing.Tuple,
          "Callable": typing.Callable, "Tensor": object, "cos_sim": None, "dot_score": None, "logger": None}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), REF, "exec"), ns)
    return ns["InformationRetrievalEvaluator"]


def main():
    rnd = random.Random(3)
    corpus = {f"c{i}": f"doc {i}" for i in range(60)}
    queries = {f"q{i}": f"query {i}" for i in range(12)}
    relevant = {f"q{i}": set(rnd.sample(sorted(corpus), rnd.randint(1, 6))) for i in range(10)}
    relevant["q10"] = set()  # dropped: no relevant docs (:43)
    Ref = reference_class()
    ev = Ref(queries, corpus, relevant, mrr_at_k=[5, 10], ndcg_at_k=[3, 10], accuracy_at_k=[1, 3], precision_recall_at_k=[1, 5],
             map_at_k=[10, 100])
    results = []
    for qid in ev.queries_ids:
        docs = rnd.sample(sorted(corpus), 40)
        results.append([{"corpus_id": d, "score": rnd.random() + (0.5 if d in relevant[qid] else 0.0)} for d in docs])
    scores = ev.compute_metrics(results)
    out = {"queries": queries, "corpus": corpus, "relevant": {k: sorted(v) for k, v in relevant.items()},
           "results": results, "csv_headers": ev.csv_headers, "csv_file": ev.csv_file,
           "scores": {m: {str(k): float(v) for k, v in d.items()} for m, d in scores.items()}}
    with open(os.path.join(HERE, "ir_eval.json"), "w") as f:
        json.dump(out, f)
    print(json.dumps(out["scores"], indent=1)[:600])


if __name__ == "__main__":
    main()
