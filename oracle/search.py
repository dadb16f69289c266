"""Oracle: dense exact search — cos_sim / dot_score, chunked top-k and the cross-chunk merge (test infra only).

Restates biencoder/nli_msmarco/sentence-transformers/sentence_transformers/util.py:24-63 (identical to
``beir.util.cos_sim/dot_score`` imported at biencoder/beir/custommodels/exact_search.py:9) and the search loop of
biencoder/beir/custommodels/exact_search.py:34-134 ("XS").
"""
from __future__ import annotations

import heapq
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch


def _as_2d(a) -> torch.Tensor:
    if not isinstance(a, torch.Tensor):
        a = torch.tensor(np.asarray(a))  # util.py:29-33
    if a.dim() == 1:
        a = a.unsqueeze(0)  # util.py:35-39
    return a


def cos_sim(a, b) -> torch.Tensor:
    """util.py:24-43: normalize rows (x / max(||x||, 1e-12)), then a_n @ b_n^T."""
    a, b = _as_2d(a), _as_2d(b)
    a_norm = a / torch.clamp(torch.linalg.vector_norm(a, dim=1, keepdim=True), min=1e-12)
    b_norm = b / torch.clamp(torch.linalg.vector_norm(b, dim=1, keepdim=True), min=1e-12)
    return torch.mm(a_norm, b_norm.transpose(0, 1))


def dot_score(a, b) -> torch.Tensor:
    """util.py:46-63."""
    a, b = _as_2d(a), _as_2d(b)
    return torch.mm(a, b.transpose(0, 1))


SCORE_FUNCTIONS = {"cos_sim": cos_sim, "dot": dot_score}


def search_embeddings(query_ids: Sequence[str], query_emb, corpus_ids: Sequence[str], corpus_emb, top_k: int,
                      score_function: str = "cos_sim", corpus_chunk_size: int = 50000) -> Dict[str, Dict[str, float]]:
    """The scoring half of DenseRetrievalExactSearch.search (XS:80-134) on already-computed embeddings.

    For each chunk of `corpus_chunk_size` docs: scores (XS:96-98) -> NaN := -1 (XS:99) -> topk(min(k+1, n)) unsorted
    (XS:102-108) -> per query: skip corpus_id == query_id (XS:118), insert into the result dict (XS:119), and after the
    first chunk keep only heapq.nlargest(min(k+1, len)) entries (XS:121-132).
    Returns {qid: {cid: score}} with up to k+1 entries per query, exactly like the reference.
    """
    if score_function not in SCORE_FUNCTIONS:
        raise ValueError(
            "score function: {} must be either (cos_sim) for cosine similarity or (dot) for dot product".format(
                score_function))  # XS:46-51
    fn = SCORE_FUNCTIONS[score_function]
    q = _as_2d(query_emb).float()
    c = _as_2d(corpus_emb).float()
    results: Dict[str, Dict[str, float]] = {qid: {} for qid in query_ids}
    for batch_num, start in enumerate(range(0, len(corpus_ids), corpus_chunk_size)):
        end = min(start + corpus_chunk_size, len(corpus_ids))
        scores = fn(q, c[start:end])
        scores[torch.isnan(scores)] = -1
        vals, idx = torch.topk(scores, min(top_k + 1, scores.shape[1]), dim=1, largest=True, sorted=False)
        vals, idx = vals.tolist(), idx.tolist()
        for qi, qid in enumerate(query_ids):
            for sub, score in zip(idx[qi], vals[qi]):
                cid = corpus_ids[start + sub]
                if cid != qid:
                    results[qid][cid] = score
            if batch_num > 0:
                keep = heapq.nlargest(min(top_k + 1, len(results[qid])), results[qid], key=results[qid].get)
                results[qid] = {k: results[qid][k] for k in keep}
    return results


def topk_ids(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Plain torch.topk on a dense score matrix, sorted descending (ties: lowest index first)."""
    k = min(k, scores.shape[1])
    # stable sort on (-score, index) to make tie order deterministic for comparisons
    order = torch.argsort(-scores.double(), dim=1, stable=True)[:, :k]
    return torch.gather(scores, 1, order), order


def semantic_search(query_emb, corpus_emb, query_chunk_size: int = 100, corpus_chunk_size: int = 500000,
                    top_k: int = 10, score_function=cos_sim) -> List[List[dict]]:
    """util.py:197-258, restated: chunked scores + topk per chunk, then sort and trim per query."""
    q = _as_2d(query_emb)
    c = _as_2d(corpus_emb)
    out: List[List[dict]] = [[] for _ in range(len(q))]
    for qs in range(0, len(q), query_chunk_size):
        for cs in range(0, len(c), corpus_chunk_size):
            sc = score_function(q[qs:qs + query_chunk_size], c[cs:cs + corpus_chunk_size])
            vals, idx = torch.topk(sc, min(top_k, sc.shape[1]), dim=1, largest=True, sorted=False)
            vals, idx = vals.tolist(), idx.tolist()
            for qi in range(len(sc)):
                for sub, score in zip(idx[qi], vals[qi]):
                    out[qs + qi].append({"corpus_id": cs + sub, "score": score})
    for i in range(len(out)):
        out[i] = sorted(out[i], key=lambda x: x["score"], reverse=True)[:top_k]
    return out
