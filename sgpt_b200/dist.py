"""Row-sharded corpus search across the GPUs of one box.

Rank g owns documents [g*ceil(N/G), (g+1)*ceil(N/G)); every rank scans its own shard (no data-path communication),
then ONE collective — an all-gather of the per-shard top-(k+1) ``(score fp32, global id int64)`` lists,
Q*(k+1)*12 bytes per rank — followed by the same merge kernel the single-GPU chunk loop uses.  This is the reference's
sequential "chunk, top-k, heapq.nlargest merge" (exact_search.py:80-132) with "chunk" = "shard".  The only collective
the reference itself has is the padded embedding all-gather of sentence_transformers/util.py:326-347, whose role
(combine per-rank partial results) this takes over.  Works with the ``nccl`` backend on GPUs and — for the host-side
logic tests — with ``gloo`` on CPU tensors when a ``local_search`` callable is supplied.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_docs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`: [rank*ceil(N/G), min(N, (rank+1)*ceil(N/G)))."""
    per = (n_docs + world - 1) // world
    lo = min(n_docs, rank * per)
    return lo, min(n_docs, lo + per)


def all_gather_topk(scores: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[Q,k] per rank -> [G,Q,k] on every rank (scores fp32, ids int64)."""
    world = dist.get_world_size(group)
    Q, k = scores.shape
    # concatenated-along-dim-0 output layout is the one both nccl and gloo accept
    gs = torch.empty((world * Q, k), dtype=scores.dtype, device=scores.device)
    gi = torch.empty((world * Q, k), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, ids.contiguous(), group=group)
    return gs.view(world, Q, k), gi.view(world, Q, k)


def sharded_search(query_emb: torch.Tensor, shard, k: int, score_function: str = "cos_sim",
                   exclude_ids: Optional[torch.Tensor] = None, group=None,
                   merge: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank: local exact top-k over its shard -> all-gather -> merge.  Returns identical (scores, ids) [Q,k] on all
    ranks, ids being global document rows.  `shard` is a CorpusShard (or anything with .search(q, k, score_function));
    `merge` defaults to the CUDA merge kernel."""
    local_s, local_i = shard.search(query_emb, k, score_function)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        gs, gi = local_s.unsqueeze(0), local_i.unsqueeze(0)
    else:
        gs, gi = all_gather_topk(local_s, local_i, group)
    if merge is None:
        from .index import merge_topk as merge
    return merge(gs, gi, exclude_ids)


class ShardedDenseRetrievalExactSearch:
    """``DenseRetrievalExactSearch.search`` (exact_search.py:34-134) across the ranks of one ``torch.distributed`` group.

    Every rank holds a replica of the embedder.  Queries are encoded by every rank (cheap, and it keeps the result
    available everywhere); the corpus is sorted by length like the reference (XS:66-70) and dealt out ``i mod G`` over
    the ranks, so every rank encodes the same mix of long and short documents.  Rank g encodes only its
    documents into one resident ``CorpusShard``, scans it once, and the per-rank top-(k+1) lists meet in one all-gather
    followed by the merge kernel — the reference's "chunk, top-k, heapq merge" with chunk = rank.  Every rank returns the
    same ``Dict[qid, Dict[cid, float]]``.  ``shard_factory(dim, capacity)`` / ``merge`` default to ``CorpusShard`` and the
    CUDA merge kernel; the CPU tests inject stand-ins and run it on ``gloo``.
    """

    def __init__(self, model, batch_size: int = 128, corpus_chunk_size: int = 50000, group=None,
                 shard_factory: Optional[Callable] = None, merge: Optional[Callable] = None, **kwargs):
        self.model, self.batch_size, self.corpus_chunk_size = model, batch_size, corpus_chunk_size
        self.group, self.shard_factory, self.merge = group, shard_factory, merge
        self.show_progress_bar, self.convert_to_tensor = True, True
        self.results = {}
        self.device = getattr(model, "device", torch.device("cuda:0"))

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def search(self, corpus, queries, top_k: int, score_function: str, return_sorted: bool = False, **kwargs):
        from .index import _check_score_function

        _check_score_function(score_function)
        rank, world = self._world()
        query_ids = list(queries.keys())
        self.results = {qid: {} for qid in query_ids}
        q_emb = torch.as_tensor(self.model.encode_queries(
            [(qid, queries[qid]) for qid in query_ids], batch_size=self.batch_size,
            show_progress_bar=self.show_progress_bar, convert_to_tensor=self.convert_to_tensor)).to(self.device).float()
        corpus_ids = sorted(corpus, key=lambda k: len(corpus[k].get("title", "") + corpus[k].get("text", "")),
                            reverse=True)
        row_of = {cid: i for i, cid in enumerate(corpus_ids)}
        exclude = torch.tensor([row_of.get(qid, -1) for qid in query_ids], dtype=torch.int64, device=self.device)
        mine = corpus_ids[rank::world]  # global sorted row of local row j is rank + j * world
        if self.shard_factory is None:
            from .index import CorpusShard

            shard = CorpusShard(q_emb.shape[1], max(len(mine), 1), device=self.device)
        else:
            shard = self.shard_factory(q_emb.shape[1], max(len(mine), 1))
        for batch_num, start in enumerate(range(0, len(mine), self.corpus_chunk_size)):
            chunk = [(cid, corpus[cid]) for cid in mine[start:start + self.corpus_chunk_size]]
            shard.add(torch.as_tensor(self.model.encode_corpus(
                chunk, batch_size=self.batch_size, show_progress_bar=self.show_progress_bar,
                convert_to_tensor=self.convert_to_tensor, batch_num=f"{rank}_{batch_num}")).to(self.device).float())
        kk = top_k + 1
        s, i = shard.search(q_emb, kk, score_function)
        i = torch.where(i >= 0, i * world + rank, i)  # local row -> row in the global length-sorted order
        if world > 1:
            gs, gi = all_gather_topk(s, i, self.group)
        else:
            gs, gi = s.unsqueeze(0), i.unsqueeze(0)
        merge = self.merge
        if merge is None:
            from .index import merge_topk as merge
        s, i = merge(gs, gi, exclude)
        for qi, (srow, irow) in enumerate(zip(s.cpu().tolist(), i.cpu().tolist())):
            res = self.results[query_ids[qi]]
            for score, row in zip(srow, irow):
                if row >= 0:
                    res[corpus_ids[row]] = score
        return self.results
