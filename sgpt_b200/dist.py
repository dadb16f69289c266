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
