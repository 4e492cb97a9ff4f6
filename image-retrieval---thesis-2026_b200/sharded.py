"""Row-sharded gallery over the GPUs of one node.

The reference all-gathers EMBEDDINGS and recomputes the whole N x N evaluation on every rank's CPU
(train.py:604-609).  Here every rank keeps a contiguous row range ``[r*N/W, (r+1)*N/W)`` of the gallery
resident in its own HBM, queries are replicated (optionally broadcast from rank 0), each rank runs the
fused distance + top-k over its shard with GLOBAL row indices, and ONE all-gather of the ``[Q, k]``
candidates (8 B index + 4 B score each) followed by an on-device k-way merge yields a result that is
independent of the shard count.  The data path has no other collective.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .search import FlatIndex, merge_topk


def shard_rows(n_total: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges ``(start, count)`` per rank: ``[r*N//W, (r+1)*N//W)``."""
    out = []
    for r in range(world_size):
        s, e = (r * n_total) // world_size, ((r + 1) * n_total) // world_size
        out.append((s, e - s))
    return out


def pack_candidates(vals: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """[Q,k] fp32 + [Q,k] int64 -> one int32 buffer of 3*Q*k words (single all-gather payload)."""
    n = vals.numel()
    buf = torch.empty((3 * n,), dtype=torch.int32, device=vals.device)
    buf[: 2 * n].view(torch.int64).copy_(idx.reshape(-1))
    buf[2 * n:].view(torch.float32).copy_(vals.reshape(-1))
    return buf


def unpack_candidates(buf: torch.Tensor, parts: int, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    n = nq * k
    buf = buf.view(parts, 3 * n)
    idx = buf[:, : 2 * n].contiguous().view(torch.int64).view(parts, nq, k)
    vals = buf[:, 2 * n:].contiguous().view(torch.float32).view(parts, nq, k)
    return vals, idx


class ShardedFlatIndex:
    """One process per GPU; this rank's shard is a :class:`FlatIndex` whose ``index_base`` is its first global row."""

    def __init__(self, local: FlatIndex, group: Optional[dist.ProcessGroup] = None):
        self.local = local
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    @classmethod
    def from_full(cls, gallery: torch.Tensor, metric: str = "cosine", precision: str = "fp32", *,
                  normalize: bool = False, group: Optional[dist.ProcessGroup] = None) -> "ShardedFlatIndex":
        """Every rank passes the same full gallery tensor (or at least its own rows); keeps only its range."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        start, count = shard_rows(gallery.shape[0], world)[rank]
        idx = FlatIndex(gallery.shape[1], metric, precision, normalize=normalize, index_base=start,
                        device=gallery.device)
        idx.add(gallery[start:start + count])
        return cls(idx, group)

    # --- the two device steps; tests replace them to exercise the exchange logic without a GPU ---------
    def _search_local(self, queries, k, self_mode, query_offset):
        return self.local.search(queries, k, self_mode=self_mode, query_offset=query_offset)

    def _merge(self, vals, idx):
        return merge_topk(vals, idx, self.local.metric)

    def search(self, queries: torch.Tensor, k: int, *, exclude_self: bool = False, self_mode: Optional[str] = None,
               query_offset: int = 0, broadcast_queries: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (distances [Q,k], indices [Q,k] global rows), identical on every rank and for every shard count."""
        if broadcast_queries and self.world_size > 1:
            dist.broadcast(queries, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        mode = self_mode or ("exclude" if exclude_self else "keep")
        vals, idx = self._search_local(queries, k, mode, query_offset)
        if self.world_size == 1:
            return vals, idx
        nq = vals.shape[0]
        payload = pack_candidates(vals, idx)
        gathered = torch.empty((self.world_size * payload.numel(),), dtype=payload.dtype, device=payload.device)
        dist.all_gather_into_tensor(gathered, payload, group=self.group)
        pv, pi = unpack_candidates(gathered, self.world_size, nq, k)
        return self._merge(pv, pi)
