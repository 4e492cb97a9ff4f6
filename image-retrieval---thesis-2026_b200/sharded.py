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

from .search import FlatIndex


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


def candidate_views(buf: torch.Tensor, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The (distances fp32 [Q,k], indices int64 [Q,k]) views of one packed candidate buffer of 3*Q*k int32 words
    (indices first: 8-byte aligned) -- the search writes its result straight into the exchange buffer."""
    n = nq * k
    return buf[2 * n:3 * n].view(torch.float32).view(nq, k), buf[:2 * n].view(torch.int64).view(nq, k)


class PeerExchange:
    """Candidate exchange over NVLink peer memory instead of an all-gather: every rank's search writes its ``[Q,k]``
    candidates into ITS symmetric-memory buffer, publishes the number of the search in every rank's flag array, and
    every rank's merge kernel (``knn_merge_topk_parts_sync``) waits for all flags and then reads all W buffers directly
    through the peer mappings -- signal, wait and merge are two launches, no barrier kernel (the library barrier of
    torch's symmetric memory cost ~1 ms per search on 8 GPUs).  Two alternating slots: a slot is rewritten two searches
    later, and a peer publishes search e+1 only behind its own merge of search e, so when this rank's merge e+1 has seen
    every flag, every peer has finished reading slot e."""

    def __init__(self, group, device: torch.device, slots: int = 2):
        self.group = group if group is not None else dist.group.WORLD
        self.device = device
        self.slots = int(slots)  # 2: synchronous exchange; 4: pipelined exchange (see ShardedFlatIndex.search_async)
        self.capacity = 0        # int32 words per slot
        self.buf = None
        self.hdl = None
        self.turn = 0
        self.epoch = 0
        self.flags = None
        self.flag_hdl = None
        self._retired = []       # outgrown buffers stay mapped: a peer's merge may still be reading them

    def _ensure(self, words: int) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        if self.flags is None:
            self.flags = symm_mem.empty((64,), dtype=torch.int32, device=self.device)
            self.flags.zero_()
            self.flag_hdl = symm_mem.rendezvous(self.flags, self.group)
            self.flag_hdl.barrier(channel=0)   # once: every rank's flags are zero before anybody publishes
        if words <= self.capacity:
            return
        # the capacity must be identical on every rank: it only depends on (Q, k), which are
        cap = max(words, 1 << 16)
        if self.buf is not None:
            self._retired.append((self.buf, self.hdl))
        self.buf = symm_mem.empty((self.slots * cap,), dtype=torch.int32, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.capacity = cap
        self.turn = 0

    def slot(self, nq: int, k: int) -> torch.Tensor:
        """This rank's buffer for the next search (3*Q*k words of the current slot)."""
        words = 3 * nq * k
        self._ensure(words + (words & 1))
        self.turn = (self.turn + 1) % self.slots
        self.epoch += 1
        return self.buf[self.turn * self.capacity: self.turn * self.capacity + words]

    def peer_pointers(self, nq: int, k: int):
        """-> the W (val, idx) addresses of the current slot and the W flag arrays (all as mapped into this process)."""
        n = nq * k
        base = [int(p) + self.turn * self.capacity * 4 for p in self.hdl.buffer_ptrs]
        return [b + 8 * n for b in base], base, [int(p) for p in self.flag_hdl.buffer_ptrs]


class PendingSearch:
    """What :meth:`ShardedFlatIndex.search_async` returns: the result tensors and the event after which they are
    complete (recorded on the side stream of the pipelined exchange; ``None`` = complete in stream order already)."""

    def __init__(self, vals: torch.Tensor, idx: torch.Tensor, done: Optional[torch.cuda.Event] = None,
                 stream: Optional[torch.cuda.Stream] = None):
        self.vals, self.idx, self.done, self.stream = vals, idx, done, stream

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(distances, indices), safe to use on the CURRENT stream (which is made to wait for the merge)."""
        if self.done is not None:
            cur = torch.cuda.current_stream(self.vals.device)
            cur.wait_event(self.done)
            self.vals.record_stream(cur)
            self.idx.record_stream(cur)
            self.done = None
        return self.vals, self.idx

    def to_host(self, out_vals: torch.Tensor, out_idx: torch.Tensor) -> torch.cuda.Event:
        """Copy the result into (pinned) host tensors WITHOUT involving the current stream: the copies are enqueued on
        the stream that produces the result -- the side stream of the pipelined exchange, right behind the merge -- so
        they run on the copy engine under the next local search instead of between two searches.  Returns the event
        after which the host tensors are complete (``event.synchronize()``, or ``stream.wait_event(event)``, before
        reading them)."""
        dev = self.vals.device
        s = self.stream if (self.done is not None and self.stream is not None) else torch.cuda.current_stream(dev)
        with torch.cuda.stream(s):
            out_vals.copy_(self.vals, non_blocking=True)
            out_idx.copy_(self.idx, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s)
        return ev


class ShardedFlatIndex:
    """One process per GPU; this rank's shard is a :class:`FlatIndex` whose ``index_base`` is its first global row.

    exchange: "allgather" = one NCCL all-gather of the packed candidates, merged in place (no unpacking copy);
    "peer" = symmetric-memory buffers read over NVLink by the merge kernel itself (:class:`PeerExchange`)."""

    pipeline = False   # class defaults (subclasses in the tests build instances without __init__)
    _side = None
    _merged: list = []

    def __init__(self, local: FlatIndex, group: Optional[dist.ProcessGroup] = None, exchange: str = "allgather",
                 pipeline: bool = False):
        if exchange not in ("allgather", "peer"):
            raise ValueError("exchange must be 'allgather' or 'peer'")
        self.local = local
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.exchange = exchange
        # pipeline: the peer exchange of search i (wait for the slowest shard + merge over NVLink) runs on a side stream
        # under the local search of i + 1 (search_async); plain search() then simply waits for its own result
        self.pipeline = bool(pipeline) and exchange == "peer"
        self._peer: Optional[PeerExchange] = None
        self._prof = None
        self._side = None            # side stream of the pipelined exchange
        self._merged = []            # events "merge j done" of the last searches (pipelined exchange)

    @classmethod
    def from_full(cls, gallery: torch.Tensor, metric: str = "cosine", precision: str = "fp32", *,
                  normalize: bool = False, group: Optional[dist.ProcessGroup] = None,
                  exchange: str = "allgather", pipeline: bool = False) -> "ShardedFlatIndex":
        """Every rank passes the same full gallery tensor (or at least its own rows); keeps only its range."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        start, count = shard_rows(gallery.shape[0], world)[rank]
        idx = FlatIndex(gallery.shape[1], metric, precision, normalize=normalize, index_base=start,
                        device=gallery.device)
        idx.add(gallery[start:start + count])
        return cls(idx, group, exchange, pipeline)

    # --- the two device steps; tests replace them to exercise the exchange logic without a GPU ---------
    def _search_local(self, queries, k, self_mode, query_offset, out=None, prepared=None):
        if prepared is not None:
            return self.local.search_prepared(prepared[0], prepared[1], k, self_mode=self_mode,
                                              query_offset=query_offset, out=out)
        return self.local.search(queries, k, self_mode=self_mode, query_offset=query_offset, out=out)

    def _merge_parts(self, bufs, nq, k):
        """bufs: one packed candidate buffer (3*Q*k int32 words) per shard, in rank order."""
        from .search import merge_topk_parts

        n = nq * k
        base = [b.data_ptr() for b in bufs]
        return merge_topk_parts([p + 8 * n for p in base], base, nq, k, self.local.metric, bufs[0].device)

    # --- optional timing of the steps of a search (CUDA events on the search stream, read back by the caller) ---
    def profile(self, on: bool = True) -> None:
        """Record events around the local search / the exchange barrier (or all-gather) / the shard merge of every
        following search; :meth:`profile_read` returns the per-call durations."""
        self._prof = [] if on else None

    def profile_read(self):
        """-> list of (local_ms, gather_ms, merge_ms) per profiled search.  Peer exchange: ``merge_ms`` holds the wait
        for the slowest shard + the merge over NVLink (one kernel), ``gather_ms`` is ~0; all-gather exchange:
        ``gather_ms`` is the NCCL all-gather."""
        out = []
        for ev in self._prof or []:
            ev[3].synchronize()
            out.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
        return out

    def _mark(self, marks):
        if self._prof is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)

    def search(self, queries: torch.Tensor, k: int, *, exclude_self: bool = False, self_mode: Optional[str] = None,
               query_offset: int = 0, broadcast_queries: bool = False, prepared=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (distances [Q,k], indices [Q,k] global rows), identical on every rank and for every shard count.
        ``prepared`` = (rows, squared norms) from ``FlatIndex.prepare_queries``: skips the query-side normalise + cast."""
        if broadcast_queries and self.world_size > 1:
            dist.broadcast(queries, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        mode = self_mode or ("exclude" if exclude_self else "keep")
        if self.world_size == 1:
            return self._search_local(queries, k, mode, query_offset, prepared=prepared)
        if self.pipeline:
            return self.search_async(queries, k, self_mode=mode, query_offset=query_offset, prepared=prepared).result()
        nq = (prepared[0] if prepared is not None else queries).shape[0]
        dev = (prepared[0] if prepared is not None else queries).device
        words = 3 * nq * k
        marks = []
        self._mark(marks)
        if self.exchange == "peer":
            if self._peer is None:
                self._peer = PeerExchange(self.group, self.local.device)
            send = self._peer.slot(nq, k)
            self._search_local(queries, k, mode, query_offset, out=candidate_views(send, nq, k), prepared=prepared)
            from .search import merge_topk_parts

            self._mark(marks)
            val_ptrs, idx_ptrs, flag_ptrs = self._peer.peer_pointers(nq, k)
            self._mark(marks)
            # publish this search's number, wait for every peer's, merge over NVLink: inside the library call
            res = merge_topk_parts(val_ptrs, idx_ptrs, nq, k, self.local.metric, self.local.device,
                                   flag_ptrs=flag_ptrs, rank=self.rank, epoch=self._peer.epoch)
        else:
            # the search writes straight into the send buffer; the gathered buffer is merged where it lies
            send = torch.empty((words + (words & 1),), dtype=torch.int32, device=dev)
            self._search_local(queries, k, mode, query_offset, out=candidate_views(send, nq, k), prepared=prepared)
            self._mark(marks)
            gathered = torch.empty((self.world_size * send.numel(),), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gathered, send, group=self.group)
            self._mark(marks)
            res = self._merge_parts(list(gathered.view(self.world_size, send.numel())), nq, k)
        self._mark(marks)
        if self._prof is not None:
            self._prof.append(marks)
            if len(self._prof) > 64:
                self._prof.pop(0)
        return res

    def search_async(self, queries: torch.Tensor, k: int, *, exclude_self: bool = False,
                     self_mode: Optional[str] = None, query_offset: int = 0, prepared=None) -> PendingSearch:
        """:meth:`search` without waiting for the exchange: with ``pipeline=True`` (peer exchange) the local search and
        the publish of its candidates run on the current stream, the wait for the slowest shard + the merge over NVLink on
        a high-priority side stream -- the caller may enqueue the next search at once and take ``.result()`` later.
        Per-step jitter between the ranks and the merge itself then hide under the next local search.

        Buffer protocol (four candidate slots): search j is enqueued behind this rank's OWN merge j-2.  When that merge
        has seen every peer's flag j-2, every peer has started search j-2, hence finished its merge j-4 -- the last reader
        of the slot search j rewrites."""
        mode = self_mode or ("exclude" if exclude_self else "keep")
        if self.world_size == 1 or not self.pipeline:
            v, i = self.search(queries, k, self_mode=mode, query_offset=query_offset, prepared=prepared)
            return PendingSearch(v, i, None)
        import ctypes

        from . import _lib as L
        from .search import merge_topk_parts

        nq = (prepared[0] if prepared is not None else queries).shape[0]
        dev = self.local.device
        if self._peer is None:
            self._peer = PeerExchange(self.group, dev, slots=4)
            # high priority: the merge takes SMs as CTAs of the next search retire.  Measured on 8 GPUs (C5): 37.24 ms per
            # step against 37.46 with the searches on a high-priority stream and the merge on the lowest one (the merge
            # then runs at the very end of the next search), 37.78 without the pipeline
            self._side = torch.cuda.Stream(dev, priority=-1)
        main = torch.cuda.current_stream(dev)
        marks = []
        self._mark(marks)
        if len(self._merged) >= 2:
            main.wait_event(self._merged[-2])
        send = self._peer.slot(nq, k)
        self._search_local(queries, k, mode, query_offset, out=candidate_views(send, nq, k), prepared=prepared)
        val_ptrs, idx_ptrs, flag_ptrs = self._peer.peer_pointers(nq, k)
        epoch = self._peer.epoch
        fp = (ctypes.c_void_p * self.world_size)(*flag_ptrs)
        with torch.cuda.device(dev):
            L.check(L.load().knn_peer_publish(fp, self.world_size, self.rank, epoch, main.cuda_stream), "knn_peer_publish")
        self._mark(marks)
        searched = torch.cuda.Event()
        searched.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(searched)
            self._mark(marks)
            vals, idx = merge_topk_parts(val_ptrs, idx_ptrs, nq, k, self.local.metric, dev, flag_ptrs=flag_ptrs,
                                         rank=self.rank, epoch=epoch, publish=False)
            self._mark(marks)
            done = torch.cuda.Event()
            done.record(self._side)
        self._merged = (self._merged + [done])[-4:]
        if self._prof is not None:
            self._prof.append(marks)
            if len(self._prof) > 64:
                self._prof.pop(0)
        return PendingSearch(vals, idx, done, self._side)

    def search_host_async(self, queries_host: torch.Tensor, k: int, *, exclude_self: bool = False,
                          self_mode: Optional[str] = None, query_offset: int = 0) -> PendingSearch:
        """:meth:`search_host` without waiting for the exchange (see :meth:`search_async`)."""
        if self.world_size == 1:
            v, i = self.search_host(queries_host, k, exclude_self=exclude_self, self_mode=self_mode,
                                    query_offset=query_offset)
            return PendingSearch(v, i, None)
        return self.search_async(None, k, exclude_self=exclude_self, self_mode=self_mode, query_offset=query_offset,
                                 prepared=self._prepare_host(queries_host))

    def search_host(self, queries_host: torch.Tensor, k: int, *, exclude_self: bool = False,
                    self_mode: Optional[str] = None, query_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """:meth:`search` for a query batch that lives in HOST memory on every rank (the same batch everywhere): each
        rank copies only ITS 1/W slice to the device, normalises + casts it to the search dtype, and the prepared slices
        are all-gathered over NVLink (bf16: Q*D*2 bytes in total) -- instead of W host-to-device copies of the whole fp32
        batch and W redundant normalisations."""
        if self.world_size == 1:
            return self.search(queries_host.to(self.local.device, non_blocking=True), k, exclude_self=exclude_self,
                               self_mode=self_mode, query_offset=query_offset)
        return self.search(None, k, exclude_self=exclude_self, self_mode=self_mode, query_offset=query_offset,
                           prepared=self._prepare_host(queries_host))

    def _prepare_host(self, queries_host: torch.Tensor):
        """This rank's 1/W slice of a host batch copied, normalised + cast; the prepared slices all-gathered."""
        nq = queries_host.shape[0]
        dev = self.local.device
        per = (nq + self.world_size - 1) // self.world_size          # equal slices (the last ones may be short / empty)
        s, e = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
        mine = queries_host[s:e].to(dev, non_blocking=True)
        want_sq = self.local.metric == "l2"
        if e > s:
            q, qsq = self.local.prepare_queries(mine)
        else:
            q, qsq = None, None
        dtype = torch.bfloat16 if self.local.precision == "bf16" else torch.float32
        dpad = self.local.rows.shape[1]
        send = torch.zeros((per, dpad), dtype=dtype, device=dev)
        if q is not None:
            send[: e - s] = q
        full = torch.empty((self.world_size * per, dpad), dtype=dtype, device=dev)
        dist.all_gather_into_tensor(full, send, group=self.group)
        full_sq = None
        if want_sq:
            sq_send = torch.zeros((per,), dtype=torch.float32, device=dev)
            if qsq is not None:
                sq_send[: e - s] = qsq
            full_sq = torch.empty((self.world_size * per,), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(full_sq, sq_send, group=self.group)
            full_sq = full_sq[:nq]
        return full[:nq], full_sq
