"""Full-ranking retrieval metrics without the N x N ranking.

The reference's full-ranking metrics -- ``compute_map`` (test.py:95-146), ``compute_map_multilabel`` (test.py:941-985),
``_compute_single_label_retrieval_metrics`` / ``_compute_multilabel_retrieval_metrics`` (train.py:399-487),
``evaluate_map`` (nih_multilabel_training.py:66-99), ``evaluate_retrieval_metrics`` (fusion_eval/metrics.py:41-94) -- all
sort every score row (``argsort`` / ``topk(N-1)``) and then only look at WHERE the relevant rows ended up.  At the NIH
scale (112 k rows) the ranking matrix alone would be 100 GB of indices.  Here the query set is walked in chunks:

    dense exact-fp32 score block [chunk, N]      knn_scores_dense        (the same score definition as knn_search)
    -> ranks of the relevant rows per query      knn_rank_of_positives   (csrc/rank_positives.cu)
    -> per-query AP variants                     knn_ap_from_ranks / knn_ap_sklearn_from_ranks

Nothing of size N x N is ever allocated; per chunk the transient memory is a few [chunk, N] 32-bit buffers.  The ranks
equal the positions a full ``rank_rows`` ranking gives (best score first, ties by ascending gallery row), so every metric
is bit-identical to the dense path it replaces.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import _lib as L
from .search import (PackedRows, _METRICS, _SELF, _prepare, _ptr, _require_cuda, _scores_dense_prepared, _stream,
                     split_bf16x3, use_packed)

REL_SINGLE, REL_JACCARD_F32, REL_JACCARD_F64, REL_ANY = 0, 1, 2, 3
_CHUNK_BYTES = 8 << 30   # transient memory budget of the query chunks in flight


def rank_of_positives(scores: torch.Tensor, rel_mode: int, q_rel: torch.Tensor, g_rel: torch.Tensor, *,
                      largest_first: bool = True, jaccard_threshold: float = 0.0, self_offset: int = 0,
                      drop_self: bool = False, q_group: Optional[torch.Tensor] = None,
                      g_group: Optional[torch.Tensor] = None, ties: bool = False,
                      ld_out: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """``knn_rank_of_positives`` on a dense score block ``[Q, N]`` fp32 (row q = query q against the whole gallery).
    q_rel / g_rel: int64 labels (REL_SINGLE) or int64 bit masks of the multi-hot labels (``metrics.pack_multihot``).
    q_group / g_group (int64, optional): gallery rows whose group equals the query's are not ranked (rows that share
    the query's image path, fusion_eval/metrics.py:67).
    -> {"pos_ranks" int32 [Q, ld], "npos" int32 [Q], "nranked" int32 [Q]} (+ "pos_ge", "pos_tgroup", "ngroups" with
    ``ties``): row q holds the ascending 0-based ranks of its npos[q] relevant gallery rows."""
    _require_cuda(scores, q_rel, g_rel)
    if scores.dim() != 2 or scores.dtype != torch.float32 or scores.stride(1) != 1:
        raise ValueError("scores must be a [Q, N] float32 block with contiguous rows")
    nq, ng = scores.shape
    dev = scores.device
    q_rel = q_rel.to(device=dev, dtype=torch.int64).contiguous().view(-1)
    g_rel = g_rel.to(device=dev, dtype=torch.int64).contiguous().view(-1)
    if q_rel.numel() != nq or g_rel.numel() != ng:
        raise ValueError("q_rel / g_rel must hold one label (or label mask) per query / gallery row")
    if (q_group is None) != (g_group is None):
        raise ValueError("q_group and g_group come together")
    if q_group is not None:
        q_group = q_group.to(device=dev, dtype=torch.int64).contiguous().view(-1)
        g_group = g_group.to(device=dev, dtype=torch.int64).contiguous().view(-1)
    ld = int(ld_out) if ld_out is not None else max(ng, 1)
    out = {"pos_ranks": torch.empty((nq, ld), dtype=torch.int32, device=dev),
           "npos": torch.empty((nq,), dtype=torch.int32, device=dev),
           "nranked": torch.empty((nq,), dtype=torch.int32, device=dev)}
    if ties:
        out["pos_ge"] = torch.empty((nq, ld), dtype=torch.int32, device=dev)
        out["pos_tgroup"] = torch.empty((nq, ld), dtype=torch.int32, device=dev)
        out["ngroups"] = torch.empty((nq,), dtype=torch.int32, device=dev)
    if nq == 0 or ng == 0:
        for key in ("npos", "nranked", "ngroups"):
            if key in out:
                out[key].zero_()
        return out
    lib = L.load()
    with torch.cuda.device(dev):
        nbytes = lib.knn_rank_of_positives_workspace(nq, ng)
        ws = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=dev)
        rc = lib.knn_rank_of_positives(_ptr(scores), scores.stride(0), nq, ng, 1 if largest_first else 0, int(rel_mode),
                                       _ptr(q_rel), _ptr(g_rel), float(jaccard_threshold), int(self_offset),
                                       1 if drop_self else 0, _ptr(q_group), _ptr(g_group), _ptr(out["pos_ranks"]), ld,
                                       _ptr(out.get("pos_ge")),
                                       _ptr(out.get("pos_tgroup")), _ptr(out["npos"]), _ptr(out["nranked"]),
                                       _ptr(out.get("ngroups")), _ptr(ws), ws.numel(), _stream(scores))
    L.check(rc, "knn_rank_of_positives")
    return out


_AP_OUTPUTS = ("ap_trapz", "prs", "nres", "prec_sum", "first", "hits_at")


def ap_from_ranks(rp: Dict[str, torch.Tensor], kappas: Sequence[int] = (), self_last_positive: bool = False,
                  outputs: Optional[Sequence[str]] = None):
    """``knn_ap_from_ranks``: -> {"ap_trapz" f64 [Q] (compute_ap, test.py:58-92), "prs" f64 [Q, len(kappas)]
    (test.py:137-140), "nres" int32 [Q], "prec_sum" f64 [Q], "first" int32 [Q] (1-based, 0 = none),
    "hits_at" int32 [Q, len(kappas)]}; ``outputs`` selects a subset (the two AP sums are chains of dependent double
    additions over all positives: ask only for the one a metric needs)."""
    pr = rp["pos_ranks"]
    nq, ld = pr.shape
    dev = pr.device
    nk = len(kappas)
    if nk > 8:
        raise ValueError("at most 8 cut-offs per call")
    kap = torch.as_tensor([int(k) for k in kappas], dtype=torch.int32, device=dev)
    want = set(_AP_OUTPUTS if outputs is None else outputs)
    if not want <= set(_AP_OUTPUTS):
        raise ValueError(f"outputs must be among {_AP_OUTPUTS}")
    shapes = {"ap_trapz": ((nq,), torch.float64), "prs": ((nq, max(nk, 1)), torch.float64), "nres": ((nq,), torch.int32),
              "prec_sum": ((nq,), torch.float64), "first": ((nq,), torch.int32),
              "hits_at": ((nq, max(nk, 1)), torch.int32)}
    out = {k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=dev) for k in _AP_OUTPUTS if k in want}
    if not out:
        return out
    with torch.cuda.device(dev):
        rc = L.load().knn_ap_from_ranks(_ptr(pr), ld, _ptr(rp["npos"]), _ptr(rp["nranked"]), nq,
                                        1 if self_last_positive else 0, _ptr(kap), nk, _ptr(out.get("ap_trapz")),
                                        _ptr(out.get("prs")), _ptr(out.get("nres")), _ptr(out.get("prec_sum")),
                                        _ptr(out.get("first")), _ptr(out.get("hits_at")), _stream(pr))
    L.check(rc, "knn_ap_from_ranks")
    for k in ("prs", "hits_at"):
        if k in out:
            out[k] = out[k][:, :nk]
    return out


def ap_sklearn_from_ranks(rp: Dict[str, torch.Tensor]) -> torch.Tensor:
    """``knn_ap_sklearn_from_ranks``: sklearn's average_precision_score per query over the full ranking (tied scores
    grouped; NaN = no relevant row).  Needs the tie outputs of :func:`rank_of_positives`."""
    ge = rp["pos_ge"]
    nq, ld = ge.shape
    dev = ge.device
    ap = torch.empty((nq,), dtype=torch.float64, device=dev)
    lib = L.load()
    with torch.cuda.device(dev):
        nbytes = lib.knn_ap_sklearn_from_ranks_workspace(nq, ld)
        ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=dev)
        rc = lib.knn_ap_sklearn_from_ranks(_ptr(ge), _ptr(rp["pos_tgroup"]), ld, _ptr(rp["npos"]), _ptr(rp["ngroups"]),
                                           nq, _ptr(ap), _ptr(ws), ws.numel(), _stream(ge))
    L.check(rc, "knn_ap_sklearn_from_ranks")
    return ap


def chunk_rows(ng: int, ties: bool, budget: int = _CHUNK_BYTES, device=None) -> int:
    """Queries per chunk so that the [chunk, N] transients (scores + ranks, + tie outputs and the sklearn workspace) of
    the two chunks in flight stay inside ``budget`` bytes; a multiple of 128 (whole query blocks of the distance kernel --
    the rank kernel claims rows dynamically and does not care)."""
    per_row = 2 * ng * (4 + 4 + (8 + 12 if ties else 0))
    rows = int(max(1, min(8192, budget // max(per_row, 1))))
    return rows - rows % 128 if rows >= 128 else rows


def query_slice(nq: int, world: int, rank: int):
    """Contiguous query range ``[s, e)`` of a rank (the same split as ``sharded.shard_rows``)."""
    return (rank * nq) // world, ((rank + 1) * nq) // world


def gather_query_sharded(local: Dict[str, torch.Tensor], nq: int, group=None) -> Dict[str, torch.Tensor]:
    """All-gather per-query tensors that every rank computed for ITS query slice (``query_slice``) into the full
    ``[Q, ...]`` tensors, identical on every rank and in the original query order.  Slices are padded to the longest one
    for the collective (one all-gather per output)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [query_slice(nq, world, r)[1] - query_slice(nq, world, r)[0] for r in range(world)]
    longest = max(sizes)
    out = {}
    for key, val in local.items():
        pad = torch.zeros((longest,) + tuple(val.shape[1:]), dtype=val.dtype, device=val.device)
        pad[: val.shape[0]] = val
        full = torch.empty((world * longest,) + tuple(val.shape[1:]), dtype=val.dtype, device=val.device)
        dist.all_gather_into_tensor(full, pad.contiguous(), group=group)
        out[key] = torch.cat([full[r * longest: r * longest + sizes[r]] for r in range(world)], 0)
    return out


def full_ranking_stats(queries: torch.Tensor, gallery: torch.Tensor, rel_mode: int, q_rel: torch.Tensor,
                       g_rel: torch.Tensor, *, metric: str = "cosine", normalize: bool = False,
                       self_mode: str = "keep", drop_self: bool = False, query_offset: int = 0,
                       q_group: Optional[torch.Tensor] = None, g_group: Optional[torch.Tensor] = None,
                       jaccard_threshold: float = 0.0, kappas: Sequence[int] = (), sklearn_ap: bool = False,
                       self_last_positive: bool = False, outputs: Optional[Sequence[str]] = None,
                       eps: float = 1e-12, eps_mode: str = "clamp",
                       rows_per_chunk: Optional[int] = None, distributed: bool = False,
                       group=None, precision: str = "fp32") -> Dict[str, torch.Tensor]:
    """Per-query full-ranking statistics of ``queries`` against the whole ``gallery`` (exact fp32 scores), chunked over
    the queries.  ``self_mode`` is the score the query's own gallery row gets (``fill_diagonal_``: "exclude" = -inf,
    "minus1" = -1, "keep"); ``drop_self`` removes that row from the ranking and from the relevant set.
    -> the outputs of :func:`ap_from_ranks` (all, or the subset ``outputs``; + "ap_sklearn" with ``sklearn_ap``) and
    "npos", each ``[Q]`` / ``[Q, k]``.

    The AP kernels of a chunk are chains of dependent double additions in rank order (latency bound, a fraction of the
    SMs): they run on a side stream while the distance and rank kernels of the next chunk fill the machine.

    ``distributed=True`` (inside an initialised ``torch.distributed`` job, every rank holding the same inputs): the
    QUERIES are sharded -- rank r ranks the rows of ``query_slice(Q, W, r)`` against the whole gallery, the per-query
    statistics are all-gathered (a few bytes per query) and every rank returns the full, identical result.  The
    reference all-gathers the EMBEDDINGS and recomputes everything on every rank (train.py:604-609).

    ``precision``: "fp32" (default) ranks the exact fp32 scores -- bit-equal to the dense path and the CPU restatement in the tests, bound by
    the FP32 pipe; "bf16x3" computes the score block on the tensor cores from the error-free bf16 split of the fp32
    rows (|score error| <~ 1e-5 |q||g|: only scores that close may swap ranks); "bf16" ranks bf16-rounded rows."""
    _require_cuda(queries, gallery)
    if precision not in ("fp32", "bf16", "bf16x3"):
        raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
    if metric not in _METRICS:
        raise ValueError(f"metric must be one of {sorted(_METRICS)}")
    if self_mode not in _SELF:
        raise ValueError(f"self_mode must be one of {sorted(_SELF)}")
    want_sq = metric == "l2"
    prep = "bf16" if precision == "bf16" else "fp32"
    q, qsq = _prepare(queries, normalize, prep, eps, eps_mode, want_sq)
    g, gsq = (q, qsq) if gallery is queries else _prepare(gallery, normalize, prep, eps, eps_mode, want_sq)
    nq, ng = q.shape[0], g.shape[0]
    if nq == 0:
        raise ValueError("no queries")
    if distributed:
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            distributed = False
    if distributed:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        s0, e0 = query_slice(nq, world, rank)
        local = full_ranking_stats(q[s0:e0], g, rel_mode, q_rel[s0:e0], g_rel, metric=metric, normalize=False,
                                   self_mode=self_mode, drop_self=drop_self, query_offset=query_offset + s0,
                                   q_group=None if q_group is None else q_group[s0:e0], g_group=g_group,
                                   jaccard_threshold=jaccard_threshold, kappas=kappas, sklearn_ap=sklearn_ap,
                                   self_last_positive=self_last_positive, outputs=outputs,
                                   rows_per_chunk=rows_per_chunk, precision=precision) if e0 > s0 else None
        if local is None:   # a rank without queries still takes part in the collectives: build empty outputs
            probe = full_ranking_stats(q[:1], g, rel_mode, q_rel[:1], g_rel, metric=metric, self_mode=self_mode,
                                       drop_self=drop_self, query_offset=query_offset, jaccard_threshold=jaccard_threshold,
                                       kappas=kappas, sklearn_ap=sklearn_ap, self_last_positive=self_last_positive,
                                       outputs=outputs, q_group=None if q_group is None else q_group[:1], g_group=g_group,
                                       precision=precision)
            local = {k: v[:0] for k, v in probe.items()}
        return gather_query_sharded(local, nq, group)
    dev = q.device
    g3 = split_bf16x3(g, "gallery") if precision == "bf16x3" else None   # once; the query chunks are split on the fly
    step = int(rows_per_chunk) if rows_per_chunk else chunk_rows(ng, sklearn_ap, device=dev)
    # exact mode: the gallery goes into the FFMA kernel's packed operand layout once for all query chunks
    gp = PackedRows.build(g) if g3 is None and g.dtype == torch.float32 and use_packed(min(nq, step), ng, g.shape[1], dev) \
        else None
    main = torch.cuda.current_stream(dev)
    nchunks = (nq + step - 1) // step
    side = torch.cuda.Stream(dev) if nchunks > 1 else main
    parts: Dict[str, list] = {}
    pending = []   # (rank outputs, event recorded after the chunk's AP kernels): at most two chunks in flight
    for s in range(0, nq, step):
        e = min(nq, s + step)
        if len(pending) >= 2:   # the buffers of chunk i-2 are free once its AP kernels have run
            old_rp, old_done = pending.pop(0)
            main.wait_event(old_done)
            del old_rp
        if g3 is not None:
            sc = _scores_dense_prepared(split_bf16x3(q[s:e], "queries"), None if qsq is None else qsq[s:e], g3, gsq,
                                        metric, self_mode, query_offset + s, split_rows=True)
        else:
            sc = _scores_dense_prepared(q[s:e], None if qsq is None else qsq[s:e], g, gsq, metric, self_mode,
                                        query_offset + s, g_packed=gp)
        rp = rank_of_positives(sc, rel_mode, q_rel[s:e], g_rel, largest_first=(metric != "l2"),
                               jaccard_threshold=jaccard_threshold, self_offset=query_offset + s, drop_self=drop_self,
                               q_group=None if q_group is None else q_group[s:e], g_group=g_group, ties=sklearn_ap)
        del sc
        ranked = torch.cuda.Event()
        ranked.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ranked)
            st = ap_from_ranks(rp, kappas, self_last_positive, outputs)
            st["npos"] = rp["npos"]
            if sklearn_ap:
                st["ap_sklearn"] = ap_sklearn_from_ranks(rp)
            done = torch.cuda.Event()
            done.record(side)
        if side is not main:
            for t in rp.values():
                t.record_stream(side)
        pending.append((rp, done))
        for key, val in st.items():
            parts.setdefault(key, []).append(val)
    main.wait_stream(side)
    out = {key: torch.cat(vals, 0) for key, vals in parts.items()}
    return out
