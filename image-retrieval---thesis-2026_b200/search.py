"""Exact k-NN search on B200: the host-side mirror of the reference's retrieval calls.

Call shapes follow what the reference's scripts consume:

* ``values, indices = S.topk(k, dim=1, largest=True, sorted=True)`` on ``S = q @ g.T`` / ``-cdist(q, g)``
  (test.py:44,1080; train.py:405-409; xai_conceptclip.py:468) -> :func:`search`
* ``scores, neighbors = index.search(xq, k)`` after ``index.add(xb)`` (ATH.py:403-410) -> :class:`FlatIndex`
* ``F.normalize(x, p=2, dim=1)`` (test.py:1005) / ``x / x.norm(dim=-1, keepdim=True)`` (test.py:251) /
  ``l2_normalize`` (fusion_eval/fuse.py:11-15) -> :func:`normalize`
* ``torch.argsort(dists, dim=1, descending=True)`` (test.py:1018) -> :func:`rank_rows`

Everything runs in the CUDA library behind include/b200knn.h; torch only owns the device buffers.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib as L

_METRICS = {"cosine": L.KNN_COSINE, "ip": L.KNN_IP, "l2": L.KNN_L2}
_EPS_MODES = {"clamp": L.KNN_EPS_CLAMP, "none": L.KNN_EPS_NONE, "add": L.KNN_EPS_ADD, "cast": L.KNN_CAST_ONLY}
_SELF = {"keep": L.KNN_SELF_KEEP, "exclude": L.KNN_SELF_EXCLUDE, "minus1": L.KNN_SELF_MINUS1}
_DT = {torch.float32: L.KNN_F32, torch.bfloat16: L.KNN_BF16}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise L.KnnError("b200knn operates on CUDA tensors only (there is no CPU fallback)")


def _as2d(x: torch.Tensor, name: str) -> torch.Tensor:
    if x.dim() != 2:
        raise ValueError(f"{name} must be [rows, dim], got {tuple(x.shape)}")
    return x.contiguous()


def normalize(
    x: torch.Tensor,
    *,
    eps: float = 1e-12,
    eps_mode: str = "clamp",
    out_dtype: Optional[torch.dtype] = None,
    return_sqnorm: bool = False,
):
    """Row-wise L2 normalisation fused with the cast to the search dtype.

    eps_mode: "clamp" = F.normalize (x / max(||x||, eps)); "none" = x / ||x|| (test.py:251);
    "add" = x / (||x|| + eps) (test.py:444); "cast" = no normalisation, cast only.
    """
    _require_cuda(x)
    x = _as2d(x, "x")
    if x.dtype not in _DT:
        x = x.float()
    out_dtype = out_dtype or x.dtype
    if out_dtype not in _DT:
        raise ValueError(f"unsupported out_dtype {out_dtype}")
    n, d = x.shape
    y = torch.empty((n, d), dtype=out_dtype, device=x.device)
    sq = torch.empty((n,), dtype=torch.float32, device=x.device) if return_sqnorm else None
    with torch.cuda.device(x.device):
        rc = L.load().knn_normalize(_ptr(x), _ptr(y), _ptr(sq), n, d, _DT[x.dtype], _DT[out_dtype], float(eps),
                                    _EPS_MODES[eps_mode], _stream(x))
    L.check(rc, "knn_normalize")
    return (y, sq) if return_sqnorm else y


def row_sqnorm(x: torch.Tensor) -> torch.Tensor:
    """fp32 squared L2 norm of every stored row (the |g|^2 term of cdist's GEMM form)."""
    _require_cuda(x)
    x = _as2d(x, "x")
    sq = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.load().knn_row_sqnorm(_ptr(x), _ptr(sq), x.shape[0], x.shape[1], _DT[x.dtype], _stream(x))
    L.check(rc, "knn_row_sqnorm")
    return sq


def _pad_dim(x: torch.Tensor, mult: int) -> torch.Tensor:
    d = x.shape[1]
    if d % mult == 0:
        return x
    pad = mult - d % mult
    out = x.new_zeros((x.shape[0], d + pad))  # zero columns change no dot product and no norm
    out[:, :d] = x
    return out


def _prepare(x: torch.Tensor, do_norm: bool, precision: str, eps: float, eps_mode: str, want_sq: bool):
    """-> (rows in the search dtype, squared norms or None)."""
    tgt = torch.bfloat16 if precision == "bf16" else torch.float32
    x = _as2d(x, "embeddings")
    if x.dtype not in _DT:
        x = x.float()
    if precision == "bf16":
        x = _pad_dim(x, 8)
    if do_norm:
        if want_sq:
            return normalize(x, eps=eps, eps_mode=eps_mode, out_dtype=tgt, return_sqnorm=True)
        return normalize(x, eps=eps, eps_mode=eps_mode, out_dtype=tgt), None
    if x.dtype != tgt:
        x = normalize(x, eps_mode="cast", out_dtype=tgt)
    return x, (row_sqnorm(x) if want_sq else None)


class FlatIndex:
    """Device-resident exact (FLAT) gallery: the drop-in for ``faiss.IndexFlatL2`` + ``add`` (ATH.py:401-403)
    and for a Milvus FLAT collection (ChestMIR/milvus_embed.py:538-539).

    Rows are stored once in the search dtype (fp32 exact mode or bf16 tensor-core mode), optionally
    L2-normalised by the fused normalise+cast kernel, with their squared norms for the L2 metric.
    ``index_base`` is the global row of local row 0 (row-sharded galleries).
    """

    def __init__(self, dim: int, metric: str = "cosine", precision: str = "fp32", *, normalize: bool = False,
                 eps: float = 1e-12, eps_mode: str = "clamp", index_base: int = 0,
                 device: Optional[torch.device] = None):
        if metric not in _METRICS:
            raise ValueError(f"metric must be one of {sorted(_METRICS)}")
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.dim, self.metric, self.precision = int(dim), metric, precision
        self.normalize, self.eps, self.eps_mode = bool(normalize), float(eps), eps_mode
        self.index_base = int(index_base)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.rows: Optional[torch.Tensor] = None
        self.sqnorm: Optional[torch.Tensor] = None
        self._filter: Optional[ExactFilterRows] = None   # bf16 split of fp32 rows (tensor-core exact engine), lazy
        self._packed: Optional["PackedRows"] = None      # fp32 rows in the FFMA kernel's packed operand layout, lazy

    @property
    def ntotal(self) -> int:
        return 0 if self.rows is None else int(self.rows.shape[0])

    def add(self, x: torch.Tensor) -> "FlatIndex":
        """Append rows (``index.add(xb)``, ATH.py:403; ``collection.insert``, nih_zilliz_utils.py:244-251)."""
        _require_cuda(x)
        if x.shape[1] != self.dim:
            raise ValueError(f"expected dim {self.dim}, got {x.shape[1]}")
        rows, sq = _prepare(x.to(self.device), self.normalize, self.precision, self.eps, self.eps_mode,
                            self.metric == "l2")
        self.rows = rows if self.rows is None else torch.cat([self.rows, rows], 0)
        if sq is not None:
            self.sqnorm = sq if self.sqnorm is None else torch.cat([self.sqnorm, sq], 0)
        self._filter = None
        self._packed = None
        return self

    def adopt(self, rows: torch.Tensor, sqnorm: Optional[torch.Tensor] = None) -> "FlatIndex":
        """Take ownership of rows that are ALREADY in the search dtype/layout (no copy)."""
        _require_cuda(rows)
        want = torch.bfloat16 if self.precision == "bf16" else torch.float32
        if rows.dtype != want or not rows.is_contiguous():
            raise ValueError("adopt() needs contiguous rows already in the search dtype")
        self.rows = rows
        self.sqnorm = sqnorm if sqnorm is not None else (row_sqnorm(rows) if self.metric == "l2" else None)
        self._filter = None
        self._packed = None
        return self

    def search(self, queries: torch.Tensor, k: int, *, exclude_self: bool = False, self_mode: Optional[str] = None,
               query_offset: int = 0, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``scores, neighbors = index.search(xq, k)`` (ATH.py:410) -> (distances [Q,k] fp32, indices [Q,k] int64).

        query_offset: global gallery row of query 0 (self-retrieval over a chunk of the gallery).
        out: optional (float32 [Q,k], int64 [Q,k]) tensors to write the result into (exchange buffers).
        """
        if self.rows is None:
            raise L.KnnError("FlatIndex is empty")
        _require_cuda(queries)
        q, qsq = self.prepare_queries(queries)
        return self.search_prepared(q, qsq, k, exclude_self=exclude_self, self_mode=self_mode,
                                    query_offset=query_offset, out=out)

    def prepare_queries(self, queries: torch.Tensor):
        """The query-side half of :meth:`search`: rows in the search dtype (normalised when the index normalises) and
        their squared norms (L2 metric only).  Row-wise, so a batch may be prepared in slices (sharded.search_host)."""
        return _prepare(queries.to(self.device), self.normalize, self.precision, self.eps, self.eps_mode,
                        self.metric == "l2")

    def search_prepared(self, q: torch.Tensor, qsq: Optional[torch.Tensor], k: int, *, exclude_self: bool = False,
                        self_mode: Optional[str] = None, query_offset: int = 0, out=None):
        """:meth:`search` for queries that went through :meth:`prepare_queries` already."""
        if self.rows is None:
            raise L.KnnError("FlatIndex is empty")
        mode = self_mode or ("exclude" if exclude_self else "keep")
        if self.precision == "fp32" and exact_engine(q.shape[0], self.ntotal, self.dim, int(k), self.device,
                                                     self._filter is not None) == "tensor":
            if self._filter is None:
                self._filter = ExactFilterRows.build(self.rows, self.sqnorm)
            return _search_exact_tensor(q, qsq, self.rows, self.sqnorm, int(k), self.metric, mode, int(query_offset),
                                        self.index_base, self._filter, out=out)
        if (self.precision == "fp32" and self._packed is None and int(k) <= L.MAX_FUSED_K
                and use_packed(q.shape[0], self.ntotal, self.dim, self.device)):
            self._packed = PackedRows.build(self.rows)
        return _search_prepared(q, qsq, self.rows, self.sqnorm, int(k), self.metric, mode, int(query_offset),
                                self.index_base, out=out, g_packed=self._packed)


# ---------------------------------------------------------------------------------------------------------------
# Exact fp32 search on the tensor cores (csrc/exact_tc.cu): bf16x3 split filter -> exact fp32 re-scoring of kc > k
# candidates per query -> per-query proof that nothing outside the candidate set can be in the answer -> FFMA
# re-run of the (rare) queries the proof does not cover.  The result is bit-identical to the FFMA engine's.
# ---------------------------------------------------------------------------------------------------------------
_EXACT_MIN_FLOP = 4.0e10   # below ~1 ms of FFMA work the detour through the tensor cores does not pay


def _filter_k(k: int) -> int:
    """Candidates per query the filter returns: k plus a slack of max(8, k/8), rounded up to the list geometry."""
    want = k + max(8, k // 8)
    kc = 32
    while kc < want:
        kc <<= 1
    return kc


def exact_engine(nq: int, ng: int, d: int, k: int, device=None, have_split: bool = False) -> str:
    """Which kernel family serves precision="fp32": "ffma" (search_f32.cu) or "tensor" (filter + re-score).
    KNN_EXACT_ENGINE=ffma|tensor forces one (tests, measurements).  The tensor engine keeps a bf16 split of the gallery
    (6 bytes per element next to the 4 of the fp32 rows): it is only chosen when that fits the device's free memory
    (``device`` given and the split not built yet)."""
    forced = os.environ.get("KNN_EXACT_ENGINE", "")
    feasible = k <= L.MAX_FUSED_K and _filter_k(k) <= L.MAX_FUSED_K and nq > 0 and ng > 0
    if forced == "ffma" or not feasible:
        return "ffma"
    if forced != "tensor" and 2.0 * nq * ng * d < _EXACT_MIN_FLOP:
        return "ffma"
    if device is not None and not have_split and torch.cuda.is_available():
        need = 6 * ng * ((d + 7) // 8 * 8) + 6 * nq * d + (1 << 28)
        # cudaMemGetInfo costs ~1 ms: only asked when the split is a sizeable part of the device memory
        if need > torch.cuda.get_device_properties(device).total_memory // 16:
            free, _ = torch.cuda.mem_get_info(device)
            if need >= 0.9 * free:
                return "ffma"
    return "tensor"


def rerun_ranges(bad_rows, nq: int, block: int = 128):
    """Query ranges to re-run through the FFMA engine for the unproven rows `bad_rows`: whole 128-row blocks merged
    into contiguous runs (a contiguous range keeps the ``query_offset + i`` self-row arithmetic; re-running proven
    rows is harmless -- both engines return the same bits); everything at once when half the blocks are affected."""
    if not bad_rows:
        return []
    blocks = sorted({int(r) // block for r in bad_rows})
    if 2 * len(blocks) >= (nq + block - 1) // block:
        return [(0, nq)]
    runs, s0, prev = [], blocks[0], blocks[0]
    for b in blocks[1:]:
        if b != prev + 1:
            runs.append((s0 * block, min(nq, (prev + 1) * block)))
            s0 = b
        prev = b
    runs.append((s0 * block, min(nq, (prev + 1) * block)))
    return runs


def split_bf16x3(x: torch.Tensor, role: str) -> torch.Tensor:
    """fp32 rows [n, d] -> bf16 [n, 3 * dpad]: ``[hi | lo | hi]`` for role "queries", ``[hi | hi | lo]`` for
    "gallery" (hi = bf16(x), lo = bf16(x - hi), dpad = d rounded up to 8)."""
    _require_cuda(x)
    x = _as2d(x, "x")
    if x.dtype != torch.float32:
        raise ValueError("split_bf16x3 takes fp32 rows")
    n, d = x.shape
    out = torch.empty((n, 3 * ((d + 7) // 8 * 8)), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.load().knn_split_bf16x3(_ptr(x), n, d, {"queries": 0, "gallery": 1}[role], _ptr(out), _stream(x))
    L.check(rc, "knn_split_bf16x3")
    return out


class PackedRows:
    """fp32 rows in the KNN_F32_PACKED operand layout of the exact FFMA kernel (``knn_pack_f32``: 128-row tiles
    ``[tile][dpad][128]``): pack a gallery once, search it with many query batches."""

    def __init__(self, data: torch.Tensor, n: int, d: int):
        self.data, self.n, self.d = data, n, d

    @staticmethod
    def build(rows: torch.Tensor) -> "PackedRows":
        _require_cuda(rows)
        rows = _as2d(rows, "rows")
        if rows.dtype != torch.float32 or not rows.is_contiguous():
            raise ValueError("PackedRows takes contiguous fp32 rows")
        n, d = rows.shape
        lib = L.load()
        out = torch.empty((max(lib.knn_pack_f32_bytes(n, d) // 4, 4),), dtype=torch.float32, device=rows.device)
        with torch.cuda.device(rows.device):
            rc = lib.knn_pack_f32(_ptr(rows), n, d, _ptr(out), _stream(rows))
        L.check(rc, "knn_pack_f32")
        return PackedRows(out, n, d)


_PACK_MIN_FLOP = 4.0e9     # below ~0.1 ms of FFMA work the two pack launches do not pay


def use_packed(nq: int, ng: int, d: int, device=None) -> bool:
    """Whether an exact-fp32 FFMA call packs its operands first (KNN_F32_PACK=0|1 forces it off / on): enough work to
    pay for the pack launches, and the packed copy (the size of the rows) fits the device's free memory."""
    forced = os.environ.get("KNN_F32_PACK", "")
    if forced in ("0", "1"):
        return forced == "1" and nq > 0 and ng > 0
    if 2.0 * nq * ng * d < _PACK_MIN_FLOP:
        return False
    if device is not None and torch.cuda.is_available():
        need = 4 * (ng + nq + 256) * ((d + 15) // 16 * 16) + (1 << 28)
        if need > torch.cuda.get_device_properties(device).total_memory // 16:
            free, _ = torch.cuda.mem_get_info(device)
            if need >= 0.5 * free:
                return False
    return True


def _packed_operands(q, g, g_packed):
    """(q pointer tensor, g pointer tensor, dtype code) of an fp32 FFMA call."""
    if g_packed is None and not use_packed(q.shape[0], g.shape[0], q.shape[1], q.device):
        return q, g, L.KNN_F32
    gp = g_packed if g_packed is not None else PackedRows.build(g)
    if gp.n != g.shape[0] or gp.d != g.shape[1]:
        raise ValueError("packed gallery does not match the gallery rows")
    qp = PackedRows.build(q if q.is_contiguous() else q.contiguous())
    return qp.data, gp.data, L.KNN_F32_PACKED


class ExactFilterRows:
    """What the tensor-core exact engine keeps per gallery besides the fp32 rows: the bf16 split rows, the squared
    norms and their maximum (a device scalar; enters the error bound of the filter)."""

    def __init__(self, split: torch.Tensor, sqnorm: torch.Tensor, max_sqnorm: torch.Tensor,
                 lo_max_sqnorm: Optional[torch.Tensor] = None):
        self.split, self.sqnorm, self.max_sqnorm = split, sqnorm, max_sqnorm
        self.lo_max_sqnorm = lo_max_sqnorm     # max |g - bf16(g)|^2 over the rows: the bound of the two-product filter

    @staticmethod
    def build(rows: torch.Tensor, sqnorm: Optional[torch.Tensor]) -> "ExactFilterRows":
        sq = sqnorm if sqnorm is not None else row_sqnorm(rows)
        mx = torch.empty((1,), dtype=torch.float32, device=rows.device)
        lo = torch.empty((1,), dtype=torch.float32, device=rows.device)
        split = split_bf16x3(rows, "gallery")
        with torch.cuda.device(rows.device):
            rc = L.load().knn_max_sqnorm(_ptr(sq), sq.numel(), _ptr(mx), _stream(rows))
            L.check(rc, "knn_max_sqnorm")
            rc = L.load().knn_split_lo_max_sqnorm(_ptr(split), rows.shape[0], rows.shape[1], _ptr(lo), _stream(rows))
            L.check(rc, "knn_split_lo_max_sqnorm")
        return ExactFilterRows(split, sq, mx, lo)


def filter_error_bound(qsq: torch.Tensor, max_gsq: torch.Tensor, d: int, metric: str,
                       lo_max_gsq: Optional[torch.Tensor] = None) -> torch.Tensor:
    """eps[q] >= |filter value - exact-mode value| for every gallery row (fp32 [Q], rounded up): split error +
    tensor-core accumulation + the exact fp32 chain's own rounding, times |q| * max|g| (``knn_filter_error_bound``,
    formula in include/b200knn.h).  With ``lo_max_gsq`` (max squared norm of the gallery's lo parts): the bound of the
    TWO-product filter, whose dropped product adds |q| * max|g_lo| (``knn_filter_error_bound2``)."""
    _require_cuda(qsq, max_gsq)
    eps = torch.empty_like(qsq)
    with torch.cuda.device(qsq.device):
        if lo_max_gsq is None:
            rc = L.load().knn_filter_error_bound(_ptr(qsq), qsq.numel(), _ptr(max_gsq), int(d), _METRICS[metric],
                                                 _ptr(eps), _stream(qsq))
        else:
            rc = L.load().knn_filter_error_bound2(_ptr(qsq), qsq.numel(), _ptr(max_gsq), _ptr(lo_max_gsq), int(d),
                                                  _METRICS[metric], _ptr(eps), _stream(qsq))
    L.check(rc, "knn_filter_error_bound")
    return eps


_TWO_PRODUCT_MIN_ROWS = 49152


def _two_product_filter(nq: int, ng: int, self_mode: str) -> bool:
    """Whether the tensor-core exact engine tries the two-product filter first (KNN_EXACT_PRODUCTS=2|3 forces it):
    queries it cannot prove are GATHERED and re-run through the three-product filter, which needs self_mode "keep" (a
    gathered batch has no ``query_offset + i`` self rows), several 256-row query blocks to be worth a second pass, and a
    gallery (shard) large enough: dropping a product saves a third of the filter, which scales with the gallery rows,
    while the wider candidate set costs ~1.1 ms of extra re-scoring per 25 000 queries whatever the gallery -- even at
    ~42 k rows (25 000 x 14 000 x 1024, the 8-GPU shard of config 3: 6.2 ms with two products, 4.9 ms with three)."""
    forced = os.environ.get("KNN_EXACT_PRODUCTS", "")
    if self_mode != "keep" or forced == "3":
        return False
    return forced == "2" or (nq >= 1024 and ng >= _TWO_PRODUCT_MIN_ROWS)


def _search_exact_tensor(q, qsq, g, gsq, k, metric, self_mode, query_offset, index_base,
                         filt: Optional[ExactFilterRows] = None, out=None, products: Optional[int] = None):
    """precision="fp32" through the tensor cores; same contract and same bits as _search_prepared on fp32 rows.
    products: 3 = the three-product split filter; 2 = the two-product filter first (2/3 of the tensor work, ~65 x the
    error bound), its unproven queries gathered and re-run with three products; None = by :func:`_two_product_filter`."""
    nq, d = q.shape
    ng = g.shape[0]
    dev = q.device
    kc = _filter_k(k)
    if filt is None:
        filt = ExactFilterRows.build(g, gsq)
    if products is None:
        products = 2 if (_two_product_filter(nq, ng, self_mode) and filt.lo_max_sqnorm is not None) else 3
    if products == 2:
        # the wide bound needs a wider candidate set: with kc = 64 for k = 50 the proof failed for 41 % of the queries
        # of BASELINE config 3 (the 14 ranks of slack span ~0.003 of score, the bound is ~0.002), with 128 for none;
        # knn_rescore_exact prunes the candidates that provably cannot reach the top k, so the re-scoring grows by
        # ~10 %, not 2 x
        kc2 = 32
        while kc2 < 2 * k + 16:
            kc2 <<= 1
        if kc2 > L.MAX_FUSED_K:
            products = 3
        else:
            kc = kc2
    q_sq = qsq if qsq is not None else row_sqnorm(q)
    eps = filter_error_bound(q_sq, filt.max_sqnorm, d, metric, filt.lo_max_sqnorm if products == 2 else None)
    lib = L.load()
    q3 = split_bf16x3(q, "queries")
    if products == 2:
        cand_val = torch.empty((nq, kc), dtype=torch.float32, device=dev)
        cand_idx = torch.empty((nq, kc), dtype=torch.int64, device=dev)
        d3 = q3.shape[1]
        with torch.cuda.device(dev):
            nbytes = lib.knn_search_workspace(nq, ng, d3, L.KNN_BF16X2, kc)
            ws = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=dev)
            rc = lib.knn_search(_ptr(q3), _ptr(filt.split), _ptr(qsq), _ptr(gsq if metric == "l2" else None), nq, ng, d3,
                                L.KNN_BF16X2, kc, _METRICS[metric], _SELF[self_mode], query_offset, index_base,
                                _ptr(cand_val), _ptr(cand_idx), _ptr(ws), ws.numel(), _stream(q))
        L.check(rc, "knn_search")
    else:
        cand_val, cand_idx = _search_prepared(q3, qsq, filt.split, gsq if metric == "l2" else None,
                                              kc, metric, self_mode, query_offset, index_base, split_rows=True)
    out_val, out_idx = _out_buffers(out, nq, k, dev)
    flags = torch.empty((nq,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.knn_rescore_exact(_ptr(q), _ptr(g), _ptr(qsq), _ptr(gsq), nq, ng, d, _METRICS[metric],
                                   _SELF[self_mode], query_offset, index_base, _ptr(cand_val), _ptr(cand_idx),
                                   kc, k, _ptr(eps), _ptr(out_val), _ptr(out_idx), _ptr(flags), _stream(q))
    L.check(rc, "knn_rescore_exact")
    bad = torch.nonzero(flags).flatten()
    if products == 2:
        # unproven under the wide bound: gather those queries and run them through the three-product filter (which
        # falls back to the FFMA engine for what IT cannot prove)
        _search_exact_tensor.last_two_product_rerun = int(bad.numel())
        if bad.numel() > 0:
            qb = q.index_select(0, bad)
            v, i = _search_exact_tensor(qb, None if qsq is None else qsq.index_select(0, bad), g, gsq, k, metric,
                                        self_mode, 0, index_base, filt=filt, products=3)
            out_val.index_copy_(0, bad, v)
            out_idx.index_copy_(0, bad, i)
        else:
            _search_exact_tensor.last_unverified = 0
        return out_val, out_idx
    # queries whose candidate set could not be proven complete (ties / near-duplicates wider than the slack):
    # re-run their 128-row blocks through the FFMA engine (contiguous runs keep the self-row arithmetic)
    runs = rerun_ranges(bad.tolist(), nq)
    gp = PackedRows.build(g) if len(runs) > 1 and use_packed(runs[0][1] - runs[0][0], ng, d, dev) else None
    for s, e in runs:
        v, i = _search_prepared(q[s:e], None if qsq is None else qsq[s:e], g, gsq, k, metric, self_mode,
                                query_offset + s, index_base, g_packed=gp)
        out_val[s:e] = v
        out_idx[s:e] = i
    _search_exact_tensor.last_unverified = int(bad.numel())
    return out_val, out_idx


_search_exact_tensor.last_two_product_rerun = 0
_search_exact_tensor.last_unverified = 0


def _out_buffers(out, nq, k, dev):
    """(distances, indices) to write into: fresh tensors, or the caller's (a send / symmetric-memory buffer)."""
    if out is None:
        return (torch.empty((nq, k), dtype=torch.float32, device=dev),
                torch.empty((nq, k), dtype=torch.int64, device=dev))
    ov, oi = out
    if (ov.shape != (nq, k) or oi.shape != (nq, k) or ov.dtype != torch.float32 or oi.dtype != torch.int64
            or not ov.is_contiguous() or not oi.is_contiguous() or ov.device != dev or oi.device != dev):
        raise ValueError("out must be contiguous (float32 [Q,k], int64 [Q,k]) tensors on the search device")
    return ov, oi


def _search_prepared(q, qsq, g, gsq, k, metric, self_mode, query_offset, index_base, split_rows=False, out=None,
                     g_packed: Optional["PackedRows"] = None):
    """split_rows: q / g are split_bf16x3 rows (KNN_BF16X3: same scores, each part loaded once per tile).
    g_packed: the fp32 gallery already in the packed operand layout (else packed here when the call is large enough)."""
    nq, d = q.shape
    ng = g.shape[0]
    if g.shape[1] != d or g.dtype != q.dtype:
        raise ValueError("queries and gallery must share dim and dtype")
    if k < 1:
        raise ValueError("k must be >= 1")
    dev = q.device
    out_val, out_idx = _out_buffers(out, nq, k, dev)
    if nq == 0:
        return out_val, out_idx
    lib = L.load()
    with torch.cuda.device(dev):
        if k <= L.MAX_FUSED_K:
            dt = L.KNN_BF16X3 if split_rows else _DT[q.dtype]
            qd, gd = q, g
            if dt == L.KNN_F32:
                qd, gd, dt = _packed_operands(q, g, g_packed)
            nbytes = lib.knn_search_workspace(nq, ng, d, dt, k)
            ws = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=dev)
            rc = lib.knn_search(_ptr(qd), _ptr(gd), _ptr(qsq), _ptr(gsq), nq, ng, d, dt, k, _METRICS[metric],
                                _SELF[self_mode], query_offset, index_base, _ptr(out_val), _ptr(out_idx),
                                _ptr(ws), ws.numel(), _stream(q))
            L.check(rc, "knn_search")
            return out_val, out_idx
    # k beyond the fused limit (train.py:409 asks for topk(N-1)): dense scores + full row ranking, in query
    # chunks so the score block stays bounded.  Still the CUDA library -- never torch.mm / torch.sort.
    kk = min(k, ng)
    chunk = max(1, min(nq, (1 << 28) // max(ng, 1)))
    largest = metric != "l2"
    out_val.fill_(float("-inf") if largest else float("inf"))
    out_idx.fill_(-1)
    qf = q if q.dtype == torch.float32 else normalize(q, eps_mode="cast", out_dtype=torch.float32)
    gf = g if g.dtype == torch.float32 else normalize(g, eps_mode="cast", out_dtype=torch.float32)
    if g_packed is None and use_packed(min(nq, chunk), ng, d, dev):
        g_packed = PackedRows.build(gf)
    for s in range(0, nq, chunk):
        e = min(nq, s + chunk)
        sc = _scores_dense_prepared(qf[s:e], None if qsq is None else qsq[s:e], gf, gsq, metric, self_mode,
                                    query_offset - index_base + s, g_packed=g_packed)
        rk = rank_rows(sc, largest_first=largest)[:, :kk]
        out_idx[s:e, :kk] = rk + index_base
        out_val[s:e, :kk] = torch.gather(sc, 1, rk)
    if self_mode == "exclude":  # the masked self entry ranks last with -inf/+inf: report it as "no candidate"
        bad = torch.isinf(out_val) & (out_idx >= 0)
        out_idx[bad] = -1
    return out_val, out_idx


def _scores_dense_prepared(q, qsq, g, gsq, metric, self_mode, self_offset_local, split_rows=False,
                           g_packed: Optional["PackedRows"] = None):
    """split_rows: q / g are split_bf16x3 rows (KNN_BF16X3); bf16 rows run the tcgen05 kernel, fp32 rows the exact one
    (g_packed: the gallery already in the packed operand layout)."""
    nq, d = q.shape
    ng = g.shape[0]
    out = torch.empty((nq, ng), dtype=torch.float32, device=q.device)
    dt = L.KNN_BF16X3 if split_rows else _DT[q.dtype]
    if dt == L.KNN_F32 and nq > 0 and ng > 0:
        q, g, dt = _packed_operands(q, g, g_packed)
    with torch.cuda.device(q.device):
        rc = L.load().knn_scores_dense(_ptr(q), _ptr(g), _ptr(qsq), _ptr(gsq), nq, ng, d, dt,
                                       _METRICS[metric], _SELF[self_mode], self_offset_local, _ptr(out), _stream(q))
    L.check(rc, "knn_scores_dense")
    return out


def search(
    queries: torch.Tensor,
    gallery,
    k: int,
    metric: str = "cosine",
    *,
    normalize: bool = False,
    exclude_self: bool = False,
    self_mode: Optional[str] = None,
    query_offset: int = 0,
    precision: str = "fp32",
    negated: bool = False,
    eps: float = 1e-12,
    eps_mode: str = "clamp",
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused distance + top-k: ``(distances [Q,k] fp32, indices [Q,k] int64)``, best first, ties by ascending
    gallery row.  cosine/ip: similarities, descending.  l2: Euclidean distances, ascending
    (``torch.argsort(cdist(q, g), dim=1)[:, :k]``, test_ath.py:87-114); ``negated=True`` returns ``-distance``
    like ``(-torch.cdist(e, e)).topk(k)`` (test.py:1080-1085).

    ``gallery`` is a ``[N, D]`` tensor or a :class:`FlatIndex` (then metric/precision/normalize come from it).
    ``exclude_self`` masks gallery row ``query_offset + i`` for query ``i`` (``fill_diagonal_(-inf)``,
    test.py:1081); ``self_mode="minus1"`` keeps it with score -1 (nih_multilabel_training.py:86).
    """
    if isinstance(gallery, FlatIndex):
        vals, idx = gallery.search(queries, k, exclude_self=exclude_self, self_mode=self_mode,
                                   query_offset=query_offset)
        metric = gallery.metric
    else:
        _require_cuda(queries, gallery)
        if metric not in _METRICS:
            raise ValueError(f"metric must be one of {sorted(_METRICS)}")
        want_sq = metric == "l2"
        q, qsq = _prepare(queries, normalize, precision, eps, eps_mode, want_sq)
        if gallery is queries:
            g, gsq = q, qsq
        else:
            g, gsq = _prepare(gallery, normalize, precision, eps, eps_mode, want_sq)
        mode = self_mode or ("exclude" if exclude_self else "keep")
        if precision == "fp32" and exact_engine(q.shape[0], g.shape[0], q.shape[1], int(k), q.device) == "tensor":
            vals, idx = _search_exact_tensor(q, qsq, g, gsq, int(k), metric, mode, int(query_offset), 0)
        else:
            vals, idx = _search_prepared(q, qsq, g, gsq, int(k), metric, mode, int(query_offset), 0)
    if negated and metric == "l2":
        vals = -vals
    return vals, idx


def scores_dense(queries: torch.Tensor, gallery: torch.Tensor, metric: str = "cosine", *, normalize: bool = False,
                 self_mode: str = "keep", query_offset: int = 0, eps: float = 1e-12,
                 eps_mode: str = "clamp", precision: str = "fp32") -> torch.Tensor:
    """The dense ``dists`` matrix of test.py:1080 / fusion_eval/metrics.py:15 (callers that save or post-process the
    full matrix).  Similarities for cosine/ip, positive distances for l2.  precision: "fp32" = the exact fp32 chain
    (bit-equal to the search), "bf16" = bf16 rows on the tensor cores, "bf16x3" = fp32 rows through the error-free
    three-product split on the tensor cores (|error| <~ 1e-5 |q||g|)."""
    _require_cuda(queries, gallery)
    if precision not in ("fp32", "bf16", "bf16x3"):
        raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
    want_sq = metric == "l2"
    prep = "bf16" if precision == "bf16" else "fp32"
    q, qsq = _prepare(queries, normalize, prep, eps, eps_mode, want_sq)
    g, gsq = (q, qsq) if gallery is queries else _prepare(gallery, normalize, prep, eps, eps_mode, want_sq)
    if precision == "bf16x3":
        q3 = split_bf16x3(q, "queries")
        g3 = split_bf16x3(g, "gallery")
        return _scores_dense_prepared(q3, qsq, g3, gsq, metric, self_mode, int(query_offset), split_rows=True)
    return _scores_dense_prepared(q, qsq, g, gsq, metric, self_mode, int(query_offset))


def pack_bits(codes: torch.Tensor) -> torch.Tensor:
    """Pack 0/1 codes ``[N, bits]`` (any float/int dtype; a position is set iff non-zero) into ``[N, ceil(bits/64)]``
    int64 words for the Hamming search (``knn_pack_bits``)."""
    _require_cuda(codes)
    x = _as2d(codes, "codes").float().contiguous()
    n, bits = x.shape
    words = (bits + 63) // 64
    out = torch.empty((n, words), dtype=torch.int64, device=x.device)
    if n == 0:
        return out
    with torch.cuda.device(x.device):
        rc = L.load().knn_pack_bits(_ptr(x), n, bits, L.KNN_F32, _ptr(out), _stream(x))
    L.check(rc, "knn_pack_bits")
    return out


def unpack_bits_pm1(words: torch.Tensor, bits: int) -> torch.Tensor:
    """Packed codes ``[N, ceil(bits/64)]`` int64 -> ``[N, dpad]`` bf16 rows of +1 / -1 (``knn_unpack_bits_pm1``)."""
    _require_cuda(words)
    words = words.contiguous()
    n = words.shape[0]
    out = torch.empty((n, (bits + 7) // 8 * 8), dtype=torch.bfloat16, device=words.device)
    if n:
        with torch.cuda.device(words.device):
            rc = L.load().knn_unpack_bits_pm1(_ptr(words), n, int(bits), _ptr(out), _stream(words))
        L.check(rc, "knn_unpack_bits_pm1")
    return out


def search_hamming(query_codes: torch.Tensor, gallery_codes: torch.Tensor, k: int, *, exclude_self: bool = False,
                   query_offset: int = 0, packed: bool = False, bits: Optional[int] = None,
                   method: str = "auto") -> Tuple[torch.Tensor, torch.Tensor]:
    """Hamming ranking of binary codes: the fused form of ``(q[:, None, :] != g[None, :, :]).sum(2).float()`` +
    ``argsort(dim=1)[:, :k]`` (test_ath.py:80-100) -> (distances [Q,k] fp32 ascending, indices [Q,k] int64), ties by
    ascending gallery row.  Codes are 0/1 tensors ``[N, bits]`` (as ``(codes >= 0).float()``, test_ath.py:68), or
    already packed words with ``packed=True`` (then pass ``bits``; default 64 * words).

    method: "popc" = xor + popcount over the packed words (``knn_search_hamming``; 8 bytes per 64 bits of gallery);
    "mma" = the codes as +-1 bf16 rows through the tcgen05 kernels (``<q, g> = bits - 2 d``, exact); "auto" = mma
    whenever the expanded rows fit the device's free memory (measured faster at every batch size), else popc."""
    _require_cuda(query_codes, gallery_codes)
    if method not in ("auto", "popc", "mma"):
        raise ValueError("method must be 'auto', 'popc' or 'mma'")
    qw = query_codes.contiguous() if packed else pack_bits(query_codes)
    gw = gallery_codes.contiguous() if packed else (qw if gallery_codes is query_codes else pack_bits(gallery_codes))
    if qw.dtype != torch.int64 or gw.dtype != torch.int64 or qw.shape[1] != gw.shape[1]:
        raise ValueError("packed codes must be int64 words with the same number of words per row")
    nq, words = qw.shape
    ng = gw.shape[0]
    nbits = int(bits) if bits is not None else (words * 64 if packed else int(query_codes.shape[1]))
    if not (words - 1) * 64 < nbits <= words * 64:
        raise ValueError(f"bits={nbits} does not match {words} words per code")
    if k < 1:
        raise ValueError("k must be >= 1")
    dev = qw.device
    if method == "auto":   # the +-1 expansion (2 bytes per code bit) must fit comfortably; it wins at every batch size
        need = 2 * (ng + nq) * ((nbits + 7) // 8 * 8)
        fits = need < torch.cuda.get_device_properties(dev).total_memory // 16 or need < 0.5 * torch.cuda.mem_get_info(dev)[0]
        method = "mma" if ng > 0 and nq > 0 and fits else "popc"
        # the popcount kernel is instantiated for 1, 2, 3, 4 or 8 words and the fused list sizes: anything else (the
        # reference ranks the whole matrix for any code length / topk, test_ath.py:100) goes through the +-1 rows,
        # whose search has the dense + full-ranking path for k beyond the fused limit
        if ng > 0 and nq > 0 and (k > L.MAX_FUSED_K or words not in (1, 2, 3, 4, 8)):
            method = "mma"
    if method == "popc" and k > L.MAX_FUSED_K:
        raise L.KnnError(f"the popcount Hamming kernel supports k <= {L.MAX_FUSED_K}; use method='mma' (or 'auto')")
    if method == "mma":
        q1 = unpack_bits_pm1(qw, nbits)
        g1 = q1 if gw is qw else unpack_bits_pm1(gw, nbits)
        score, idx = _search_prepared(q1, None, g1, None, int(k), "ip", "exclude" if exclude_self else "keep",
                                      int(query_offset), 0)
        if k > ng:   # the large-k path pads with (-inf, -1) like the fused one
            score, idx = score[:, :k], idx[:, :k]
        # d = (bits - score) / 2: exact small integers; empty slots (score -inf) -> +inf like the popcount path
        dist_out = torch.empty_like(score)
        with torch.cuda.device(dev):
            rc = L.load().knn_hamming_from_scores(_ptr(score), score.numel(), nbits, _ptr(dist_out), _stream(score))
        L.check(rc, "knn_hamming_from_scores")
        return dist_out, idx
    if words not in (1, 2, 3, 4, 8):
        raise L.KnnError(f"the popcount kernel is instantiated for 1, 2, 3, 4 or 8 words per code, got {words}")
    out_val = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    if nq == 0:
        return out_val, out_idx
    lib = L.load()
    with torch.cuda.device(dev):
        nbytes = lib.knn_search_hamming_workspace(nq, ng, k)
        ws = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=dev)
        rc = lib.knn_search_hamming(_ptr(qw), _ptr(gw), nq, ng, words, k, _SELF["exclude" if exclude_self else "keep"],
                                    int(query_offset), 0, _ptr(out_val), _ptr(out_idx), _ptr(ws), ws.numel(), _stream(qw))
    L.check(rc, "knn_search_hamming")
    return out_val, out_idx


def rank_rows(scores: torch.Tensor, largest_first: bool = True) -> torch.Tensor:
    """Stable full ranking of every row: best first, ties by ascending column (the deterministic form of
    ``torch.argsort(dists, dim=1, descending=True)``, test.py:1018).  -> int64 [Q, N]."""
    _require_cuda(scores)
    scores = _as2d(scores.float(), "scores")
    nq, ng = scores.shape
    ranks = torch.empty((nq, ng), dtype=torch.int64, device=scores.device)
    if nq == 0 or ng == 0:
        return ranks
    lib = L.load()
    with torch.cuda.device(scores.device):
        nbytes = lib.knn_rank_rows_workspace(nq, ng)
        ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=scores.device)
        rc = lib.knn_rank_rows(_ptr(scores), nq, ng, 1 if largest_first else 0, _ptr(ranks), _ptr(ws), ws.numel(),
                               _stream(scores))
    L.check(rc, "knn_rank_rows")
    return ranks


def merge_topk_parts(val_ptrs, idx_ptrs, nq: int, k: int, metric: str, device, flag_ptrs=None, rank: int = 0,
                     epoch: int = 0, publish: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """k-way merge of separately placed per-shard lists given as raw DEVICE addresses (``knn_merge_topk_parts``):
    slices of an all-gathered buffer, or the peers' symmetric-memory buffers (fused exchange + merge).  With
    ``flag_ptrs`` (every rank's flag array as mapped here) the synchronisation of the exchange runs inside
    (``knn_merge_topk_parts_sync``): this rank publishes ``epoch``, the merge waits for every peer's;
    ``publish=False`` only waits and merges (``knn_merge_topk_parts_wait``: the publish ran elsewhere)."""
    import ctypes

    parts = len(val_ptrs)
    out_val = torch.empty((nq, k), dtype=torch.float32, device=device)
    out_idx = torch.empty((nq, k), dtype=torch.int64, device=device)
    vp = (ctypes.c_void_p * parts)(*[int(p) for p in val_ptrs])
    ip = (ctypes.c_void_p * parts)(*[int(p) for p in idx_ptrs])
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        if flag_ptrs is None:
            rc = L.load().knn_merge_topk_parts(vp, ip, parts, nq, k, _METRICS[metric], _ptr(out_val), _ptr(out_idx), stream)
        else:
            fp = (ctypes.c_void_p * parts)(*[int(p) for p in flag_ptrs])
            fn = L.load().knn_merge_topk_parts_sync if publish else L.load().knn_merge_topk_parts_wait
            rc = fn(vp, ip, parts, nq, k, _METRICS[metric], fp, int(rank), int(epoch), _ptr(out_val), _ptr(out_idx), stream)
    L.check(rc, "knn_merge_topk_parts")
    return out_val, out_idx


def merge_topk(vals: torch.Tensor, idx: torch.Tensor, metric: str = "cosine") -> Tuple[torch.Tensor, torch.Tensor]:
    """k-way merge of per-shard results ``[parts, Q, k]`` -> ``[Q, k]`` (after the candidate all-gather)."""
    _require_cuda(vals, idx)
    vals = vals.contiguous().float()
    idx = idx.contiguous().long()
    parts, nq, k = vals.shape
    out_val = torch.empty((nq, k), dtype=torch.float32, device=vals.device)
    out_idx = torch.empty((nq, k), dtype=torch.int64, device=vals.device)
    with torch.cuda.device(vals.device):
        rc = L.load().knn_merge_topk(_ptr(vals), _ptr(idx), parts, nq, k, _METRICS[metric], _ptr(out_val),
                                     _ptr(out_idx), _stream(vals))
    L.check(rc, "knn_merge_topk")
    return out_val, out_idx
