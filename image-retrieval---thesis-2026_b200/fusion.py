"""Late fusion and re-ranking on top of the fused search.

Mirrors, without ever building an N x N similarity matrix:

* ``fusion_eval/fuse.py``      -- ``l2_normalize`` (:11), ``concat_fusion`` (:18), ``weighted_sum_fusion`` (:35)
* ``fusion_eval/evaluate.py``  -- score-level fusion ``alpha*S_conv + (1-alpha)*S_dino`` (:64-84) with the row-wise
  ``zscore`` / ``minmax`` normalisation of ``normalize_similarity_matrix`` (:152-177), the confidence fusion with a
  per-query alpha from top1-top2 margins (:180-214), and the experiment loop ``run_late_fusion_experiments`` (:30-149)
* ``test.py:612-621, 769-777`` -- re-scoring of the top ``rerank_k`` hits with an (item, query-class) text-score table
* ``retrieval_analysis/rerank.py`` -- the ``Reranker`` protocol and ``IdentityReranker``

How the N x N matrices disappear: a weighted sum of two similarities is ONE inner product of concatenated,
per-query-weighted embeddings,  a_q * (q1 . g1) + b_q * (q2 . g2) = [a_q q1, b_q q2] . [g1, g2],  so score fusion is a
single fused top-k search over the concatenated gallery; the row statistics the normalisations need (mean, std, min,
max of every similarity row) come from ``knn_score_stats`` (a statistics epilogue of the exact-fp32 distance kernel);
the per-query additive constants of zscore / minmax do not change a row's ranking and are added to the returned
values afterwards.  torch is used for embedding plumbing only (concatenation, per-row scaling of the query copies).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Protocol, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .search import _as2d, _ptr, _require_cuda, _stream, normalize, search


# --------------------------------------------------------------------------------------------------------------
# embedding-level fusion (fusion_eval/fuse.py)
# --------------------------------------------------------------------------------------------------------------
def _cuda_f32(x) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))
    t = t.float()
    return (t if t.is_cuda else t.cuda()).contiguous()


def l2_normalize(embeddings, eps: float = 1e-12) -> torch.Tensor:
    """fuse.py:11-15: ``x / maximum(||x||, eps)`` row-wise (fused normalise kernel)."""
    return normalize(_cuda_f32(embeddings), eps=eps, eps_mode="clamp")


def concat_fusion(conv_embeddings, dino_embeddings) -> torch.Tensor:
    """fuse.py:18-23: normalise, concatenate, normalise again."""
    return l2_normalize(torch.cat([l2_normalize(conv_embeddings), l2_normalize(dino_embeddings)], dim=1))


@dataclass(frozen=True)
class WeightedSumResult:
    """fuse.py:26-31."""

    embeddings: Optional[torch.Tensor]
    skipped_reason: Optional[str] = None


def weighted_sum_fusion(conv_embeddings, dino_embeddings, alpha: float) -> WeightedSumResult:
    """fuse.py:35-52: ``normalize(alpha*normalize(conv) + (1-alpha)*normalize(dino))`` when the dims match."""
    conv, dino = _cuda_f32(conv_embeddings), _cuda_f32(dino_embeddings)
    if conv.shape[1] != dino.shape[1]:
        return WeightedSumResult(None, "weighted_sum_skipped_dimension_mismatch:"
                                       f" conv_dim={conv.shape[1]}, dino_dim={dino.shape[1]}")
    fused = alpha * l2_normalize(conv) + (1.0 - alpha) * l2_normalize(dino)
    return WeightedSumResult(l2_normalize(fused))


# --------------------------------------------------------------------------------------------------------------
# similarity-row statistics (normalize_similarity_matrix, evaluate.py:152-177)
# --------------------------------------------------------------------------------------------------------------
def score_stats(queries: torch.Tensor, gallery: torch.Tensor, metric: str = "ip", self_mode: str = "keep",
                query_offset: int = 0) -> Dict[str, torch.Tensor]:
    """Row statistics of ``score(q, g)`` over the whole gallery: ``{"mean","std","min","max"}`` (float64, [Q]).
    ``std`` is the population standard deviation (``np.std``).  fp32 inputs, exact-fp32 scores."""
    from .search import _METRICS, _SELF, _packed_operands, row_sqnorm

    _require_cuda(queries, gallery)
    q, g = _as2d(queries.float().contiguous(), "queries"), _as2d(gallery.float().contiguous(), "gallery")
    nq, ng = q.shape[0], g.shape[0]
    qsq = gsq = None
    if metric == "l2":
        qsq, gsq = row_sqnorm(q), row_sqnorm(g)
    out = torch.empty((nq, 4), dtype=torch.float64, device=q.device)
    lib = L.load()
    d = q.shape[1]
    dt = L.KNN_F32
    if nq > 0 and ng > 0:
        q, g, dt = _packed_operands(q, g, None)
    with torch.cuda.device(out.device):
        nbytes = lib.knn_score_stats_workspace(nq, ng)
        ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=out.device)
        rc = lib.knn_score_stats(_ptr(q), _ptr(g), _ptr(qsq), _ptr(gsq), nq, ng, d, dt, _METRICS[metric],
                                 _SELF[self_mode], query_offset, _ptr(out), _ptr(ws), ws.numel(), _stream(q))
    L.check(rc, "knn_score_stats")
    n = float(ng - (1 if self_mode == "exclude" else 0))
    mean = out[:, 0] / n
    var = (out[:, 1] / n - mean * mean).clamp_min(0.0)
    return {"mean": mean, "std": var.sqrt(), "min": out[:, 2], "max": out[:, 3]}


def _row_affine(emb_n: torch.Tensor, mode: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-query (scale, shift) with  normalised_similarity = scale * similarity + shift  for the modes of
    ``normalize_similarity_matrix``; statistics are taken over the full row INCLUDING the self-similarity, as the
    reference does.  Arithmetic in fp32 like the reference's ``(similarity - means) / stds``."""
    nq = emb_n.shape[0]
    one = torch.ones((nq,), dtype=torch.float32, device=emb_n.device)
    if mode == "none":
        return one, torch.zeros_like(one)
    st = score_stats(emb_n, emb_n, "ip")
    if mode == "zscore":
        std = st["std"].float().clamp_min(1e-12)
        return one / std, -(st["mean"].float() / std)
    if mode == "minmax":
        scale = (st["max"] - st["min"]).float().clamp_min(1e-12)
        return one / scale, -(st["min"].float() / scale)
    raise ValueError(f"Unsupported score normalization mode: {mode}. Use one of: none, zscore, minmax")


def _fused_search(conv_n, dino_n, wa, wb, shift, k, exclude_self):
    """top-k of  wa_q * (conv_q . conv_g) + wb_q * (dino_q . dino_g) + shift_q  (self excluded when asked)."""
    qcat = torch.cat([conv_n * wa[:, None], dino_n * wb[:, None]], dim=1).contiguous()
    gcat = torch.cat([conv_n, dino_n], dim=1).contiguous()
    vals, idx = search(qcat, gcat, k, "ip", exclude_self=exclude_self)
    return vals + shift[:, None], idx


def score_fusion_search(conv_embeddings, dino_embeddings, alpha: float, k: int, score_normalization: str = "none",
                        exclude_self: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """evaluate.py:64-84 without the N x N matrices: top-k of ``alpha*norm(S_conv) + (1-alpha)*norm(S_dino)`` for
    the self-retrieval of an aligned embedding set -> (fused scores [N,k], indices [N,k])."""
    conv_n, dino_n = l2_normalize(conv_embeddings), l2_normalize(dino_embeddings)
    sa, ta = _row_affine(conv_n, score_normalization)
    sb, tb = _row_affine(dino_n, score_normalization)
    a, b = float(alpha), 1.0 - float(alpha)
    return _fused_search(conv_n, dino_n, a * sa, b * sb, a * ta + b * tb, k, exclude_self)


def top12_margin_from_embeddings(emb_n: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """evaluate.py:207-214 on the (normalised) self-similarity with the diagonal removed: top1 - top2 per query."""
    vals, _ = search(emb_n, emb_n, 2, "ip", exclude_self=True)
    return (vals[:, 0] - vals[:, 1]) * scale   # an additive row shift cancels in the margin


def confidence_fusion_search(conv_embeddings, dino_embeddings, k: int, score_normalization: str = "none"):
    """evaluate.py:180-204: query-adaptive ``alpha = c_conv / (c_conv + c_dino + 1e-8)`` from the top1-top2
    margins -> ((scores, indices), info) with the reference's bookkeeping fields."""
    conv_n, dino_n = l2_normalize(conv_embeddings), l2_normalize(dino_embeddings)
    sa, ta = _row_affine(conv_n, score_normalization)
    sb, tb = _row_affine(dino_n, score_normalization)
    ca = top12_margin_from_embeddings(conv_n, sa)
    cb = top12_margin_from_embeddings(dino_n, sb)
    alpha = ca / (ca + cb + 1e-8)
    out = _fused_search(conv_n, dino_n, alpha * sa, (1.0 - alpha) * sb, alpha * ta + (1.0 - alpha) * tb, k, True)
    a = alpha.double().cpu().numpy()
    info = {"conv_selected_queries": int(np.sum(a >= 0.5)), "dino_selected_queries": int(np.sum(a < 0.5)),
            "alpha_mean": float(np.mean(a)), "alpha_std": float(np.std(a)), "alpha": alpha}
    return out, info


# --------------------------------------------------------------------------------------------------------------
# experiment loop (run_late_fusion_experiments, evaluate.py:30-149)
# --------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ExperimentResult:
    """evaluate.py:19-27."""

    experiment_name: str
    num_samples: int
    metrics: Dict[str, float]
    skipped: bool = False
    skipped_reason: Optional[str] = None


@dataclass(frozen=True)
class AlignedEmbeddings:
    """Intersection-aligned embedding payload of the fusion experiments (fusion_eval/align.py:27-35)."""

    image_paths: List[str]
    labels: List[str]
    conv_embeddings: np.ndarray
    dino_embeddings: np.ndarray
    coverage: Dict[str, List[str]] = field(default_factory=dict)


def run_late_fusion_experiments(conv_embeddings, dino_embeddings=None, labels: Optional[Sequence] = None,
                                image_paths: Optional[Sequence] = None,
                                alpha_values: Sequence[float] = (0.2, 0.4, 0.5, 0.6, 0.8),
                                k_values: Iterable[int] = (1, 5, 10), include_score_fusion: bool = True,
                                score_normalization: str = "none",
                                include_confidence_fusion: bool = True) -> List[ExperimentResult]:
    """Same experiments, names and metric keys as the reference's loop.  Call it like the reference --
    ``run_late_fusion_experiments(aligned, alpha_values=...)`` with an :class:`AlignedEmbeddings` -- or with the four
    arrays.  Full-ranking metrics (mAP) use k = N-1 searches, so this is meant for the evaluation-set sizes the
    reference runs it on."""
    from . import metrics as M

    if isinstance(conv_embeddings, AlignedEmbeddings):
        aligned = conv_embeddings
        conv_embeddings, dino_embeddings = aligned.conv_embeddings, aligned.dino_embeddings
        labels, image_paths = aligned.labels, aligned.image_paths
    if dino_embeddings is None or labels is None or image_paths is None:
        raise ValueError("pass an AlignedEmbeddings, or conv / dino embeddings with labels and image paths")
    conv, dino = _cuda_f32(conv_embeddings), _cuda_f32(dino_embeddings)
    n = conv.shape[0]
    k_values = list(k_values)
    results: List[ExperimentResult] = []
    baselines = {"convnext_baseline": l2_normalize(conv), "dino_baseline": l2_normalize(dino),
                 "concat_fusion": concat_fusion(conv, dino)}
    for name, emb in baselines.items():
        results.append(ExperimentResult(name, n, M.evaluate_retrieval_metrics(emb, labels, image_paths, k_values)))
    if include_score_fusion:
        for alpha in alpha_values:
            _, idx = score_fusion_search(conv, dino, alpha, n - 1, score_normalization)
            results.append(ExperimentResult(f"score_fusion_alpha_{alpha:.1f}", n,
                                            M.retrieval_metrics_from_ranking(idx, labels, k_values)))
    if include_confidence_fusion:
        (_, idx), info = confidence_fusion_search(conv, dino, n - 1, score_normalization)
        m = M.retrieval_metrics_from_ranking(idx, labels, k_values)
        m["conv_selected_queries"] = float(info["conv_selected_queries"])
        m["dino_selected_queries"] = float(info["dino_selected_queries"])
        results.append(ExperimentResult("confidence_fusion_top12_margin", n, m))
    for alpha in alpha_values:
        fusion = weighted_sum_fusion(conv, dino, alpha)
        if fusion.embeddings is None:
            results.append(ExperimentResult(f"weighted_sum_alpha_{alpha:.1f}", n, {}, True, fusion.skipped_reason))
            continue
        results.append(ExperimentResult(f"weighted_sum_alpha_{alpha:.1f}", n,
                                        M.evaluate_retrieval_metrics(fusion.embeddings, labels, image_paths, k_values)))
    return results


# --------------------------------------------------------------------------------------------------------------
# re-ranking (test.py:612-621, 769-777; retrieval_analysis/rerank.py)
# --------------------------------------------------------------------------------------------------------------
def sort_topk(vals: torch.Tensor, idx: torch.Tensor, largest_first: bool = True):
    """Order every row's k candidates best-first, ties by ascending index (``knn_sort_topk``)."""
    _require_cuda(vals, idx)
    vals, idx = vals.float().contiguous(), idx.long().contiguous()
    nq, k = vals.shape
    ov, oi = torch.empty_like(vals), torch.empty_like(idx)
    lib = L.load()
    with torch.cuda.device(vals.device):
        rc = lib.knn_sort_topk(_ptr(vals), _ptr(idx), nq, k, 1 if largest_first else 0, _ptr(ov), _ptr(oi), _stream(vals))
    L.check(rc, "knn_sort_topk")
    return ov, oi


def rescore_topk(vals: torch.Tensor, idx: torch.Tensor, table: torch.Tensor, query_columns, alpha: float, beta: float,
                 first_m: Optional[int] = None, query_offset: int = 0, exclude_self: bool = True,
                 mask_self: bool = False):
    """``vals[q, j] <- alpha*vals[q, j] + beta*table[idx[q, j], query_columns[q]]`` for the first ``first_m``
    candidates of every query (the query's own entry is left untouched, or set to -inf with ``mask_self``), then the
    row is re-sorted best-first with ties by ascending index -> (vals, idx).  fp32 arithmetic exactly as
    ``alpha * img_sim[i, j] + beta * text_score`` (test.py:618-621)."""
    _require_cuda(vals, idx, table)
    vals, idx = vals.float().contiguous(), idx.long().contiguous()
    table = _as2d(table.float().contiguous(), "table")
    nq, k = vals.shape
    qcol = torch.as_tensor(query_columns, dtype=torch.int64, device=vals.device).contiguous().view(-1)
    if qcol.numel() != nq:
        raise ValueError("one table column per query expected")
    out = torch.empty_like(vals)
    lib = L.load()
    with torch.cuda.device(vals.device):
        rc = lib.knn_rescore_topk(_ptr(vals), _ptr(idx), nq, k, _ptr(table), table.shape[0], table.shape[1], _ptr(qcol),
                                  float(alpha), float(beta), k if first_m is None else int(first_m),
                                  int(query_offset) if exclude_self else -1, 1 if mask_self else 0, _ptr(out),
                                  _stream(vals))
    L.check(rc, "knn_rescore_topk")
    return sort_topk(out, idx, largest_first=True)


def rerank_search(queries: torch.Tensor, gallery: torch.Tensor, table: torch.Tensor, query_columns, k: int,
                  rerank_k: int, alpha: float, beta: Optional[float] = None, metric: str = "cosine",
                  normalize: bool = False, exclude_self: bool = True, query_offset: int = 0):
    """The re-ranking strategy of test.py:596-623 for the top-k only: retrieve, re-score the best ``rerank_k`` hits
    with the text-score table, re-sort.  Exactness of the final top-k: a hit outside the first ``rerank_k`` keeps its
    original score, and at most k of them can enter the final top-k -- they are the original ranks
    ``rerank_k+1 .. rerank_k+k``, so ``rerank_k + k`` candidates are retrieved.  Note (reference quirk kept): the
    reference picks its ``rerank_k`` candidates BEFORE masking the diagonal, so the query itself occupies one slot;
    the diagonal is masked afterwards (``fill_diagonal_(-inf)``, test.py:623)."""
    beta = 1.0 - alpha if beta is None else beta
    kk = min(gallery.shape[0], rerank_k + k)
    vals, idx = search(queries, gallery, kk, metric, normalize=normalize, exclude_self=False, query_offset=query_offset)
    v2, i2 = rescore_topk(vals, idx, table, query_columns, alpha, beta, first_m=min(rerank_k, kk),
                          query_offset=query_offset, exclude_self=True, mask_self=exclude_self)
    if exclude_self:  # the masked self entry sorted last with -inf: report it as "no candidate"
        i2 = torch.where(torch.isinf(v2) & (v2 < 0), torch.full_like(i2, -1), i2)
    return v2[:, :k].contiguous(), i2[:, :k].contiguous()


class Reranker(Protocol):
    """retrieval_analysis/rerank.py:10-17: ``rerank(query, results) -> iterable of results``."""

    def rerank(self, query, results: Iterable) -> Iterable:
        ...


class IdentityReranker:
    """retrieval_analysis/rerank.py:20-25."""

    def rerank(self, query, results: Iterable) -> Iterable:
        return list(results)


# --------------------------------------------------------------------------------------------------------------
# lesion (region) re-ranking  (ChestMIR/chestmir_eval.py:275-318, 461-650)
# --------------------------------------------------------------------------------------------------------------
def normalize_lesion_text(name) -> str:
    """chestmir_eval.py:281-285: lower-case, `_ - /` -> space, collapse blanks."""
    text = str(name).strip().lower().replace("_", " ").replace("-", " ").replace("/", " ")
    return " ".join(text.split())


def canonical_lesion_name(name, alias_to_canon: Optional[Dict[str, str]] = None) -> str:
    """chestmir_eval.py:288-290; the alias table (LESION_ALIAS_GROUPS) is the caller's data."""
    normalized = normalize_lesion_text(name)
    return (alias_to_canon or {}).get(normalized, normalized)


class LesionIndex:
    """Device-resident region vectors of a gallery: ``lesion_maps[i]`` = {canonical lesion -> [unit vectors]} (the
    output of ``build_lesion_vector_map``), stored CSR over (image, lesion slot) for ``knn_lesion_rerank``."""

    def __init__(self, lesion_maps: Sequence[Dict[str, Sequence[np.ndarray]]], device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.maps = lesion_maps
        names = sorted({k for m in lesion_maps for k, v in m.items() if len(v)})
        self.slot = {name: s for s, name in enumerate(names)}
        self.n_slots = max(1, len(names))
        self.n_img = len(lesion_maps)
        dl = next((len(v[0]) for m in lesion_maps for v in m.values() if len(v)), 1)
        offsets = np.zeros(self.n_img * self.n_slots + 1, dtype=np.int64)
        rows: List[np.ndarray] = []
        for i, m in enumerate(lesion_maps):
            for name, vecs in m.items():
                if len(vecs):
                    offsets[i * self.n_slots + self.slot[name] + 1] = len(vecs)
        np.cumsum(offsets, out=offsets)
        flat = np.zeros((max(1, int(offsets[-1])), dl), dtype=np.float32)
        for i, m in enumerate(lesion_maps):
            for name, vecs in m.items():
                if len(vecs):
                    b = offsets[i * self.n_slots + self.slot[name]]
                    flat[b:b + len(vecs)] = np.asarray(vecs, dtype=np.float32)
        self.dl = dl
        self.offsets = torch.from_numpy(offsets).to(self.device)
        self.vectors = torch.from_numpy(flat).to(self.device)

    def query_vector(self, i: int, lesion: str):
        """choose_query_lesion_vector (chestmir_eval.py:461-469): the image's FIRST vector of that lesion."""
        cands = self.maps[i].get(lesion, [])
        return cands[0] if len(cands) else None

    def adaptive_query_vector(self, i: int, target_keys: Sequence[str]):
        """choose_query_adaptive_lesion_vector (chestmir_eval.py:484-505): the target lesion with the most regions in
        the image (first such target on ties)."""
        best_name, best_vec, best_count = None, None, -1
        for name in target_keys:
            cands = self.maps[i].get(name, [])
            if len(cands) > best_count and len(cands):
                best_count, best_name, best_vec = len(cands), name, cands[0]
        return best_name, best_vec


def lesion_rerank_search(global_vectors, lesion_maps, k: int, rerank_topk: int, global_weight: float,
                         lesion_name: Optional[str] = None, target_lesions: Optional[Sequence[str]] = None,
                         alias_to_canon: Optional[Dict[str, str]] = None, index: Optional[LesionIndex] = None):
    """Two-stage self-retrieval of chestmir_eval.py: global cosine ranking, then the top ``rerank_topk`` candidates of
    every query re-ordered by ``gw * global + (1 - gw) * best region cosine`` for ONE lesion (``lesion_name``:
    rerank_with_specific_lesion) or for the query's dominant target lesion (``target_lesions``:
    rerank_with_adaptive_lesion).  -> (indices [N, K] with K = max(k, rerank_topk), stats dict with the reference's
    keys).  No N x N matrix: stage 1 is the fused search, stage 2 ``knn_lesion_rerank`` on its candidate lists."""
    if (lesion_name is None) == (target_lesions is None):
        raise ValueError("pass exactly one of lesion_name / target_lesions")
    g = _cuda_f32(global_vectors)
    n = g.shape[0]
    topk = min(int(rerank_topk), n - 1)
    K = min(max(int(k), topk), n - 1)
    vals, idx = search(g, g, K, "ip", exclude_self=True)
    index = index or LesionIndex(lesion_maps, g.device)
    qslot = np.full(n, -1, dtype=np.int32)
    qvec = np.zeros((n, index.dl), dtype=np.float32)
    usage: Dict[str, int] = {}
    chosen: List[Optional[str]] = [None] * n
    if lesion_name is not None:
        key = canonical_lesion_name(lesion_name, alias_to_canon)
        for i in range(n):
            v = index.query_vector(i, key)
            if v is not None:
                qslot[i], qvec[i], chosen[i] = index.slot[key], v, key
    else:
        keys = [canonical_lesion_name(x, alias_to_canon) for x in target_lesions]
        for i in range(n):
            name, v = index.adaptive_query_vector(i, keys)
            if v is not None:
                qslot[i], qvec[i], chosen[i] = index.slot[name], v, name
    out_idx = torch.empty_like(idx)
    matched = torch.empty((n,), dtype=torch.int32, device=g.device)
    qv, qs = torch.from_numpy(qvec).to(g.device), torch.from_numpy(qslot).to(g.device)
    with torch.cuda.device(g.device):
        rc = L.load().knn_lesion_rerank(_ptr(vals), _ptr(idx), n, K, topk, _ptr(qv), _ptr(qs), _ptr(index.offsets),
                                        _ptr(index.vectors), index.n_img, index.n_slots, index.dl, float(global_weight),
                                        _ptr(out_idx), None, _ptr(matched), _stream(g))
    L.check(rc, "knn_lesion_rerank")
    mt = matched.cpu().numpy()
    reranked = int((mt > 0).sum())
    for i in np.nonzero(mt > 0)[0]:
        usage[chosen[i]] = usage.get(chosen[i], 0) + 1
    total_topk = n * topk
    total_matched = int(mt[mt > 0].sum())
    stats = {
        "queries_total": n, "queries_reranked": reranked, "queries_fallback_global": n - reranked,
        "queries_with_candidate_match": reranked, "matched_candidates_in_topk": total_matched,
        "candidate_match_rate_pct": (100.0 * total_matched / total_topk) if total_topk > 0 else 0.0,
        "rerank_topk": rerank_topk, "global_weight": global_weight, "region_weight": 1.0 - global_weight,
    }
    if lesion_name is not None:
        stats = {"lesion": lesion_name, **stats}
    else:
        stats = {"mode": "adaptive", **stats, "lesion_usage": usage}
    return out_idx, stats
