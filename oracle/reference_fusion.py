"""CPU restatement (numpy) of the reference's late-fusion and re-ranking arithmetic -- TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows.  Pinned against the real reference by
``oracle/make_golden_fusion.py`` -> ``tests/golden/golden_fusion*.{json,npz}`` (``tests/test_oracle_golden.py``).
Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline leg may import this.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np


def l2_normalize(embeddings: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """fusion_eval/fuse.py:11-15."""
    norms = np.linalg.norm(embeddings, axis=1, keepdims=True)
    return embeddings / np.maximum(norms, eps)


def concat_fusion(conv: np.ndarray, dino: np.ndarray) -> np.ndarray:
    """fusion_eval/fuse.py:18-23."""
    return l2_normalize(np.concatenate([l2_normalize(conv), l2_normalize(dino)], axis=1))


def weighted_sum_fusion(conv: np.ndarray, dino: np.ndarray, alpha: float):
    """fusion_eval/fuse.py:35-52 (None when the dimensions differ)."""
    if conv.shape[1] != dino.shape[1]:
        return None
    return l2_normalize(alpha * l2_normalize(conv) + (1.0 - alpha) * l2_normalize(dino))


def compute_similarity_matrix(embeddings: np.ndarray) -> np.ndarray:
    """fusion_eval/metrics.py:12-15."""
    n = l2_normalize(embeddings.astype(np.float32))
    return n @ n.T


def normalize_similarity_matrix(similarity: np.ndarray, mode: str = "none") -> np.ndarray:
    """fusion_eval/evaluate.py:152-177: row-wise zscore / minmax over the FULL row (diagonal included in the
    statistics, restored afterwards)."""
    similarity = similarity.astype(np.float32, copy=True)
    if mode == "none":
        return similarity
    diag = np.diag(similarity).copy()
    if mode == "zscore":
        means = np.mean(similarity, axis=1, keepdims=True)
        stds = np.maximum(np.std(similarity, axis=1, keepdims=True), 1e-12)
        out = (similarity - means) / stds
    elif mode == "minmax":
        mins = np.min(similarity, axis=1, keepdims=True)
        scales = np.maximum(np.max(similarity, axis=1, keepdims=True) - mins, 1e-12)
        out = (similarity - mins) / scales
    else:
        raise ValueError(mode)
    np.fill_diagonal(out, diag)
    return out


def top12_margin(similarity: np.ndarray) -> np.ndarray:
    """fusion_eval/evaluate.py:207-214."""
    top2 = np.partition(similarity, kth=-2, axis=1)[:, -2:]
    return np.max(top2, axis=1) - np.min(top2, axis=1)


def confidence_based_fusion(conv_similarity: np.ndarray, dino_similarity: np.ndarray) -> Dict:
    """fusion_eval/evaluate.py:180-204."""
    cs = conv_similarity.astype(np.float32, copy=True)
    ds = dino_similarity.astype(np.float32, copy=True)
    np.fill_diagonal(cs, -np.inf)
    np.fill_diagonal(ds, -np.inf)
    cc, dc = top12_margin(cs), top12_margin(ds)
    alpha = cc / (cc + dc + 1e-8)
    return {"similarity": alpha[:, None] * cs + (1.0 - alpha[:, None]) * ds, "alpha": alpha,
            "conv_selected_queries": int(np.sum(alpha >= 0.5)), "dino_selected_queries": int(np.sum(alpha < 0.5))}


def score_fusion_similarity(conv: np.ndarray, dino: np.ndarray, alpha: float, mode: str = "none") -> np.ndarray:
    """fusion_eval/evaluate.py:64-73."""
    cs = normalize_similarity_matrix(compute_similarity_matrix(conv), mode)
    ds = normalize_similarity_matrix(compute_similarity_matrix(dino), mode)
    return alpha * cs + (1.0 - alpha) * ds


def stable_self_ranking(similarity: np.ndarray):
    """fusion_eval/metrics.py:18-22 made deterministic: diagonal -inf, best-first, ties by ascending
    index; the self entry (ranked last) is dropped -> (values [N, N-1], indices [N, N-1])."""
    s = similarity.astype(np.float32, copy=True)
    np.fill_diagonal(s, -np.inf)
    order = np.argsort(-s, axis=1, kind="stable")[:, :-1]
    return np.take_along_axis(s, order, axis=1), order


def rerank_rows(img_sim: np.ndarray, table: np.ndarray, labels: Sequence[int], rerank_k: int, alpha: float,
                beta: float) -> np.ndarray:
    """test.py:608-623 (and :766-777): per query the top ``rerank_k`` of the UNMASKED image-similarity row are
    re-scored as ``alpha*img_sim[i, j] + beta*table[j, labels[i]]`` (j != i) in fp32, everything else keeps its
    image similarity, then the diagonal is set to -inf.  -> dists [N, N] fp32."""
    n = img_sim.shape[0]
    dists = img_sim.astype(np.float32, copy=True)
    a32, b32 = np.float32(alpha), np.float32(beta)
    for i in range(n):
        top = np.argsort(-img_sim[i], kind="stable")[: min(rerank_k, n)]
        for j in top:
            if i != j:
                dists[i, j] = np.float32(a32 * img_sim[i, j]) + np.float32(b32 * table[j, labels[i]])
    np.fill_diagonal(dists, -np.inf)
    return dists
