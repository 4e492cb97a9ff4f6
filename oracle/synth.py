"""Seeded synthetic embeddings / labels for the parity cases (TEST INFRASTRUCTURE).

numpy's legacy ``RandomState`` stream is frozen across numpy versions, so the golden fixtures only need to
store seeds: every host regenerates bit-identical inputs (SURVEY 8(d) input table).
"""
from __future__ import annotations

import numpy as np

NIH_PRIORS = np.array([0.10, 0.025, 0.12, 0.18, 0.05, 0.055, 0.012, 0.045, 0.04, 0.02, 0.022, 0.015, 0.03, 0.002])


def clustered(n: int, d: int, n_classes: int, seed: int, noise: float = 0.8, priors=None):
    """Class-clustered Gaussians: x = mu_c + noise * N(0, I); labels int64.  -> (x fp32 [n,d], labels [n])"""
    rs = np.random.RandomState(seed)
    mu = rs.standard_normal((n_classes, d))
    if priors is None:
        labels = np.arange(n) % n_classes
        rs.shuffle(labels)
    else:
        labels = rs.choice(n_classes, size=n, p=np.asarray(priors) / np.sum(priors))
    x = mu[labels] + noise * rs.standard_normal((n, d))
    return x.astype(np.float32), labels.astype(np.int64)


def exact_grid(n: int, d: int, seed: int, n_dup: int = 8):
    """Entries in {-1, 0, +1} / 32 with `n_dup` duplicated rows: every product and every partial sum is exactly
    representable in fp32 (and in bf16 inputs / fp32 accumulate), so scores are identical for ANY summation
    order and ties are real.  -> x fp32 [n, d]"""
    rs = np.random.RandomState(seed)
    x = rs.randint(-1, 2, size=(n, d)).astype(np.float32) / 32.0
    if n_dup > 0 and n > 2 * n_dup:
        src = rs.choice(n, size=n_dup, replace=False)
        dst = rs.choice(n, size=n_dup, replace=False)
        x[dst] = x[src]
    return x


def multihot(n: int, seed: int, n_labels: int = 14, max_per_row: int = 4, priors=NIH_PRIORS):
    """NIH-like multi-hot rows: 1..max_per_row labels each, drawn from the pathology priors.  -> fp32 [n, C]"""
    rs = np.random.RandomState(seed)
    p = np.asarray(priors[:n_labels], dtype=np.float64)
    p = p / p.sum()
    out = np.zeros((n, n_labels), dtype=np.float32)
    counts = rs.choice(np.arange(1, max_per_row + 1), size=n, p=[0.6, 0.25, 0.1, 0.05][:max_per_row] /
                       np.sum([0.6, 0.25, 0.1, 0.05][:max_per_row]))
    for i in range(n):
        out[i, rs.choice(n_labels, size=counts[i], replace=False, p=p)] = 1.0
    return out


def labelset_clustered(labels_multihot: np.ndarray, d: int, seed: int, noise: float = 0.8):
    """Embeddings keyed by the label SET: mean of the per-label centres + noise."""
    rs = np.random.RandomState(seed)
    mu = rs.standard_normal((labels_multihot.shape[1], d))
    base = labels_multihot @ mu / np.maximum(labels_multihot.sum(axis=1, keepdims=True), 1.0)
    return (base + noise * rs.standard_normal((labels_multihot.shape[0], d))).astype(np.float32)
