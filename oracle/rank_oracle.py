"""CPU restatement of knn_rank_of_positives / knn_ap_from_ranks / knn_ap_sklearn_from_ranks -- TEST INFRASTRUCTURE ONLY.

What the reference does with a full ranking (``argsort`` of a whole score row, test.py:1090 / test.py:962 / train.py:409,
then a look at where the relevant rows sit) restated over the deterministic order (best score first, ties by ascending
gallery row): positions of the relevant rows, the end of each positive's run of equal scores and the number of distinct
scores above it.  The AP variants built on top are the ones of oracle/reference_metrics.py (compute_ap,
average_precision_ranked ...), which are pinned against the real reference by the goldens.
"""
from __future__ import annotations

import numpy as np


def order_row(scores: np.ndarray, largest_first: bool = True) -> np.ndarray:
    """Gallery rows best first, ties by ascending row (a stable sort of the negated / plain scores; -0.0 == +0.0)."""
    s = np.asarray(scores, dtype=np.float32) + np.float32(0.0)
    return np.argsort(-s if largest_first else s, kind="stable")


def rank_of_positives(scores: np.ndarray, rel: np.ndarray, largest_first: bool = True, dropped: np.ndarray = None):
    """scores [N] fp32, rel [N] bool, dropped [N] bool (rows that are not ranked) ->
    (ranks of the positives ascending, ge, tgroup, nranked, ngroups)."""
    keep = np.ones(len(scores), dtype=bool) if dropped is None else ~np.asarray(dropped, dtype=bool)
    rows = np.flatnonzero(keep)
    order = rows[order_row(np.asarray(scores)[rows], largest_first)]
    s_sorted = (np.asarray(scores, dtype=np.float32)[order] + np.float32(0.0))
    r_sorted = np.asarray(rel, dtype=bool)[order]
    n = len(order)
    start = np.r_[True, s_sorted[1:] != s_sorted[:-1]] if n else np.zeros(0, bool)
    tg_all = np.cumsum(start) - 1
    starts = np.flatnonzero(start)
    ends = np.r_[starts[1:], n] if n else starts
    ge_all = ends[tg_all] if n else tg_all
    pos = np.flatnonzero(r_sorted)
    return pos.astype(np.int64), ge_all[pos].astype(np.int64), tg_all[pos].astype(np.int64), n, int(start.sum())
