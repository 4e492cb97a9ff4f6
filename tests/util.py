"""Shared helpers for the parity tests."""
import numpy as np


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units of last place between two float32 arrays (same sign region assumed)."""
    ai = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2**31) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2**31) - bi, bi)
    return np.abs(ai - bi)


def tie_aware_mismatches(val_a, idx_a, val_b, idx_b, tol):
    """Tie-aware comparison of two top-k results.

    A position may differ only if the two candidates' scores differ by <= tol (a near-tie that two fp32
    summation orders may legitimately resolve differently).  Returns (n_positions_differing, n_violations).
    """
    val_a, val_b = np.asarray(val_a, dtype=np.float64), np.asarray(val_b, dtype=np.float64)
    idx_a, idx_b = np.asarray(idx_a), np.asarray(idx_b)
    diff = idx_a != idx_b
    bad = 0
    for r, j in zip(*np.nonzero(diff)):
        where = np.nonzero(idx_b[r] == idx_a[r, j])[0]
        if where.size:  # same candidate, other slot: slots must hold near-equal scores
            ok = abs(val_b[r, where[0]] - val_b[r, j]) <= tol
        else:  # swapped across the k boundary: must be within tol of the k-th score
            ok = abs(val_a[r, j] - val_b[r, -1]) <= tol
        bad += 0 if ok else 1
    return int(diff.sum()), bad


def rel_close(a, b, rtol=1e-5, atol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    return np.all(both_inf | (np.abs(a - b) <= atol + rtol * np.abs(b)))
