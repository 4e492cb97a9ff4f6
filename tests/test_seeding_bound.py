"""CPU: the exactness argument of the chunk-maxima seeding pre-pass (csrc/select.cuh: seed_tile_tmem), restated in numpy.

The pre-pass owes the main pass a LOWER BOUND of every query's final k-th best score.  It takes the k-th largest of the
maxima of disjoint groups of gallery rows (runs of `stride` 32-row chunks per selection thread; chunks that hold the
query's own row or padding columns are skipped).  Maxima of n disjoint groups are the scores of n distinct rows, so at
least k eligible rows reach the k-th largest of them -- whatever the ties, the self mode or the group geometry."""
import numpy as np
import pytest


def _seed_bound(scores, k, stride, groups, unit_len, self_row):
    """k-th largest group maximum over a sample laid out as the kernel lays it out: units of `unit_len` rows, tiles of 256
    rows = 8 chunks of 32, selection thread `g` of a row owns chunks g, g + groups, ... of every tile and flushes one
    maximum per `stride` of ITS chunks (and what is left at the end of the unit).  None = fewer than k maxima."""
    maxima = []
    for u0 in range(0, len(scores), unit_len):
        unit = scores[u0:u0 + unit_len]
        for g in range(groups):
            best, since = -np.inf, 0
            for t0 in range(0, len(unit), 256):
                for ch in range(g, 8, groups):
                    c0 = u0 + t0 + ch * 32
                    whole = c0 + 32 <= min(u0 + len(unit), len(scores))
                    if whole and not (c0 <= self_row < c0 + 32):
                        m = scores[c0:c0 + 32].max()
                        best = max(best, m)
                    since += 1
                    if since >= stride:
                        if best > -np.inf:
                            maxima.append(best)
                        best, since = -np.inf, 0
            if since > 0 and best > -np.inf:
                maxima.append(best)
    if len(maxima) < k:
        return None
    return np.sort(np.asarray(maxima))[::-1][k - 1]


@pytest.mark.parametrize("seed", range(6))
def test_kth_largest_group_maximum_never_exceeds_the_kth_best_score(seed):
    rs = np.random.RandomState(seed)
    for _ in range(40):
        n = int(rs.choice([2048, 4096, 8192, 10240]))
        unit_len = int(rs.choice([768, 1024, 2048, 4096]))
        stride = int(rs.choice([1, 2, 3, 4, 1 << 30]))
        groups = 2
        k = int(rs.choice([1, 10, 32, 50, 100]))
        kind = rs.randint(4)
        if kind == 0:      # few distinct values: thousands of ties at every cut-off
            s = rs.randint(0, 6, size=n).astype(np.float32) / 8.0
        elif kind == 1:    # descending by row: every chunk maximum sits in the chunk's first column
            s = np.linspace(1.0, -1.0, n).astype(np.float32)
        elif kind == 2:    # the best rows are contiguous (a class-sorted gallery): many of them share a group
            s = rs.standard_normal(n).astype(np.float32) * 0.05
            s[100:100 + 3 * k] += 1.0
        else:
            s = rs.standard_normal(n).astype(np.float32)
        self_row = int(rs.randint(-1, n))            # -1: no self row (keep mode)
        bound = _seed_bound(s, k, stride, groups, unit_len, self_row)
        if bound is None:
            continue
        eligible = np.delete(s, self_row) if self_row >= 0 else s
        kth = np.sort(eligible)[::-1][k - 1]
        assert bound <= kth, (n, unit_len, stride, k, kind, self_row, bound, kth)


def test_bound_is_tight_on_iid_scores():
    """With 32..64 rows per group the bound sits a few ranks below the sample's exact k-th best."""
    rs = np.random.RandomState(11)
    s = rs.standard_normal(65536).astype(np.float32)
    bound = _seed_bound(s, 100, 2, 2, 32768, -1)
    rank = int((s >= bound).sum())
    assert 100 <= rank <= 125, rank
