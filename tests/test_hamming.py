"""Hamming row: packed-code search through the C ABI against the reference's own
`pairwise_distance(..., binary_codes=True)` + ranking (golden vectors) and against a numpy restatement."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_hamming as mgh
from oracle import reference_metrics as rm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def hamming_oracle(q, g):
    """test_ath.py:80-86: `(q[:, None, :] != g[None, :, :]).sum(dim=2).float()` on int16 codes."""
    return (q.astype(np.int16)[:, None, :] != g.astype(np.int16)[None, :, :]).sum(axis=2).astype(np.float32)


def test_oracle_restatement_matches_the_reference_goldens():
    ga = dict(np.load(os.path.join(GOLDEN, "golden_hamming_arrays.npz")))
    gj = json.load(open(os.path.join(GOLDEN, "golden_hamming.json")))
    for name, c in (("h36", mgh.CASE), ("h200", mgh.CASE_WIDE)):
        q, ql, g, gl = mgh.codes(c)
        d = hamming_oracle(q, g)
        order = np.argsort(d, axis=1, kind="stable")
        assert np.array_equal(order[:, :10], ga[f"{name}_top10_idx"])
        assert np.array_equal(np.take_along_axis(d, order[:, :10], 1), ga[f"{name}_top10_dist"])
        got = rm.ath_compute_metrics(order, ql, gl, (1, 5, 10))
        for k, want in gj[name].items():
            for m, v in want.items():
                # the reference ranks ties with an unstable argsort: rank-sensitive metrics may differ in the last digits
                assert abs(got[int(k)][m] - v) <= (0.02 if m in ("map", "mrr", "majority_acc", "mhr", "mp@k", "r@k") else 0), (name, k, m)


@pytest.mark.gpu
@pytest.mark.parametrize("name,case", [("h36", mgh.CASE), ("h200", mgh.CASE_WIDE)])
def test_hamming_search_bit_exact(name, case):
    import b200knn

    ga = dict(np.load(os.path.join(GOLDEN, "golden_hamming_arrays.npz")))
    q, ql, g, gl = mgh.codes(case)
    tq, tg = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    words = b200knn.pack_bits(tg)
    assert words.shape == (g.shape[0], (case["bits"] + 63) // 64)
    w = words.cpu().numpy().view(np.uint64)
    bit = ((w[:, :, None] >> np.arange(64, dtype=np.uint64)[None, None, :]) & np.uint64(1)).reshape(g.shape[0], -1)
    assert np.array_equal(bit[:, : case["bits"]].astype(np.float32), g) and not bit[:, case["bits"]:].any()
    for method in ("popc", "mma", "auto"):   # xor + popcount over packed words / +-1 rows on the tensor cores
        v, i = b200knn.search_hamming(tq, tg, 10, method=method)
        assert np.array_equal(i.cpu().numpy(), ga[f"{name}_top10_idx"])   # integer distances: bit-exact, ties by row
        assert np.array_equal(v.cpu().numpy(), ga[f"{name}_top10_dist"])
        v2, i2 = b200knn.search_hamming(b200knn.pack_bits(tq), words, 10, packed=True, method=method)
        assert torch.equal(i, i2) and torch.equal(v, v2)
        v3, i3 = b200knn.search_hamming(b200knn.pack_bits(tq)[:5], words, 10, packed=True, bits=case["bits"], method=method)
        assert torch.equal(i[:5], i3) and torch.equal(v[:5], v3)
    # full metric dict through the reference-named entry point
    got = b200knn.metrics.compute_metrics(tq, torch.from_numpy(ql), tg, torch.from_numpy(gl), None, (1, 5, 10), True)
    d = hamming_oracle(q, g)
    want = rm.ath_compute_metrics(np.argsort(d, axis=1, kind="stable"), ql, gl, (1, 5, 10))
    for k in (1, 5, 10):
        for m, val in want[k].items():
            assert abs(got["retrieval"][k][m] - val) < 1e-12, (k, m)


@pytest.mark.gpu
def test_hamming_self_exclusion_large_k_and_mass_ties():
    import b200knn

    rs = np.random.RandomState(2)
    x = rs.randint(0, 2, size=(5000, 64)).astype(np.float32)
    x[100:400] = x[7]                                    # 300 identical codes: a block of exact ties
    t = torch.from_numpy(x).cuda()
    d = hamming_oracle(x[:300], x)
    d[np.arange(300), np.arange(300)] = np.inf
    order = np.argsort(d, axis=1, kind="stable")[:, :256]
    for method in ("popc", "mma"):
        v, i = b200knn.search_hamming(t[:300], t, 256, exclude_self=True, method=method)
        assert np.array_equal(i.cpu().numpy(), order)
        assert np.array_equal(v.cpu().numpy(), np.take_along_axis(d, order, 1))
    v, i = b200knn.search_hamming(t[:3], t[:40], 64, method="mma")      # fewer rows than k: (+inf, -1) tail
    assert (i.cpu().numpy()[:, 40:] == -1).all() and np.isposinf(v.cpu().numpy()[:, 40:]).all()
    v, i = b200knn.search_hamming(t[1000:1003], t, 5, exclude_self=True, query_offset=1000)
    assert not (i.cpu().numpy() == np.arange(1000, 1003)[:, None]).any()


@pytest.mark.gpu
def test_hamming_any_code_length_and_k_beyond_the_fused_limit():
    """The reference ranks the whole Hamming matrix (test_ath.py:80-100), so any code length and any topk work: 300-bit
    codes (5 words: no popcount instantiation) and k = 300 > 256 go through the +-1 rows (dense + full ranking)."""
    import b200knn

    rs = np.random.RandomState(9)
    q = (rs.rand(37, 300) < 0.5).astype(np.float32)
    g = (rs.rand(900, 300) < 0.5).astype(np.float32)
    g[5] = g[700]                                             # ties
    dist = (q[:, None, :] != g[None, :, :]).sum(2).astype(np.float32)
    for k in (10, 300):
        v, i = b200knn.search_hamming(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), k)
        order = np.argsort(dist, axis=1, kind="stable")[:, :k]
        assert np.array_equal(i.cpu().numpy(), order)
        assert np.array_equal(v.cpu().numpy(), np.take_along_axis(dist, order, 1))
    with pytest.raises(b200knn.KnnError):
        b200knn.search_hamming(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), 300, method="popc")
