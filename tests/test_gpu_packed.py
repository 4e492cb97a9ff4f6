"""GPU parity tests of the packed operand path of the exact-fp32 FFMA kernel (KNN_F32_PACKED, csrc/search_f32.cu):
the rows are transposed once into 128-row tiles and the kernel fills its operand ring with bulk copies.  Same fmaf
chain, so everything is compared bit for bit -- against the CPU oracle and against the row-major path."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knn():
    import b200knn

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    b200knn.load_library()
    return b200knn


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture()
def packed(monkeypatch):
    monkeypatch.setenv("KNN_F32_PACK", "1")
    monkeypatch.setenv("KNN_EXACT_ENGINE", "ffma")


@pytest.mark.parametrize("n,d", [(1, 1), (127, 16), (128, 17), (129, 100), (1000, 33), (300, 1024)])
def test_pack_layout(knn, n, d):
    from b200knn.search import PackedRows

    rs = np.random.RandomState(n + d)
    x = rs.standard_normal((n, d)).astype(np.float32)
    p = PackedRows.build(dev(x))
    tiles, dpad = (n + 127) // 128, (d + 15) // 16 * 16
    want = np.zeros((tiles * 128, dpad), np.float32)
    want[:n, :d] = x
    want = want.reshape(tiles, 128, dpad).transpose(0, 2, 1)            # [tile][k][row in tile]
    assert p.n == n and p.d == d
    assert np.array_equal(host(p.data)[: tiles * dpad * 128].reshape(tiles, dpad, 128), want)


@pytest.mark.parametrize("nq,ng,d,k", [
    (1, 1, 8, 1), (1, 100, 36, 5), (7, 127, 64, 32), (128, 128, 64, 33), (129, 129, 100, 50), (300, 1000, 96, 100),
    (33, 5000, 48, 128), (40, 3000, 32, 200), (17, 2000, 17, 256), (5, 3, 8, 10), (260, 40_000, 8, 64),
])
@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_packed_search_bit_exact_vs_oracle(knn, packed, nq, ng, d, k, metric):
    rs = np.random.RandomState(nq * 7 + ng)
    g = oracle.normalize(rs.standard_normal((ng, d)).astype(np.float32))
    q = oracle.normalize(rs.standard_normal((nq, d)).astype(np.float32))
    v, i = knn.search(dev(q), dev(g), k, metric)
    ov, oi = oracle.search(q, g, k, metric, "keep", 0)
    assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov)


@pytest.mark.parametrize("metric,self_mode", [("cosine", "exclude"), ("ip", "minus1"), ("l2", "exclude"), ("l2", "keep")])
def test_packed_dense_and_self_modes_bit_exact(knn, packed, metric, self_mode):
    rs = np.random.RandomState(2)
    g = oracle.normalize(rs.standard_normal((333, 50)).astype(np.float32))
    q = g[100:230].copy()
    s = knn.scores_dense(dev(q), dev(g), metric, self_mode=self_mode, query_offset=100)
    assert np.array_equal(host(s), oracle.scores(q, g, metric, self_mode, 100))
    if self_mode != "minus1" or metric != "l2":
        v, i = knn.search(dev(q), dev(g), 20, metric, self_mode=self_mode, query_offset=100)
        ov, oi = oracle.search(q, g, 20, metric, self_mode, 100)
        assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov)


def test_packed_equals_row_major_on_a_long_unit(knn, monkeypatch):
    """Many tiles per unit and d = 1024 (128 k-blocks per tile): the ring wraps thousands of times, the producer runs
    ahead across tile boundaries.  Packed and row-major kernels must agree bit for bit (search, dense, statistics)."""
    from b200knn.fusion import score_stats

    gen = torch.Generator(device="cuda").manual_seed(5)
    g = knn.normalize(torch.randn((60_000, 1024), generator=gen, device="cuda"))
    q = knn.normalize(torch.randn((700, 1024), generator=gen, device="cuda"))
    monkeypatch.setenv("KNN_EXACT_ENGINE", "ffma")
    res = {}
    for pack in ("0", "1"):
        monkeypatch.setenv("KNN_F32_PACK", pack)
        v, i = knn.search(q, g, 100, "cosine")
        vl, il = knn.search(q[:200], g, 10, "l2")
        s = knn.scores_dense(q[:300], g[:20_001], "cosine")
        st = score_stats(q[:130], g, "ip")
        res[pack] = (v, i, vl, il, s, st["mean"], st["std"], st["min"], st["max"])
    for a, b in zip(res["0"], res["1"]):
        assert torch.equal(a, b)
    # and against the oracle on a slice of the queries
    ov, oi = oracle.search(host(q[:8]), host(g), 100, "cosine", "keep", 0)
    assert np.array_equal(host(res["1"][1][:8]), oi) and np.array_equal(host(res["1"][0][:8]), ov)


def test_flat_index_keeps_the_packed_gallery(knn, packed):
    rs = np.random.RandomState(3)
    g = rs.standard_normal((3000, 40)).astype(np.float32)
    q = rs.standard_normal((150, 40)).astype(np.float32)
    index = knn.FlatIndex(40, "cosine", "fp32", normalize=True).add(dev(g))
    v1, i1 = index.search(dev(q), 10)
    assert index._packed is not None and index._packed.n == 3000
    keep = index._packed
    v2, i2 = index.search(dev(q[:60]), 10, exclude_self=True, query_offset=5)
    assert index._packed is keep
    ov, oi = oracle.search(oracle.normalize(q), oracle.normalize(g), 10, "cosine", "keep", 0)
    assert np.array_equal(host(i1), oi) and np.array_equal(host(v1), ov)
    ov, oi = oracle.search(oracle.normalize(q[:60]), oracle.normalize(g), 10, "cosine", "exclude", 5)
    assert np.array_equal(host(i2), oi) and np.array_equal(host(v2), ov)
    index.add(dev(g[:10]))
    assert index._packed is None


def test_packed_abi_validation(knn):
    from b200knn import _lib as L

    lib = L.load()
    assert lib.knn_pack_f32_bytes(0, 8) == 0 and lib.knn_pack_f32_bytes(129, 17) == 2 * 32 * 128 * 4
    assert lib.knn_pack_f32(None, 4, 0, None, None) == -1                  # bad shape
    x = torch.zeros((4, 8), device="cuda")
    assert lib.knn_pack_f32(x.data_ptr(), 4, 8, None, None) == -1          # null out
