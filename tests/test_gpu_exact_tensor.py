"""GPU parity tests of the tensor-core exact engine (csrc/exact_tc.cu): precision="fp32" served by the bf16x3 split
filter + exact fp32 re-scoring + completeness proof + FFMA re-run.  Bar: bit-identical to the CPU oracle (and hence
to the FFMA engine) -- indices AND distances -- on every input, including ties and mass duplicates."""
import importlib

import numpy as np
import pytest
import torch

import oracle
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knn():
    import b200knn

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    b200knn.load_library()
    return b200knn


@pytest.fixture()
def tensor_engine(monkeypatch):
    monkeypatch.setenv("KNN_EXACT_ENGINE", "tensor")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def _check(knn, q, g, k, metric, self_mode="keep", offset=0):
    S = importlib.import_module("b200knn.search")

    assert S.exact_engine(q.shape[0], g.shape[0], q.shape[1], k) == "tensor"
    v, i = knn.search(dev(q), dev(g), k, metric, self_mode=self_mode, query_offset=offset, precision="fp32")
    ov, oi = oracle.search(q, g, k, metric, self_mode, offset)
    assert np.array_equal(host(i), oi), f"indices differ at {np.argwhere(host(i) != oi)[:5]}"
    assert np.array_equal(host(v), ov)
    return S._search_exact_tensor.last_unverified


@pytest.mark.parametrize("d", [1, 5, 36, 100, 1024])
@pytest.mark.parametrize("role", ["queries", "gallery"])
def test_split_is_error_free_and_laid_out_as_documented(knn, d, role):
    split_bf16x3 = importlib.import_module("b200knn.search").split_bf16x3

    rs = np.random.RandomState(d)
    x = (rs.standard_normal((77, d)) * np.exp(rs.uniform(-6, 6, (77, 1)))).astype(np.float32)
    x[3] = 0.0
    y = host(split_bf16x3(dev(x), role).float())
    dpad = (d + 7) // 8 * 8
    hi = oracle.bf16_round(x)
    lo = oracle.bf16_round(x - hi)
    assert np.all(np.abs(x - hi - lo) <= 2.0 ** -16 * np.abs(x))      # two bf16 parts carry >= 16 significand bits
    parts = (hi, lo, hi) if role == "queries" else (hi, hi, lo)
    for p, want in enumerate(parts):
        assert np.array_equal(y[:, p * dpad:p * dpad + d], want)
        assert not y[:, p * dpad + d:(p + 1) * dpad].any()


@pytest.mark.parametrize("nq,ng,d,k,metric", [
    (300, 5000, 64, 10, "cosine"), (300, 5000, 64, 50, "l2"), (129, 3000, 100, 100, "ip"), (64, 4000, 36, 128, "l2"),
    (700, 20000, 1024, 50, "cosine"), (40, 900, 256, 200, "cosine"), (5, 40, 8, 10, "l2"), (3, 7, 16, 10, "ip"),
    (1, 2000, 768, 1, "cosine"),
])
def test_tensor_engine_is_bit_identical_to_the_oracle(knn, tensor_engine, nq, ng, d, k, metric):
    rs = np.random.RandomState(nq + ng + d)
    g = rs.standard_normal((ng, d)).astype(np.float32)
    q = rs.standard_normal((nq, d)).astype(np.float32)
    if metric == "cosine":
        g, q = oracle.normalize(g), oracle.normalize(q)
    else:  # un-normalised rows with very different norms: the error bound scales with |q| * max|g|
        g *= np.exp(rs.uniform(-2, 2, (ng, 1))).astype(np.float32)
        q *= np.exp(rs.uniform(-2, 2, (nq, 1))).astype(np.float32)
    _check(knn, q, g, k, metric)


def test_large_batch_filter_seeds_from_chunk_maxima(knn, tensor_engine):
    """A large batch over a large gallery: the split filter's pre-pass collects chunk maxima (select.cuh:
    seed_tile_tmem) like the plain bf16 kernel's.  Bit-identical to the oracle on a slice of the queries."""
    import ctypes as C

    from b200knn import _lib

    nq, ng, d, k = 1280, 820_000, 64, 50
    out = (C.c_int64 * 8)()
    assert _lib.load().knn_search_geometry(nq, ng, 3 * d, 2, 64, out) == 0 and out[7] >= 1, list(out)   # kc = 64 rows
    rs = np.random.RandomState(17)
    g = oracle.normalize(rs.standard_normal((ng, d)).astype(np.float32))
    q = oracle.normalize(g[:nq] + 0.5 * rs.standard_normal((nq, d)).astype(np.float32))
    for metric in ("cosine", "l2"):
        v, i = knn.search(dev(q), dev(g), k, metric, precision="fp32")
        ov, oi = oracle.search(q[100:132], g, k, metric)
        assert np.array_equal(host(i)[100:132], oi) and np.array_equal(host(v)[100:132], ov), metric


def test_self_modes_offsets_and_clustered_data(knn, tensor_engine):
    x, _ = synth.clustered(2000, 128, 3, seed=7, noise=0.8)
    e = oracle.normalize(x)
    _check(knn, e, e, 10, "cosine", "exclude")
    _check(knn, e, e, 10, "l2", "exclude")
    _check(knn, e, e, 50, "cosine", "minus1")
    _check(knn, e[500:900], e, 20, "cosine", "exclude", 500)          # a chunk of the gallery queries itself
    _check(knn, e[1900:], e[:1950], 20, "l2", "exclude", 1900)        # self rows partly outside the gallery


def test_ties_and_mass_duplicates_take_the_ffma_rerun(knn, tensor_engine):
    """Exactly-representable rows with duplicates: ties are real, the completeness proof must refuse whenever a tie
    group straddles the candidate set, and the flagged blocks must come back exact from the FFMA engine."""
    x = synth.exact_grid(3000, 96, 3, 64)
    _check(knn, x, x, 10, "ip", "exclude")
    rs = np.random.RandomState(8)
    base = oracle.normalize(rs.standard_normal((40, 64)).astype(np.float32))
    g = np.repeat(base, 300, axis=0)                                   # 300-fold duplicates: 300 > kc for k = 100
    q = oracle.normalize(rs.standard_normal((260, 64)).astype(np.float32))
    assert _check(knn, q, g, 100, "cosine") > 0
    assert _check(knn, q, g, 100, "l2") > 0
    same = np.repeat(base[:1], 5000, axis=0)                           # everything ties: rows 0..k-1 win
    _check(knn, q[:3], same, 100, "cosine")
    # only SOME query blocks are flagged: queries of block 1 sit on a duplicated row, the others see distinct rows
    g2 = oracle.normalize(rs.standard_normal((6000, 64)).astype(np.float32))
    g2[1000:1200] = g2[999]
    q2 = oracle.normalize(rs.standard_normal((640, 64)).astype(np.float32))
    q2[130:140] = g2[999]
    n_bad = _check(knn, q2, g2, 100, "cosine")
    assert 10 <= n_bad < 640


def test_observed_filter_error_is_far_inside_the_bound(knn):
    """The one assumption behind the completeness proof is the tensor cores' accumulation error (2^-21 of the magnitude
    sum per K=16 step).  Measure |filter value - exact value| on the worst case for accumulation -- all-positive rows,
    no cancellation -- and on Gaussian rows: it must stay below a QUARTER of the bound."""
    S = importlib.import_module("b200knn.search")

    rs = np.random.RandomState(21)
    for d, positive in ((1024, True), (1024, False), (2048, True), (96, True), (1024, "tiny-tail"), (4096, "ulp-tail")):
        g = rs.standard_normal((20000, d)).astype(np.float32)
        q = rs.standard_normal((256, d)).astype(np.float32)
        if positive == "tiny-tail":
            # one unit term followed by d - 1 products just below half an ulp of the running sum: the fp32 chain
            # drops every one of them, a wider accumulator keeps them -- the largest chain-vs-filter gap by design
            g = np.full((20000, d), 2.0 ** -12, dtype=np.float32) * rs.uniform(0.9, 1.0, (20000, d)).astype(np.float32)
            q = np.full((256, d), 2.0 ** -12, dtype=np.float32) * rs.uniform(0.9, 1.0, (256, d)).astype(np.float32)
            g[:, 0], q[:, 0] = 1.0, 1.0
        elif positive == "ulp-tail":
            # one unit term, then d - 1 products just BELOW one ulp of the running sum: the chain (round to nearest)
            # keeps every one, an accumulator that truncated each product to the sum's ulp would keep none -- the
            # worst case for the tensor cores' side of the bound
            q = np.full((256, d), 2.0 ** -11, dtype=np.float32)
            g = (0.99 * 2.0 ** -12 * rs.uniform(0.98, 1.0, (20000, d))).astype(np.float32)
            g[:, 0], q[:, 0] = 1.0, 1.0
        elif positive:
            g, q = np.abs(g), np.abs(q)
        gd, qd = dev(g), dev(q)
        kc = 64
        av, ai = S._search_prepared(S.split_bf16x3(qd, "queries"), None, S.split_bf16x3(gd, "gallery"), None, kc,
                                    "ip", "keep", 0, 0)
        exact = knn.scores_dense(qd, gd, "ip")                         # the fp32 chain of the exact mode
        ev = torch.gather(exact, 1, ai)
        eps = S.filter_error_bound(S.row_sqnorm(qd), S.ExactFilterRows.build(gd, None).max_sqnorm, d, "ip")
        ratio = ((av - ev).abs() / eps[:, None]).max().item()
        assert ratio < 0.25, (d, positive, ratio)


def test_flat_index_caches_the_split_rows_and_shards_merge_exactly(knn, tensor_engine):
    rs = np.random.RandomState(5)
    g = oracle.normalize(rs.standard_normal((9000, 128)).astype(np.float32))
    q = oracle.normalize(rs.standard_normal((300, 128)).astype(np.float32))
    ov, oi = oracle.search(q, g, 50, "l2")
    index = knn.FlatIndex(128, "l2", "fp32").add(dev(g[:4000])).add(dev(g[4000:]))
    for _ in range(2):
        v, i = index.search(dev(q), 50)
        assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov)
    assert index._filter is not None and index._filter.split.shape == (9000, 3 * 128)
    parts = [knn.FlatIndex(128, "l2", "fp32", index_base=s).add(dev(g[s:e])).search(dev(q), 50)
             for s, e in ((0, 3000), (3000, 9000))]
    mv, mi = knn.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), "l2")
    assert np.array_equal(host(mi), oi) and np.array_equal(host(mv), ov)


def test_c3_full_size_tensor_engine_equals_ffma_engine(knn, monkeypatch):
    """BASELINE config 3 at full size (25 000 x 112 000 x 1024, top-50, cosine): the default engine for this size is
    the tensor-core one; it must reproduce the FFMA engine bit for bit."""
    S = importlib.import_module("b200knn.search")

    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    g = knn.normalize(torch.randn((112_000, 1024), generator=gen, device="cuda"))
    q = knn.normalize(torch.randn((25_000, 1024), generator=gen, device="cuda"))
    assert S.exact_engine(25_000, 112_000, 1024, 50) == "tensor"
    v, i = knn.search(q, g, 50, "cosine", precision="fp32")
    assert S._search_exact_tensor.last_unverified < 25          # Gaussian rows: (almost) every query is proven
    monkeypatch.setenv("KNN_EXACT_ENGINE", "ffma")
    fv, fi = knn.search(q, g, 50, "cosine", precision="fp32")
    assert torch.equal(i, fi) and torch.equal(v, fv)


def test_c3_sized_gallery_tensor_engine_against_the_oracle(knn, tensor_engine):
    """The tensor-core engine FORCED at BASELINE config 3's gallery size (112 000 x 1024, top-50) against the CPU oracle
    directly -- not through the FFMA engine: 128 queries, indices and distances bit for bit, cosine and L2."""
    S = importlib.import_module("b200knn.search")
    gen = torch.Generator(device="cuda").manual_seed(31)
    g = knn.normalize(torch.randn((112_000, 1024), generator=gen, device="cuda"))
    q = knn.normalize(torch.randn((128, 1024), generator=gen, device="cuda"))
    gh, qh = host(g), host(q)
    for metric in ("cosine", "l2"):
        assert S.exact_engine(128, 112_000, 1024, 50) == "tensor"
        v, i = knn.search(q, g, 50, metric, precision="fp32")
        ov, oi = oracle.search(qh, gh, 50, metric, "keep", 0)
        assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov), metric
        assert S._search_exact_tensor.last_unverified < 8


@pytest.mark.parametrize("d", [4096, 1024])
def test_all_positive_rows_stress_the_accumulation_bound(knn, tensor_engine, d):
    """The proof's one hardware assumption is the accumulation error of the tensor cores.  All-positive rows are the
    adversarial case: every product has the same sign, the fp32 accumulator grows monotonically to ~d/4 and every
    truncation goes the same way.  The result must still be the oracle's, bit for bit -- through the proof or, for a
    query whose candidates contradict the error model (the run-time check of knn_rescore_exact), through the FFMA re-run."""
    S = importlib.import_module("b200knn.search")
    rs = np.random.RandomState(d)
    g = rs.rand(24_000, d).astype(np.float32)                 # uniform [0, 1): un-normalised, inner product
    q = rs.rand(128, d).astype(np.float32)
    g[77] = g[19_000]                                          # an exact tie
    v, i = knn.search(dev(q), dev(g), 50, "ip", precision="fp32")
    ov, oi = oracle.search(q, g, 50, "ip", "keep", 0)
    assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov)
    # the observed filter error stays inside the bound on this data too (the bound scales with |q| |g| ~ d / 3)
    eps = host(S.filter_error_bound(knn.row_sqnorm(dev(q)), S.ExactFilterRows.build(dev(g), None).max_sqnorm, d, "ip"))
    fv, fi = S._search_prepared(S.split_bf16x3(dev(q), "queries"), None, S.split_bf16x3(dev(g), "gallery"), None, 64,
                                "ip", "keep", 0, 0, split_rows=True)
    exact = (q.astype(np.float64)[:, None, :] * g.astype(np.float64)[host(fi)]).sum(-1)
    err = np.abs(host(fv).astype(np.float64) - exact).max(axis=1)
    print(f"\nd={d}: max |filter - exact| / eps = {(err / eps).max():.3f}")
    assert (err <= eps).all()


# ------------------------------------------------------------------------------------------ two-product filter
@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_two_product_filter_is_bit_identical_and_reruns_what_it_cannot_prove(knn, monkeypatch, metric):
    """KNN_BF16X2 (q_hi.g_hi + q_lo.g_hi, bound |q| max|g_lo|): same bits as the three-product filter and the FFMA
    engine.  Gaussian rows verify almost everywhere; rows with tight clusters (score gaps below the wide bound) must
    come back through the gathered three-product re-run."""
    S = importlib.import_module("b200knn.search")

    monkeypatch.setenv("KNN_EXACT_ENGINE", "tensor")
    gen = torch.Generator(device="cuda").manual_seed(31)
    g = knn.normalize(torch.randn((50_000, 256), generator=gen, device="cuda"))
    q = knn.normalize(torch.randn((2100, 256), generator=gen, device="cuda"))
    # 300 queries sit inside tight clusters of 200 gallery rows (more than the kc = 128 candidates of k = 50): the proof
    # cannot separate the cluster's members under the wide bound
    centre = g[:12].repeat_interleave(200, 0)
    g[10_000:12_400] = knn.normalize(centre + 2e-3 * torch.randn(centre.shape, generator=gen, device="cuda"))
    q[500:800] = knn.normalize(g[10_000:10_300] + 1e-3 * torch.randn((300, 256), generator=gen, device="cuda"))
    res = {}
    for products in ("2", "3", "ffma"):
        if products == "ffma":
            monkeypatch.setenv("KNN_EXACT_ENGINE", "ffma")
        else:
            monkeypatch.setenv("KNN_EXACT_PRODUCTS", products)
        S._search_exact_tensor.last_two_product_rerun = -1
        res[products] = knn.search(q, g, 50, metric, precision="fp32")
        if products == "2":
            rerun = S._search_exact_tensor.last_two_product_rerun
            assert 100 <= rerun < 1500, rerun          # the clustered queries, not everything
        if products == "3":
            assert S._search_exact_tensor.last_two_product_rerun == -1
    for products in ("2", "3"):
        assert torch.equal(res[products][1], res["ffma"][1]) and torch.equal(res[products][0], res["ffma"][0])
    ov, oi = oracle.search(host(q[490:510]), host(g), 50, metric, "keep", 0)
    assert np.array_equal(host(res["2"][1][490:510]), oi) and np.array_equal(host(res["2"][0][490:510]), ov)


def test_two_product_bound_holds_and_is_about_the_lo_norm(knn):
    """|two-product filter value - exact value| must stay below the bound knn_filter_error_bound2 on Gaussian, all-positive
    and wide-dynamic-range rows; the bound is dominated by |q| max|g_lo| (a row's |g_lo| is 0.3 - 0.6 of 2^-8 |g|)."""
    S = importlib.import_module("b200knn.search")
    L = importlib.import_module("b200knn._lib")

    rs = np.random.RandomState(5)
    for d, kind in ((1024, "gauss"), (1024, "positive"), (96, "range"), (2048, "positive")):
        g = rs.standard_normal((20000, d)).astype(np.float32)
        q = rs.standard_normal((1024, d)).astype(np.float32)
        if kind == "positive":
            g, q = np.abs(g), np.abs(q)
        if kind == "range":
            g *= np.exp(rs.uniform(-4, 4, g.shape)).astype(np.float32)
        gd, qd = dev(g), dev(q)
        filt = S.ExactFilterRows.build(gd, None)
        q3 = S.split_bf16x3(qd, "queries")
        kc = 64
        lib = L.load()
        av = torch.empty((1024, kc), dtype=torch.float32, device="cuda")
        ai = torch.empty((1024, kc), dtype=torch.int64, device="cuda")
        nbytes = lib.knn_search_workspace(1024, 20000, q3.shape[1], L.KNN_BF16X2, kc)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
        rc = lib.knn_search(q3.data_ptr(), filt.split.data_ptr(), None, None, 1024, 20000, q3.shape[1], L.KNN_BF16X2, kc,
                            1, 0, 0, 0, av.data_ptr(), ai.data_ptr(), ws.data_ptr(), ws.numel(),
                            torch.cuda.current_stream().cuda_stream)
        L.check(rc, "knn_search")
        exact = knn.scores_dense(qd, gd, "ip")
        ev = torch.gather(exact, 1, ai)
        eps2 = S.filter_error_bound(S.row_sqnorm(qd), filt.max_sqnorm, d, "ip", filt.lo_max_sqnorm)
        eps3 = S.filter_error_bound(S.row_sqnorm(qd), filt.max_sqnorm, d, "ip")
        ratio = ((av - ev).abs() / eps2[:, None]).max().item()
        assert ratio < 0.5, (d, kind, ratio)
        if kind == "gauss":   # the dropped product is really there: outside the three-product bound somewhere
            assert ((av - ev).abs() > eps3[:, None]).any(), (d, kind)
        lo = float(filt.lo_max_sqnorm.item()) ** 0.5 / float(filt.max_sqnorm.item()) ** 0.5
        assert 2.0 ** -10 < lo < 2.0 ** -8, lo
