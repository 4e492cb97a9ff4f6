"""GPU parity tests (B200): the CUDA path, called through the Python API -> C ABI, against the CPU oracle on the
same seeded inputs, against the golden vectors of the real reference, and through size-independent properties.

Bars: fp32 mode -- indices AND distances bit-exact against the oracle (same fp32 summation order), tie-aware
equal against the reference's MKL-ordered results with distances within 1e-5 relative; bf16 mode -- bit-exact
on exactly-representable data, recall@k >= 0.999 otherwise.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import synth
from util import rel_close, tie_aware_mismatches, ulp_diff

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


# ------------------------------------------------------------------------------------------ normalise
@pytest.mark.parametrize("d", [1, 3, 36, 64, 100, 768, 1024, 2050])
@pytest.mark.parametrize("mode", ["clamp", "none", "add"])
def test_normalize_fp32_bit_exact(knn, d, mode):
    rs = np.random.RandomState(d)
    x = rs.standard_normal((77, d)).astype(np.float32)
    x[5] = 0.0
    eps = 1e-12 if mode != "add" else 1e-8
    y, sq = knn.normalize(dev(x), eps=eps, eps_mode=mode, return_sqnorm=True)
    with np.errstate(all="ignore"):
        oy, osq = oracle.normalize(x, eps, mode, return_sqnorm=True)
    assert np.array_equal(host(y), oy, equal_nan=True)
    assert np.array_equal(host(sq), osq, equal_nan=True)


def test_normalize_matches_reference_F_normalize(knn, golden, golden_arrays):
    c = golden["cases"]["c1"]
    x, _ = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    y = host(knn.normalize(dev(x)))
    assert ulp_diff(y[:8], golden_arrays["c1_normalized_rows_0_8"]).max() <= 2


@pytest.mark.parametrize("d", [8, 64, 768])
def test_normalize_bf16_and_cast(knn, d):
    rs = np.random.RandomState(3)
    x = rs.standard_normal((50, d)).astype(np.float32)
    y, sq = knn.normalize(dev(x), out_dtype=torch.bfloat16, return_sqnorm=True)
    oy, osq = oracle.normalize(x, to_bf16=True, return_sqnorm=True)
    assert np.array_equal(host(y.float()), oy) and np.array_equal(host(sq), osq)
    c = knn.normalize(dev(x), eps_mode="cast", out_dtype=torch.bfloat16)
    assert np.array_equal(host(c.float()), oracle.bf16_round(x))
    assert np.array_equal(host(knn.row_sqnorm(y)), osq)


# ------------------------------------------------------------------------------------------ fp32 search
def _check_exact(knn, q, g, k, metric, self_mode="keep", offset=0, precision="fp32"):
    v, i = knn.search(dev(q), dev(g), k, metric, self_mode=self_mode, query_offset=offset, precision=precision)
    ov, oi = oracle.search(q, g, k, metric, self_mode, offset)
    assert np.array_equal(host(i), oi), f"indices differ at {np.argwhere(host(i) != oi)[:5]}"
    assert np.array_equal(host(v), ov)
    return host(v), host(i)


def test_c1_cosine_bit_exact_vs_oracle_and_tie_aware_vs_reference(knn, golden, golden_arrays):
    c = golden["cases"]["c1"]
    x, _ = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    e = host(knn.normalize(dev(x)))
    v, i = _check_exact(knn, e, e, 10, "cosine", "exclude")
    tol = 4 * 2.0**-23 * 32
    ndiff, bad = tie_aware_mismatches(v, i, golden_arrays["c1_cosine_top10_val"], golden_arrays["c1_cosine_top10_idx"], tol)
    assert bad == 0
    assert rel_close(v, golden_arrays["c1_cosine_top10_val"], 1e-5, 1e-6)
    v2, i2 = _check_exact(knn, e, e, 10, "l2", "exclude")
    _, bad = tie_aware_mismatches(v2, i2, golden_arrays["c1_cdist_top10_val"], golden_arrays["c1_cdist_top10_idx"], 1e-5)
    assert bad == 0 and rel_close(v2, golden_arrays["c1_cdist_top10_val"], 1e-5, 1e-6)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_exact_grid_matches_the_reference_bit_for_bit(knn, golden, golden_arrays, metric, precision):
    """Exactly representable data: any summation order gives the same fp32 scores, ties are real (duplicated
    rows) -> indices must equal the reference's stable ranking exactly, in BOTH precisions."""
    c = golden["cases"]["c1_exact"]
    x = synth.exact_grid(c["n"], c["d"], c["seed"], c["n_dup"])
    v, i = knn.search(dev(x), dev(x), 10, metric, exclude_self=True, precision=precision)
    assert np.array_equal(host(i), golden_arrays[f"c1_exact_{metric}_top10_idx"])
    if metric == "ip":
        assert np.array_equal(host(v), golden_arrays[f"c1_exact_{metric}_top10_val"])
    else:
        assert ulp_diff(host(v), golden_arrays[f"c1_exact_{metric}_top10_val"]).max() <= 1  # torch sqrt, see oracle test


def test_c2_l2_query_gallery(knn, golden, golden_arrays):
    c = golden["cases"]["c2"]
    x, _ = synth.clustered(c["nq"] + c["ng"], c["d"], c["classes"], c["seed"], c["noise"], c["priors"])
    e = host(knn.normalize(dev(x)))
    v, i = _check_exact(knn, e[: c["nq"]], e[c["nq"]:], 10, "l2")
    _, bad = tie_aware_mismatches(v, i, golden_arrays["c2_top10_val"], golden_arrays["c2_top10_idx"], 1e-5)
    assert bad == 0 and rel_close(v, golden_arrays["c2_top10_val"], 1e-5, 1e-6)


@pytest.mark.parametrize("nq,ng,d,k", [
    (1, 1, 8, 1), (1, 100, 36, 5), (7, 127, 64, 32), (128, 128, 64, 33), (129, 129, 100, 50),
    (300, 1000, 96, 100), (33, 5000, 48, 128), (40, 3000, 32, 200), (17, 2000, 16, 256), (5, 3, 8, 10),
])
@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_shape_and_k_sweep_fp32(knn, nq, ng, d, k, metric):
    rs = np.random.RandomState(nq * 7 + ng)
    g = oracle.normalize(rs.standard_normal((ng, d)).astype(np.float32))
    q = oracle.normalize(rs.standard_normal((nq, d)).astype(np.float32))
    v, i = _check_exact(knn, q, g, k, metric)
    if k > ng:  # ragged: fewer candidates than k -> (-inf | +inf, -1) tail
        assert (i[:, ng:] == -1).all() and np.isinf(v[:, ng:]).all()


def test_k_beyond_fused_limit_goes_dense(knn):
    rs = np.random.RandomState(9)
    e = oracle.normalize(rs.standard_normal((400, 64)).astype(np.float32))
    _check_exact(knn, e, e, 399, "cosine", "exclude")  # train.py:409 topk(N-1)
    _check_exact(knn, e, e, 400, "cosine", "minus1")   # nih_multilabel_training.py:86


def test_empty_inputs(knn):
    q = torch.zeros((0, 16), device="cuda")
    g = torch.randn((10, 16), device="cuda")
    v, i = knn.search(q, g, 3)
    assert v.shape == (0, 3) and i.shape == (0, 3)
    v, i = knn.search(g, q, 3)
    assert (host(i) == -1).all() and np.isneginf(host(v)).all()


def test_adversarial_orders_and_mass_ties(knn):
    rs = np.random.RandomState(4)
    d, ng = 32, 6000
    qv = oracle.normalize(rs.standard_normal((3, d)).astype(np.float32))
    g = oracle.normalize(rs.standard_normal((ng, d)).astype(np.float32))
    order = np.argsort(g @ qv[0])  # ascending score for query 0: every new row beats the threshold
    _check_exact(knn, qv, g[order], 100, "cosine")
    _check_exact(knn, qv, g[order[::-1]], 100, "cosine")
    same = np.repeat(g[:1], ng, axis=0)  # all scores tie: order must be gallery row 0..k-1
    v, i = _check_exact(knn, qv, same, 100, "cosine")
    assert np.array_equal(i[0], np.arange(100))
    blocks = np.repeat(g[:60], 100, axis=0)  # 100-fold duplicates
    _check_exact(knn, qv, blocks, 128, "l2")


def test_list_compaction_falls_back_to_the_sort_without_duplicates(knn):
    """Regression (found by the shard-invariance property at 8192 x 2M): when the selection-based list compaction
    makes too little progress it has already rewritten the list, and the sorting fallback must read the NEW count.
    The first unit's list is filled in gallery order (fp32 kernel, d = 1), so the 32 keys the compaction samples
    (position (l % 8) * 32 + l for lane l, list capacity 256 for k = 100) can be made the 32 worst of the first
    256 rows: the pivot keeps 225 of 256 entries -> fallback.  Rows 225..255 hold the best scores, so stale copies
    of them would surface as duplicate indices in the top-k.  8192 queries x 10 000 rows: 5 gallery splits of 2048
    rows (the first starts at row 0) and no threshold-seeding pre-pass, so unit 0 really accepts its first 256 rows."""
    nq, ng, k = 8192, 10_000, 100
    rs = np.random.RandomState(11)
    score = rs.permutation(ng).astype(np.float32)            # distinct integers: exact in fp32
    head = np.sort(score[:256].copy())
    sampled = sorted({(l % 8) * 32 + l for l in range(32)})
    rest = [p for p in range(256) if p not in sampled]
    score[sampled] = head[:32]                               # the sample = the 32 lowest of the first list
    score[rest] = head[32:]                                  # ascending: positions 225..255 are the best of the list
    score[:256] += 2.0 * ng                                  # ... and all of them belong to the final top-k region
    g = (score / 2.0 ** 20).reshape(ng, 1).astype(np.float32)
    qv = (1.0 + (np.arange(nq) % 7)).reshape(nq, 1).astype(np.float32) / 4.0
    v, i = _check_exact(knn, qv, g, k, "ip")
    assert all(len(set(r)) == k for r in i[::257].tolist())


@pytest.mark.parametrize("precision,nq,d", [("fp32", 5, 2), ("fp32", 300, 2), ("bf16", 5, 8), ("bf16", 300, 8),
                                            ("bf16", 5, 1024)])
def test_every_row_beats_the_threshold_for_every_list_geometry(knn, precision, nq, d):
    """Regression: with the shortest lists (k <= 32: capacity 64) the selection-based compaction left more than
    capacity - 32 keys, and the next chunk's hits were written past the end of the list into the neighbouring row's
    (lost candidates; bf16 at k = 32 missed a neighbour in 10 % of the rows).  Worst case for every list geometry:
    scores ascend with the gallery row, so EVERY new row beats the threshold.  score_j = j / 65536 exactly, built
    from bf16-representable parts so both precisions are bit-exact; negative queries see the descending order."""
    ng = 60_000
    j = np.arange(ng)
    g = np.zeros((ng, d), dtype=np.float32)
    g[:, 0] = (j // 256) / 256.0
    g[:, 1] = (j % 256).astype(np.float32)
    qv = np.zeros((nq, d), dtype=np.float32)
    scale = (2.0 ** (np.arange(nq) % 5)) * np.where(np.arange(nq) % 3 == 2, -1.0, 1.0)
    qv[:, 0] = scale
    qv[:, 1] = scale * 2.0 ** -16
    for k in (1, 10, 20, 31, 32, 33, 64, 100, 128, 256):
        _check_exact(knn, qv, g, k, "ip", precision=precision)


def test_self_modes_with_offset(knn):
    rs = np.random.RandomState(5)
    g = oracle.normalize(rs.standard_normal((700, 40)).astype(np.float32))
    q = g[300:420].copy()
    v, i = _check_exact(knn, q, g, 10, "cosine", "exclude", 300)
    assert not (i == (np.arange(120) + 300)[:, None]).any()
    _check_exact(knn, q, g, 10, "cosine", "minus1", 300)
    v, i = _check_exact(knn, q, g, 1, "cosine", "keep", 300)
    assert np.array_equal(i[:, 0], np.arange(120) + 300)


def test_nih_sized_gallery_subsample_vs_oracle(knn):
    """BASELINE config 3 gallery size (112k x 1024) against the oracle for a 96-query subsample, top-50."""
    rs = np.random.RandomState(3)
    ng, d = 112_000, 1024
    g = knn.normalize(torch.from_numpy(rs.standard_normal((ng, d)).astype(np.float32)).cuda())
    q = g[rs.choice(ng, 96, replace=False)] + 0.25 * torch.from_numpy(rs.standard_normal((96, d)).astype(np.float32)).cuda()
    q = knn.normalize(q)
    v, i = knn.search(q, g, 50, "cosine")
    ov, oi = oracle.search(host(q), host(g), 50, "cosine")
    assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov)


# ------------------------------------------------------------------------------------------ bf16 tensor-core path
@pytest.mark.parametrize("nq,ng,d,k", [(1, 256, 64, 10), (64, 5000, 768, 100), (200, 9000, 512, 100), (130, 777, 128, 50),
                                       (8, 20000, 72, 32),
                                       # several splits + threshold-seeding pre-pass (ng >= 32768), CTA pairs
                                       (300, 34000, 256, 100), (140, 33000, 768, 10),
                                       # d > 768: query tile does not fit TMEM -> shared-memory-A kernels (1 and 2 CTAs)
                                       (70, 3000, 1024, 20), (300, 3000, 1024, 20),
                                       # d not a multiple of the 64-element k-block, k at the fused limit; tiny d
                                       (33, 2000, 200, 256), (129, 1500, 8, 5)])
@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_bf16_recall_and_scores(knn, nq, ng, d, k, metric):
    rs = np.random.RandomState(nq + ng)
    g = oracle.normalize(rs.standard_normal((ng, d)).astype(np.float32), to_bf16=True)
    q = oracle.normalize(rs.standard_normal((nq, d)).astype(np.float32), to_bf16=True)
    v, i = knn.search(dev(q), dev(g), k, metric, precision="bf16")
    v, i = host(v), host(i)
    kk = min(k, ng)
    ov, oi = oracle.search(q, g, k, metric)  # same bf16-rounded inputs, fp32 fmaf-chain accumulation
    recall = np.mean([len(set(i[r, :kk]) & set(oi[r, :kk])) / kk for r in range(nq)])
    assert recall >= 0.999, recall
    assert np.allclose(v[:, :kk], ov[:, :kk], rtol=1e-4, atol=2e-5)  # fp32 accumulate: only the order differs
    assert np.all(np.diff(v[:, :kk], axis=1) * (1 if metric == "l2" else -1) >= 0)  # sorted best-first
    # returned scores are the scores OF the returned rows
    rescored = oracle.scores(q, g, metric)
    assert np.allclose(v[:, :kk], np.take_along_axis(rescored, i[:, :kk], axis=1), rtol=1e-4, atol=2e-5)


def test_bf16_recall_against_fp32_mode(knn):
    """north star: recall@k of bf16 mode >= 0.999 against the fp32 engine on the same (bf16-representable) rows."""
    rs = np.random.RandomState(11)
    g = dev(oracle.normalize(rs.standard_normal((30000, 768)).astype(np.float32), to_bf16=True))
    q = dev(oracle.normalize(rs.standard_normal((64, 768)).astype(np.float32), to_bf16=True))
    _, i32 = knn.search(q, g, 100, "cosine", precision="fp32")
    _, i16 = knn.search(q, g, 100, "cosine", precision="bf16")
    i32, i16 = host(i32), host(i16)
    recall = np.mean([len(set(i32[r]) & set(i16[r])) / 100 for r in range(64)])
    assert recall >= 0.999, recall


def test_bf16_exact_grid_adversarial_and_self(knn):
    x = synth.exact_grid(3000, 256, 21, 64)
    v, i = knn.search(dev(x), dev(x), 100, "ip", exclude_self=True, precision="bf16")
    ov, oi = oracle.search(x, x, 100, "ip", "exclude", 0)
    assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov)
    same = np.repeat(x[:1], 4000, axis=0)
    v, i = knn.search(dev(x[:5]), dev(same), 64, "ip", precision="bf16")
    assert np.array_equal(host(i)[0], np.arange(64))


def _geometry(knn, nq, ng, d, dtype, k):
    import ctypes as C

    from b200knn import _lib

    out = (C.c_int64 * 8)()
    assert _lib.load().knn_search_geometry(nq, ng, d, dtype, k, out) == 0
    return dict(zip(("qblocks", "splits", "groups", "split_len", "L", "seed_units", "seed_len", "seed_stride"), out))


@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_threshold_seeding_from_chunk_maxima_is_exact(knn, metric):
    """Large query batches on the CTA-pair kernel seed their thresholds from the chunk maxima of a gallery sample
    (select.cuh: seed_tile_tmem): the k-th largest maximum must be a LOWER bound of every query's final k-th best
    score, or neighbours are lost.  Exactly-representable rows with real ties (thousands of equal scores around every
    cut-off), the queries ARE rows of the sampled range (self exclusion inside the sample), every list geometry."""
    ng, nq, d = 820_000, 1280, 64
    x = synth.exact_grid(ng, d, 7, 64)
    for k in (10, 100, 256):
        geo = _geometry(knn, nq, ng, d, 1, k)
        assert geo["seed_stride"] >= 1 and geo["seed_units"] * geo["seed_len"] >= 1280, geo   # the mode under test
        v, i = knn.search(dev(x[:nq]), dev(x), k, metric, self_mode="exclude", precision="bf16")
        v, i = host(v), host(i)
        for lo in (0, 1000):
            ov, oi = oracle.search(x[lo:lo + 24], x, k, metric, "exclude", lo)
            assert np.array_equal(i[lo:lo + 24], oi), (k, lo, np.argwhere(i[lo:lo + 24] != oi)[:5])
            assert np.array_equal(v[lo:lo + 24], ov)
    # keep mode, queries that are not gallery rows, descending-by-row scores inside the sample (every chunk maximum
    # sits in the chunk's first column) and an all-equal stretch
    g2 = x.copy()
    g2[:4096, 0] = (255 - np.arange(4096) // 16).astype(np.float32) / 256.0      # 8 significant bits: exact in bf16
    g2[4096:6144] = g2[4096]
    q2 = synth.exact_grid(nq, d, 8, 0)
    q2[:, 0] = 0.5
    v, i = knn.search(dev(q2), dev(g2), 100, metric, precision="bf16")
    ov, oi = oracle.search(q2[600:632], g2, 100, metric)
    assert np.array_equal(host(i)[600:632], oi) and np.array_equal(host(v)[600:632], ov)


def test_small_batch_seeding_keeps_one_maximum_per_unit(knn):
    """One query block on the TMEM-resident kernel: the pre-pass keeps a single running maximum per (unit, selection
    thread) -- all the list-maxima seeding ever reads.  Bit-exact with real ties, self rows inside the sample."""
    ng, nq, d = 3_400_000, 16, 64
    geo = _geometry(knn, nq, ng, d, 1, 100)
    assert geo["seed_stride"] == 1 << 30 and geo["qblocks"] == 1, geo
    x = synth.exact_grid(ng, d, 9, 64)
    for metric in ("ip", "l2"):
        v, i = knn.search(dev(x[:nq]), dev(x), 100, metric, self_mode="exclude", precision="bf16")
        ov, oi = oracle.search(x[:nq], x, 100, metric, "exclude", 0)
        assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov), metric


# ------------------------------------------------------------------------------------------ dense / rank / merge
@pytest.mark.parametrize("metric,self_mode", [("cosine", "exclude"), ("ip", "minus1"), ("l2", "exclude"), ("l2", "keep")])
def test_scores_dense_bit_exact(knn, metric, self_mode):
    rs = np.random.RandomState(2)
    g = oracle.normalize(rs.standard_normal((333, 50)).astype(np.float32))
    q = g[100:230].copy()
    s = knn.scores_dense(dev(q), dev(g), metric, self_mode=self_mode, query_offset=100)
    assert np.array_equal(host(s), oracle.scores(q, g, metric, self_mode, 100))


@pytest.mark.parametrize("n", [1, 2, 100, 4096, 4097, 10000, 20000])
def test_rank_rows_stable(knn, n):
    rs = np.random.RandomState(n)
    s = np.round(rs.standard_normal((5, n)), 2).astype(np.float32)  # rounding -> many exact ties
    for largest in (True, False):
        r = host(knn.rank_rows(dev(s), largest_first=largest))
        assert np.array_equal(r, oracle.rank_rows(s, largest))


@pytest.mark.parametrize("parts,k,metric", [(2, 10, "cosine"), (8, 100, "cosine"), (4, 50, "l2"), (3, 1, "ip"), (8, 256, "l2")])
def test_merge_topk(knn, parts, k, metric):
    rs = np.random.RandomState(parts * k)
    nq = 37
    vals = np.round(rs.standard_normal((parts, nq, k)), 1).astype(np.float32)  # ties across parts
    vals = np.sort(vals, axis=2)
    if metric != "l2":
        vals = vals[:, :, ::-1].copy()
    idx = np.stack([np.sort(rs.choice(10_000, (nq, k)) + p * 10_000, axis=1) for p in range(parts)]).astype(np.int64)
    # make idx consistent with the tie order inside each list (ascending idx within equal values)
    for p in range(parts):
        for r in range(nq):
            order = np.lexsort((idx[p, r], vals[p, r] if metric == "l2" else -vals[p, r]))
            vals[p, r], idx[p, r] = vals[p, r][order], idx[p, r][order]
    idx[0, 0, k // 2:] = -1
    vals[0, 0, k // 2:] = np.inf if metric == "l2" else -np.inf
    v, i = knn.merge_topk(dev(vals), dev(idx), metric)
    ov, oi = oracle.merge_topk(vals, idx, metric)
    assert np.array_equal(host(i), oi) and np.array_equal(host(v), ov)


@pytest.mark.parametrize("precision,metric", [("fp32", "cosine"), ("bf16", "cosine"), ("fp32", "l2")])
def test_row_sharded_result_is_independent_of_shard_count(knn, precision, metric):
    """shard the gallery by contiguous rows, search every shard with global indices, merge: the result
    must equal the unsharded search bit for bit (all shards emulated on one GPU)."""
    from b200knn.sharded import shard_rows

    rs = np.random.RandomState(8)
    g = oracle.normalize(rs.standard_normal((10_000, 128)).astype(np.float32), to_bf16=True)
    g[123] = g[9000]
    q = g[:150].copy()
    gq, gg = dev(q), dev(g)
    full = knn.FlatIndex(128, metric, precision).add(gg)
    v1, i1 = full.search(gq, 100, exclude_self=True)
    for world in (2, 4, 8):
        pv, pi = [], []
        for start, count in shard_rows(g.shape[0], world):
            shard = knn.FlatIndex(128, metric, precision, index_base=start).add(gg[start:start + count])
            v, i = shard.search(gq, 100, exclude_self=True)
            pv.append(v)
            pi.append(i)
        v, i = knn.merge_topk(torch.stack(pv), torch.stack(pi), metric)
        assert torch.equal(i, i1) and torch.equal(v, v1), world


# ------------------------------------------------------------------------------------------ kernel variants by env
_VARIANT_SNIPPET = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import b200knn, oracle
from oracle import synth
x = synth.exact_grid(900, 256, 5, 32)                      # exactly representable: every kernel must be bit-exact
v, i = b200knn.search(torch.from_numpy(x).cuda(), torch.from_numpy(x).cuda(), 100, "ip", exclude_self=True, precision="bf16")
ov, oi = oracle.search(x, x, 100, "ip", "exclude", 0)
assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(v.cpu().numpy(), ov), "grid"
rs = np.random.RandomState(3)
g = oracle.normalize(rs.standard_normal((40000, {d})).astype(np.float32), to_bf16=True)
q = oracle.normalize(rs.standard_normal(({nq}, {d})).astype(np.float32), to_bf16=True)
for metric in ("cosine", "l2"):
    v, i = b200knn.search(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), 50, metric, precision="bf16")
    ov, oi = oracle.search(q, g, 50, metric)
    i = i.cpu().numpy()
    rec = np.mean([len(set(i[r]) & set(oi[r])) / 50 for r in range({nq})])
    assert rec >= 0.999, (metric, rec)
    assert np.allclose(v.cpu().numpy(), ov, rtol=1e-4, atol=2e-5), metric
print("VARIANT-OK")
"""


@pytest.mark.parametrize("env,nq,d", [
    ({"KNN_BF16_TS": "1"}, 300, 512),     # TMEM-resident-query kernel on CTA pairs, two accumulator stages
    ({"KNN_BF16_TS": "1"}, 300, 768),     # ... one accumulator stage
    ({"KNN_BF16_TS": "0"}, 64, 768),      # one-CTA shared-memory-A kernel where the TMEM-resident one is the default
    ({"KNN_SEED_ROWS": "0"}, 300, 512),   # no threshold-seeding pre-pass
    ({"KNN_WAVES": "1"}, 64, 768),        # fewest splits
])
def test_bf16_kernel_variants_selected_by_environment(env, nq, d):
    """The dispatch switches are read once per process, so every non-default kernel variant is exercised in a
    subprocess: bit-exact on exactly-representable data, recall/score bars otherwise."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", _VARIANT_SNIPPET.format(root=root, nq=nq, d=d)],
                         env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "VARIANT-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


# ------------------------------------------------------------------------------------------ full-size properties
def _gen_rows(knn, n, d, seed, dtype=torch.bfloat16):
    gen = torch.Generator(device="cuda")
    out = torch.empty((n, d), dtype=dtype, device="cuda")
    for s in range(0, n, 1 << 20):
        e = min(n, s + (1 << 20))
        gen.manual_seed(seed + s)
        out[s:e] = knn.normalize(torch.randn((e - s, d), generator=gen, device="cuda"), out_dtype=dtype)
    return out


@pytest.mark.parametrize("nq,ng,d,k", [
    (64, 10_000_000, 768, 100),      # BASELINE config 4 at full size (one GPU): TMEM-resident-query kernel
    (8192, 2_000_000, 512, 100),     # BASELINE config 5's shape on a gallery slice: CTA-pair kernel, many splits
])
def test_full_size_properties(knn, nq, ng, d, k):
    """Size-independent properties where the oracle cannot run: (1) rows sorted best-first with unique indices,
    (2) the returned scores are the scores of the returned rows (re-scored through the exact fp32 engine),
    (3) exactness on a superset: the exact top-k over {random subsample} U {returned rows} is the returned set,
    (4) shard invariance: searching two halves and merging the candidates gives the identical result."""
    g = _gen_rows(knn, ng, d, 1234)
    q = _gen_rows(knn, nq, d, 99)
    index = knn.FlatIndex(d, "cosine", "bf16").adopt(g)
    v, i = index.search(q, k)
    assert bool((v[:, :-1] >= v[:, 1:]).all()) and bool((i >= 0).all()) and bool((i < ng).all())
    assert all(len(set(r)) == k for r in host(i[:: max(1, nq // 16)]).tolist())
    rows = torch.randperm(nq, device="cuda")[:16]
    sub = torch.randint(0, ng, (200_000,), device="cuda")
    for r in rows.tolist():
        cand = torch.unique(torch.cat([sub, i[r]]))                      # sorted ascending -> tie order preserved
        gv = g[cand].float()
        ev, ei = knn.search(q[r:r + 1].float(), gv, k, "ip", precision="fp32")
        got = cand[ei[0]]
        assert set(host(got).tolist()) == set(host(i[r]).tolist()), r
        assert torch.allclose(ev[0], v[r], rtol=1e-4, atol=2e-5)         # (2): fp32 re-scoring of bf16 rows
    half = ng // 2
    parts_v, parts_i = [], []
    for s, e in ((0, half), (half, ng)):
        pv, pi = knn.FlatIndex(d, "cosine", "bf16", index_base=s).adopt(g[s:e]).search(q, k)
        parts_v.append(pv)
        parts_i.append(pi)
    mv, mi = knn.merge_topk(torch.stack(parts_v), torch.stack(parts_i), "cosine")
    assert torch.equal(mi, i) and torch.equal(mv, v)


def test_randomised_stress_against_brute_force():
    """20 s of tools/stress.py: random (precision, metric, shape, k) against torch brute force over the same rows --
    sortedness, unique in-range indices, tie-aware set equality and score profiles (the kind of check that surfaced
    the two list-compaction bugs)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "stress.py"), "--seconds", "20", "--seed", "7"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert '"failures": 0' in res.stdout


def test_search_is_cuda_graph_capturable(knn):
    """The library only enqueues work on the caller's stream (no hidden synchronisation, no allocation): a whole
    FlatIndex.search -- normalise + cast, threshold seeding, distance + selection, two-phase merge -- can be captured
    into a CUDA graph and replayed on new query contents (small-batch serving without launch overhead)."""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(12)
    g = torch.randn((300_000, 256), generator=gen, device="cuda")
    index = knn.FlatIndex(256, "cosine", "bf16", normalize=True).add(g)
    q_static = torch.randn((16, 256), generator=gen, device="cuda")
    index.search(q_static, 50)                                   # warm-up outside the capture (lazy initialisation)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        v_static, i_static = index.search(q_static, 50)
    for seed in (1, 2, 3):
        gen.manual_seed(seed)
        q_new = torch.randn((16, 256), generator=gen, device="cuda")
        q_static.copy_(q_new)
        graph.replay()
        torch.cuda.synchronize()
        v, i = index.search(q_new, 50)
        assert torch.equal(i, i_static) and torch.equal(v, v_static)
