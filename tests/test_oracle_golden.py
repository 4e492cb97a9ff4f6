"""CPU: the oracle (oracle/knn_oracle.c + oracle/reference_metrics.py) against the golden vectors that the REAL
reference produced (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import oracle
from oracle import reference_metrics as rm
from oracle import synth
from util import rel_close, tie_aware_mismatches, ulp_diff

TIE_TOL = 4 * 2.0**-23 * 32  # 4 ulp * sqrt(D=1024) on unit-scale scores


@pytest.fixture(scope="module")
def c1(golden):
    c = golden["cases"]["c1"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    return oracle.normalize(x), lab, x


def test_normalize_matches_F_normalize(c1, golden_arrays):
    e, _, _ = c1
    ref = golden_arrays["c1_normalized_rows_0_8"]
    assert ulp_diff(e[:8], ref).max() <= 2  # fp32 summation order of the norm only


def test_normalize_eps_modes():
    x = np.zeros((3, 8), dtype=np.float32)
    x[1, 0] = 3.0
    x[2] = 1e-20
    y = oracle.normalize(x, eps=1e-12, eps_mode="clamp")
    assert np.all(y[0] == 0) and y[1, 0] == 1.0 and np.all(np.isfinite(y))
    with np.errstate(all="ignore"):
        assert np.isnan(oracle.normalize(x, eps_mode="none")[0]).all()  # test.py:251 convention: 0/0
    assert np.allclose(oracle.normalize(x, eps=1e-8, eps_mode="add")[1, 0], 3.0 / (3.0 + 1e-8))


def test_c1_cosine_topk_and_metrics(c1, golden, golden_arrays):
    e, lab, _ = c1
    S = oracle.scores(e, e, "cosine", "exclude", 0)
    val, idx = oracle.topk(S, 10)
    ndiff, bad = tie_aware_mismatches(val, idx, golden_arrays["c1_cosine_top10_val"], golden_arrays["c1_cosine_top10_idx"], TIE_TOL)
    assert bad == 0
    assert rel_close(val, golden_arrays["c1_cosine_top10_val"], 1e-5, 1e-6)
    g = golden["c1_cosine"]
    assert [float(v) for v in rm.retrieval_accuracy(idx, lab, lab, (1, 5, 10))] == g["acc"]
    # (1) the metric restatements are bit-exact when fed the reference's OWN full ranking
    ref_ranks = golden_arrays["c1_cosine_ranks_rowmajor"].astype(np.int64)
    mAP, aps, pr, prs = rm.compute_map(ref_ranks, lab, lab, [1, 5, 10])
    assert mAP == g["mAP"] and list(pr) == g["pr"]
    assert np.array_equal(aps, golden_arrays["c1_cosine_aps"]) and np.array_equal(prs, golden_arrays["c1_cosine_prs"])
    cls = rm.compute_classification_metrics(ref_ranks, lab, lab, (1, 5, 10, 15, 20))
    for k, d in g["classification"].items():
        for m, v in d.items():
            assert cls[int(k)][m] == v, (k, m)
    # (2) the oracle's own ranking differs from MKL's only inside near-ties (fp32 summation order):
    #     top-k metrics identical, the full-ranking mAP within 1e-6
    ranks = oracle.rank_rows(S)
    frac = np.mean(ranks != ref_ranks)
    assert frac < 2e-3, frac
    mAP2, _, pr2, _ = rm.compute_map(ranks, lab, lab, [1, 5, 10])
    assert abs(mAP2 - g["mAP"]) < 1e-6 and list(pr2) == g["pr"]
    cls2 = rm.compute_classification_metrics(ranks, lab, lab, (1, 5, 10, 15, 20))
    for k, d in g["classification"].items():
        for m, v in d.items():
            assert cls2[int(k)][m] == v, (k, m)


def test_c1_cdist_pipeline(c1, golden, golden_arrays):
    """The reference's own evaluate() (test.py:1066-1126, -cdist) end to end."""
    e, lab, _ = c1
    D = oracle.scores(e, e, "l2", "exclude", 0)
    val, idx = oracle.topk(D, 10, largest_first=False)
    _, bad = tie_aware_mismatches(val, idx, golden_arrays["c1_cdist_top10_val"], golden_arrays["c1_cdist_top10_idx"], 1e-5)
    assert bad == 0
    assert rel_close(val, golden_arrays["c1_cdist_top10_val"], 1e-5, 1e-6)
    g = golden["c1_cdist"]
    assert [float(v) for v in rm.retrieval_accuracy(idx, lab, lab, (1, 5, 10))] == g["acc"]
    ref_ranks = golden_arrays["c1_cdist_ranks_rowmajor"].astype(np.int64)
    mAP, _, pr, _ = rm.compute_map(ref_ranks, lab, lab, [1, 5, 10])
    # evaluate() ranks with torch's UNSTABLE argsort and fp32 cdist values collide exactly a few times per
    # matrix, so even the reference's own full-ranking mAP is only defined up to those ties (~1e-7)
    assert abs(mAP - g["mAP"]) < 1e-6 and list(pr) == g["pr"]
    cls = rm.compute_classification_metrics(ref_ranks, lab, lab, (1, 5, 10, 15, 20))
    for k, vec in g["classification"].items():
        assert list(cls[int(k)].values()) == vec
    ranks = oracle.rank_rows(D, largest_first=False)
    mAP2, _, pr2, _ = rm.compute_map(ranks, lab, lab, [1, 5, 10])
    assert abs(mAP2 - g["mAP"]) < 1e-6 and list(pr2) == g["pr"]


@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_c1_exact_grid_is_bit_exact(golden, golden_arrays, metric):
    c = golden["cases"]["c1_exact"]
    x = synth.exact_grid(c["n"], c["d"], c["seed"], c["n_dup"])
    val, idx = oracle.search(x, x, 10, metric, "exclude", 0)
    assert np.array_equal(idx, golden_arrays[f"c1_exact_{metric}_top10_idx"])  # includes real ties
    if metric == "ip":
        assert np.array_equal(val, golden_arrays[f"c1_exact_{metric}_top10_val"])
    else:
        # d^2 is exact on this grid; torch's vectorised CPU sqrt is not correctly rounded (e.g. sqrt(1230/1024)
        # comes out 1 ulp low) while the engine/oracle use IEEE sqrt -> distances agree to 1 ulp
        assert ulp_diff(val, golden_arrays[f"c1_exact_{metric}_top10_val"]).max() <= 1


def test_c2_l2_query_gallery(golden, golden_arrays):
    c = golden["cases"]["c2"]
    x, lab = synth.clustered(c["nq"] + c["ng"], c["d"], c["classes"], c["seed"], c["noise"], c["priors"])
    e = oracle.normalize(x)
    q, g, ql, gl = e[: c["nq"]], e[c["nq"]:], lab[: c["nq"]], lab[c["nq"]:]
    val, idx = oracle.search(q, g, 10, "l2")
    _, bad = tie_aware_mismatches(val, idx, golden_arrays["c2_top10_val"], golden_arrays["c2_top10_idx"], 1e-5)
    assert bad == 0
    assert rel_close(val, golden_arrays["c2_top10_val"], 1e-5, 1e-6)
    got = rm.ath_compute_metrics(idx, ql, gl, (1, 5, 10))
    for k, d in golden["c2_l2"]["retrieval"].items():
        for m, v in d.items():
            assert got[int(k)][m] == v, (k, m)


def test_c3_small_multilabel_top50(golden, golden_arrays):
    c = golden["cases"]["c3s"]
    lab_all = synth.multihot(c["nq"] + c["ng"], c["seed"])
    emb = oracle.normalize(synth.labelset_clustered(lab_all, c["d"], c["seed"] + 100, c["noise"]))
    q, g = emb[: c["nq"]], emb[c["nq"]:]
    val, idx = oracle.search(q, g, c["k"], "cosine")
    _, bad = tie_aware_mismatches(val, idx, golden_arrays["c3s_top50_val"], golden_arrays["c3s_top50_idx"], 1e-6)
    assert bad == 0
    got = rm.evaluate_results(val, idx, lab_all[: c["nq"]], lab_all[c["nq"]:], 0.4, (1, 5, 10, 20, 50))
    for k, v in golden["c3s_nih"].items():
        assert got[k] == pytest.approx(v, rel=1e-12), k


def test_multilabel_self_retrieval(golden):
    c = golden["cases"]["ml_self"]
    mlab = synth.multihot(c["n"], c["seed"])
    emb = oracle.normalize(synth.labelset_clustered(mlab, c["d"], c["seed"] + 100, c["noise"]))
    n = c["n"]
    S = oracle.scores(emb, emb, "cosine", "exclude", 0)
    fv, fi = oracle.topk(S, n - 1)
    got = rm.multilabel_retrieval_metrics(fv, fi, mlab, (1, 5, 10), 0.4)
    for k, v in golden["ml_self_train"].items():
        assert got[k] == pytest.approx(v, rel=1e-9), k
    for t, v in golden["ml_self_map_multilabel"].items():
        assert rm.compute_map_multilabel(fi, mlab, float(t)) == pytest.approx(v, rel=1e-12)
    prk = rm.multilabel_precision_recall_at_k(fi, mlab, mlab, (1, 5, 10, 15, 20))
    for k, (p, r) in golden["ml_self_prk_printed"].items():
        assert round(prk[int(k)][0], 2) == p and round(prk[int(k)][1], 2) == r
    S1 = oracle.scores(emb, emb, "cosine", "minus1", 0)
    v1, i1 = oracle.topk(S1, n, drop_masked=False)
    assert rm.evaluate_map(v1, i1, mlab, 0.4) == pytest.approx(golden["ml_self_evaluate_map"], rel=1e-9)


def test_single_label_self_retrieval(golden):
    c = golden["cases"]["sl_self"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    e = oracle.normalize(x)
    S = oracle.scores(e, e, "cosine", "exclude", 0)
    _, fi = oracle.topk(S, c["n"] - 1)
    got = rm.single_label_retrieval_metrics(fi, lab, (1, 5, 10))
    for k, v in golden["sl_self_train"].items():
        assert got[k] == pytest.approx(v, rel=1e-5), k  # reference accumulates AP in torch float32
    got = rm.fusion_evaluate_retrieval_metrics(fi, [f"class{v}" for v in lab], (1, 5, 10))
    for k, v in golden["sl_self_fusion"].items():
        assert got[k] == pytest.approx(v, rel=1e-12), k


def test_sklearn_ap_restatement_matches_sklearn():
    from sklearn.metrics import average_precision_score

    rs = np.random.RandomState(5)
    for n in (1, 2, 7, 8, 9, 50, 129, 300):
        for _ in range(5):
            s = np.sort(np.round(rs.rand(n), 1 if n > 20 else 3).astype(np.float32))[::-1]  # many ties
            rel = (rs.rand(n) < 0.3).astype(np.float64)
            if rel.sum() == 0:
                rel[rs.randint(n)] = 1.0
            assert rm.average_precision_ranked(s, rel) == average_precision_score(rel, s)
