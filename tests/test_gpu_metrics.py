"""GPU parity tests for the retrieval metrics: b200knn.metrics (CUDA kernels behind the C ABI) against the
restated reference metrics (oracle/reference_metrics.py) and the golden outputs of the real reference."""
import numpy as np
import pytest
import torch

import oracle
from oracle import reference_metrics as rm
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knn():
    import b200knn

    assert torch.cuda.is_available()
    b200knn.load_library()
    return b200knn


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def c1(knn, golden):
    c = golden["cases"]["c1"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    return knn.normalize(dev(x)), lab


def test_c1_pipeline_cosine(knn, c1, golden, golden_arrays):
    """test.py evaluate()-style pipeline on config 1 (cosine): R@K, compute_map, majority-vote classification."""
    e, lab = c1
    M = knn.metrics
    g = golden["c1_cosine"]
    S = knn.scores_dense(e, e, "cosine", self_mode="exclude")
    labels = dev(lab)
    acc = M.retrieval_accuracy(S, labels, topk=(1, 5, 10))
    assert [float(a) for a in acc] == g["acc"] and all(a.dtype == torch.float32 for a in acc)
    ranks_cols = knn.rank_rows(S.t().contiguous()).t().contiguous()  # [db, nq] like argsort(dim=0)
    mAP, aps, pr, prs = M.compute_map(ranks_cols, lab, [1, 5, 10])
    # engine == oracle bit for bit (same fp32 summation order -> same ranking -> same float64 arithmetic)
    o_ranks = oracle.rank_rows(oracle.scores(host(e), host(e), "cosine", "exclude", 0))
    o_mAP, o_aps, o_pr, o_prs = rm.compute_map(o_ranks, lab, lab, [1, 5, 10])
    assert mAP == o_mAP and np.array_equal(aps, o_aps) and np.array_equal(pr, o_pr) and np.array_equal(prs, o_prs)
    # engine vs the reference: top-k metrics identical, full-ranking mAP within the near-tie noise
    assert abs(mAP - g["mAP"]) < 1e-6 and list(pr) == g["pr"]
    cls = M.compute_classification_metrics(labels, S, [1, 5, 10, 15, 20])
    for k, d in g["classification"].items():
        for m, v in d.items():
            assert cls[int(k)][m] == pytest.approx(v, rel=1e-12), (k, m)
    # the metric kernels are bit-exact when fed the reference's OWN ranking
    ref_ranks = golden_arrays["c1_cosine_ranks_rowmajor"].astype(np.int64)
    mAP_r, aps_r, pr_r, prs_r = M.compute_map(dev(ref_ranks.T.copy()), lab, [1, 5, 10])
    assert mAP_r == g["mAP"] and list(pr_r) == g["pr"]
    assert np.array_equal(aps_r, golden_arrays["c1_cosine_aps"]) and np.array_equal(prs_r, golden_arrays["c1_cosine_prs"])


def test_c1_pipeline_topk_only_never_materialises(knn, c1, golden):
    e, lab = c1
    M = knn.metrics
    _, idx = knn.search(e, e, 20, "cosine", exclude_self=True)
    acc = M.recall_at_k_from_topk(idx, lab, lab, (1, 5, 10))
    assert [float(a) for a in acc] == golden["c1_cosine"]["acc"]
    cls = M.classification_metrics_from_topk(idx, lab, lab, (1, 5, 10, 15, 20))
    for k, d in golden["c1_cosine"]["classification"].items():
        for m, v in d.items():
            assert cls[int(k)][m] == pytest.approx(v, rel=1e-12)


def test_c1_cdist_pipeline(knn, c1, golden):
    e, lab = c1
    M = knn.metrics
    g = golden["c1_cdist"]
    D = knn.scores_dense(e, e, "l2", self_mode="exclude")
    acc = M.retrieval_accuracy(-D, dev(lab), topk=(1, 5, 10))
    assert [float(a) for a in acc] == g["acc"]
    ranks_cols = knn.rank_rows(D.t().contiguous(), largest_first=False).t().contiguous()
    mAP, _, pr, _ = M.compute_map(ranks_cols, lab, [1, 5, 10])
    assert abs(mAP - g["mAP"]) < 1e-6 and list(pr) == g["pr"]


def test_evaluate_embeddings_is_the_reference_evaluate_body(knn, c1, golden, tmp_path):
    """One call = test.py:1077-1126 after the forward pass, for both distance conventions, + the npz bundle."""
    e, lab = c1
    for metric, key in (("cosine", "c1_cosine"), ("l2", "c1_cdist")):
        g = golden[key]
        path = str(tmp_path / f"eval_{metric}.npz")
        out = knn.metrics.evaluate_embeddings(e, dev(lab), metric, save_path=path)
        assert out["acc"].dtype == np.float32 and [float(a) for a in out["acc"]] == g["acc"]
        assert abs(out["mAP"] - g["mAP"]) < 1e-6 and list(out["pr"]) == g["pr"]
        for k, d in g["classification"].items():
            got = out["classification"][int(k)]
            want = d if isinstance(d, dict) else dict(zip(got.keys(), d))   # cdist golden: values in the npz order
            for m, v in want.items():
                assert got[m] == pytest.approx(v, rel=1e-12), (k, m)
        z = np.load(path, allow_pickle=True)
        assert set(z.files) >= {"embeds", "labels", "dists", "kappas", "acc", "mAP", "pr", "classification_k_values"}
        assert np.isposinf(np.diag(z["dists"])).all() and float(z["mAP"]) == out["mAP"]   # stored negated, as test.py:1123


def test_c2_test_ath_compute_metrics(knn, golden):
    c = golden["cases"]["c2"]
    x, lab = synth.clustered(c["nq"] + c["ng"], c["d"], c["classes"], c["seed"], c["noise"], c["priors"])
    e = knn.normalize(dev(x))
    out = knn.metrics.compute_metrics(e[: c["nq"]], dev(lab[: c["nq"]]), e[c["nq"]:], dev(lab[c["nq"]:]), None, (1, 5, 10))
    for k, d in golden["c2_l2"]["retrieval"].items():
        for m, v in d.items():
            assert out["retrieval"][int(k)][m] == v, (k, m)  # bit-exact float64


def test_c3_small_evaluate_results_and_hit_rate(knn, golden, golden_arrays):
    c = golden["cases"]["c3s"]
    lab_all = synth.multihot(c["nq"] + c["ng"], c["seed"])
    emb = knn.normalize(dev(synth.labelset_clustered(lab_all, c["d"], c["seed"] + 100, c["noise"])))
    q, g = emb[: c["nq"]], emb[c["nq"]:]
    ql, gl = dev(lab_all[: c["nq"]]), dev(lab_all[c["nq"]:])
    vals, idx = knn.search(q, g, c["k"], "cosine")
    got = knn.metrics.evaluate_results_from_topk(vals, idx, ql, gl, 0.4, (1, 5, 10, 20, 50))
    want = rm.evaluate_results(host(vals), host(idx), lab_all[: c["nq"]], lab_all[c["nq"]:], 0.4, (1, 5, 10, 20, 50))
    assert got == want  # bit-exact vs the restated reference on the engine's own hit lists
    for k, v in golden["c3s_nih"].items():
        assert got[k] == pytest.approx(v, rel=1e-9), k
    hr = knn.metrics.multilabel_hit_rate_from_topk(idx, ql, gl, (1, 5, 10, 15, 20))
    want = rm.multilabel_precision_recall_at_k(host(idx), lab_all[: c["nq"]], lab_all[c["nq"]:], (1, 5, 10, 15, 20))
    for k in want:
        assert hr[k] == (float(want[k][0]), float(want[k][1]))


def test_multilabel_self_retrieval_metrics(knn, golden):
    c = golden["cases"]["ml_self"]
    mlab = synth.multihot(c["n"], c["seed"])
    emb = dev(synth.labelset_clustered(mlab, c["d"], c["seed"] + 100, c["noise"]))
    tl = dev(mlab)
    M = knn.metrics
    got = M._compute_multilabel_retrieval_metrics(emb, tl)
    for k, v in golden["ml_self_train"].items():
        assert got[k] == pytest.approx(v, rel=1e-9), k
    for t, v in golden["ml_self_map_multilabel"].items():
        assert M.compute_map_multilabel_from_embeddings(emb, tl, float(t)) == pytest.approx(v, rel=1e-12)
        # the reference's own signature: dense dists with the diagonal at -inf (test.py:1007 -> :941)
        dense = knn.scores_dense(emb, emb, "cosine", normalize=True, self_mode="exclude")
        assert M.compute_map_multilabel(dense, tl, float(t)) == pytest.approx(v, rel=1e-12)
    assert M.evaluate_map_embeddings(emb, tl, 0.4) == pytest.approx(golden["ml_self_evaluate_map"], rel=1e-9)
    whole = M.evaluate_multilabel_embeddings(emb, tl)                       # test.py:987-1062 in one call
    for t, v in golden["ml_self_map_multilabel"].items():
        assert whole["mAP"][float(t)] == pytest.approx(v, rel=1e-12)
    for k, (p, r) in golden["ml_self_prk_printed"].items():
        assert round(whole["precision_recall_at_k"][int(k)][0], 2) == p
        assert round(whole["precision_recall_at_k"][int(k)][1], 2) == r
    _, idx = knn.search(emb, emb, 20, "cosine", normalize=True, exclude_self=True)
    hr = M.multilabel_hit_rate_from_topk(idx, tl, tl, (1, 5, 10, 15, 20))
    for k, (p, r) in golden["ml_self_prk_printed"].items():
        assert round(hr[int(k)][0], 2) == p and round(hr[int(k)][1], 2) == r


def test_single_label_self_retrieval_metrics(knn, golden):
    c = golden["cases"]["sl_self"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    M = knn.metrics
    got = M._compute_single_label_retrieval_metrics(dev(x), dev(lab))
    for k, v in golden["sl_self_train"].items():
        assert got[k] == pytest.approx(v, rel=1e-5), k  # reference accumulates in torch float32
    got = M.evaluate_retrieval_metrics(x, [f"class{v}" for v in lab], [f"img{i}.png" for i in range(len(lab))], (1, 5, 10))
    for k, v in golden["sl_self_fusion"].items():
        assert got[k] == pytest.approx(v, rel=1e-12), k


def test_ap_sklearn_kernel_matches_sklearn(knn):
    from sklearn.metrics import average_precision_score

    rs = np.random.RandomState(5)
    for k in (1, 2, 7, 8, 9, 50, 129, 300):
        s = np.sort(np.round(rs.rand(40, k), 1 if k > 20 else 3).astype(np.float32), axis=1)[:, ::-1].copy()
        rel = (rs.rand(40, k) < 0.3).astype(np.uint8)
        rel[3] = 0
        ap = host(knn.metrics.ap_sklearn(dev(s), dev(rel)))
        for r in range(40):
            if rel[r].sum() == 0:
                assert np.isnan(ap[r])
            else:
                assert ap[r] == average_precision_score(rel[r], s[r]), (k, r)


def test_majority_vote_tie_rules(knn):
    lab = np.array([[2, 1, 1, 2, 0], [3, 3, 1, 1, 0], [5, 4, 3, 2, 1]], dtype=np.int64)
    M = knn.metrics
    assert host(M.majority_vote_labels(dev(lab), 4, "first")).tolist() == [2, 3, 5]
    assert host(M.majority_vote_labels(dev(lab), 4, "smallest")).tolist() == [1, 1, 2]
    assert host(M.majority_vote_labels(dev(lab), 1, "first")).tolist() == [2, 3, 5]
    for r in range(3):
        assert rm.majority_vote(lab[r, :4]) == [2, 3, 5][r] and rm.majority_vote(lab[r, :4], "smallest") == [1, 1, 2][r]


def test_jaccard_threshold_edge(knn):
    """inter=2, union=5 -> J = 0.4 is NOT > 0.4 in either arithmetic."""
    q = np.zeros((1, 14), dtype=np.float32); q[0, :4] = 1           # {0,1,2,3}
    g = np.zeros((3, 14), dtype=np.float32)
    g[0, [0, 1, 4]] = 1        # inter 2, union 5 -> 0.4
    g[1, [0, 1, 2]] = 1        # inter 3, union 4 -> 0.75
    g[2, [5]] = 1              # disjoint
    M = knn.metrics
    idx = dev(np.array([[0, 1, 2]], dtype=np.int64))
    for arith in ("fp32", "fp64"):
        rj, ra = M.relevance_multilabel(idx, M.pack_multihot(dev(q)), M.pack_multihot(dev(g)), 0.4, arith)
        assert host(rj).tolist() == [[0, 1, 0]] and host(ra).tolist() == [[1, 1, 0]]
