"""GPU parity tests for the late-fusion / re-ranking row: b200knn.fusion through the C ABI against
the oracle restatement and against the golden vectors the real reference produced.

Bars: score statistics within 1e-6 relative of numpy's; fused / re-ranked top-k tie-aware equal to the reference with
values within 1e-5; experiment-loop metric dicts within 2e-3 percentage points (full-ranking mAP sees fp32 near-ties,
DESIGN.md section 5)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_fusion as mg
from oracle import reference_fusion as rf
from util import tie_aware_mismatches

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def knn():
    import b200knn

    assert torch.cuda.is_available()
    b200knn.load_library()
    return b200knn


@pytest.fixture(scope="module")
def gf():
    with open(os.path.join(GOLDEN, "golden_fusion.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def ga():
    return dict(np.load(os.path.join(GOLDEN, "golden_fusion_arrays.npz")))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def test_embedding_level_fusion_matches_fuse_py(knn):
    conv, dino, _ = mg.fusion_inputs(mg.CASES["fusion_same_dim"])
    F = knn.fusion
    assert np.allclose(host(F.l2_normalize(dev(conv))), rf.l2_normalize(conv), rtol=0, atol=2e-7)
    assert np.allclose(host(F.concat_fusion(dev(conv), dev(dino))), rf.concat_fusion(conv, dino), rtol=0, atol=3e-7)
    ws = F.weighted_sum_fusion(dev(conv), dev(dino), 0.4)
    assert np.allclose(host(ws.embeddings), rf.weighted_sum_fusion(conv, dino, 0.4), rtol=0, atol=3e-7)
    conv2, dino2, _ = mg.fusion_inputs(mg.CASES["fusion"])
    skipped = F.weighted_sum_fusion(dev(conv2), dev(dino2), 0.4)
    assert skipped.embeddings is None and skipped.skipped_reason.startswith("weighted_sum_skipped_dimension_mismatch")


@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_score_stats_match_numpy_row_statistics(knn, metric):
    rs = np.random.RandomState(5)
    q = rs.standard_normal((130, 72)).astype(np.float32)
    g = rs.standard_normal((3000, 72)).astype(np.float32)
    st = knn.fusion.score_stats(dev(q), dev(g), metric)
    import oracle
    s = oracle.scores(q, g, metric).astype(np.float64)
    assert np.allclose(host(st["mean"]), s.mean(1), rtol=1e-6, atol=1e-6)
    assert np.allclose(host(st["std"]), s.std(1), rtol=1e-6, atol=1e-6)
    assert np.array_equal(host(st["min"]).astype(np.float32), s.min(1).astype(np.float32))
    assert np.array_equal(host(st["max"]).astype(np.float32), s.max(1).astype(np.float32))
    # self excluded: statistics over the other rows only
    e = rf.l2_normalize(g[:500])
    st = knn.fusion.score_stats(dev(e), dev(e), "ip", self_mode="exclude")
    s = oracle.scores(e, e, "ip").astype(np.float64)
    np.fill_diagonal(s, np.nan)
    assert np.allclose(host(st["mean"]), np.nanmean(s, 1), rtol=1e-6, atol=1e-7)
    assert np.allclose(host(st["max"]), np.nanmax(s, 1), rtol=0, atol=0)


@pytest.mark.parametrize("mode", ["zscore", "minmax"])
def test_score_fusion_topk_against_the_reference(knn, ga, mode):
    conv, dino, _ = mg.fusion_inputs(mg.CASES["fusion"])
    v, i = knn.fusion.score_fusion_search(dev(conv), dev(dino), 0.5, 10, mode)
    gv, gi = ga[f"fusion_{mode}_a0.5_top10_val"], ga[f"fusion_{mode}_a0.5_top10_idx"]
    assert np.allclose(host(v), gv, rtol=1e-5, atol=2e-5)
    _, bad = tie_aware_mismatches(host(v), host(i), gv, gi, tol=5e-5)
    assert bad == 0


def test_confidence_fusion_against_the_reference(knn, ga, gf):
    conv, dino, _ = mg.fusion_inputs(mg.CASES["fusion"])
    (v, i), info = knn.fusion.confidence_fusion_search(dev(conv), dev(dino), 10, "none")
    assert np.allclose(host(v), ga["fusion_conf_top10_val"], rtol=1e-5, atol=2e-5)
    _, bad = tie_aware_mismatches(host(v), host(i), ga["fusion_conf_top10_val"], ga["fusion_conf_top10_idx"], tol=5e-5)
    assert bad == 0
    assert info["conv_selected_queries"] == gf["fusion_conf"]["conv_selected_queries"]
    assert info["dino_selected_queries"] == gf["fusion_conf"]["dino_selected_queries"]
    assert abs(info["alpha_mean"] - gf["fusion_conf"]["alpha_mean"]) < 1e-5


@pytest.mark.parametrize("case", ["fusion", "fusion_same_dim"])
@pytest.mark.parametrize("mode", ["none", "zscore", "minmax"])
def test_experiment_loop_reproduces_the_reference_metric_dicts(knn, gf, case, mode):
    conv, dino, lab = mg.fusion_inputs(mg.CASES[case])
    labels = [f"class{v}" for v in lab]
    paths = [f"img{i}.png" for i in range(len(lab))]
    res = knn.fusion.run_late_fusion_experiments(dev(conv), dev(dino), labels, paths, alpha_values=(0.2, 0.5, 0.8),
                                                 k_values=(1, 5, 10), include_score_fusion=True,
                                                 score_normalization=mode, include_confidence_fusion=True)
    if mode == "none":  # the reference's own call shape: one AlignedEmbeddings payload (fusion_eval/evaluate.py:30)
        aligned = knn.fusion.AlignedEmbeddings(image_paths=paths, labels=labels, conv_embeddings=conv,
                                               dino_embeddings=dino, coverage={})
        res2 = knn.fusion.run_late_fusion_experiments(aligned, alpha_values=(0.2, 0.5, 0.8), k_values=(1, 5, 10),
                                                      score_normalization=mode)
        assert [(r.experiment_name, r.metrics) for r in res2] == [(r.experiment_name, r.metrics) for r in res]
    want = gf[f"{case}_{mode}"]
    assert [r.experiment_name for r in res] == [w["experiment_name"] for w in want]
    for r, w in zip(res, want):
        assert r.skipped == w["skipped"] and r.num_samples == w["num_samples"]
        if w["skipped"]:
            assert r.skipped_reason == w["skipped_reason"]
            continue
        assert set(r.metrics) == set(w["metrics"])
        for key, v in w["metrics"].items():
            assert abs(r.metrics[key] - v) <= 2e-3, (r.experiment_name, key, r.metrics[key], v)


def test_rerank_search_against_test_py_608_623(knn, ga):
    c = mg.CASES["rerank"]
    x, lab, _ = mg.rerank_inputs(c)
    e = knn.normalize(dev(x), eps_mode="none")
    table = dev(ga["rerank_table"])
    v, i = knn.fusion.rerank_search(e, e, table, lab, 10, c["rerank_k"], c["text_weight"], metric="ip")
    assert np.allclose(host(v), ga["rerank_top10_val"], rtol=1e-5, atol=2e-6)
    _, bad = tie_aware_mismatches(host(v), host(i), ga["rerank_top10_val"], ga["rerank_top10_idx"], tol=1e-5)
    assert bad == 0
    # oracle restatement on the engine's own fp32 similarities: bit-exact re-scoring arithmetic
    import oracle
    en = host(e)
    d = rf.rerank_rows(oracle.scores(en, en, "ip"), ga["rerank_table"], lab, c["rerank_k"], c["text_weight"],
                       1.0 - c["text_weight"])
    order = np.argsort(-d, axis=1, kind="stable")[:, :10]
    assert np.array_equal(host(i), order)
    assert np.array_equal(host(v), np.take_along_axis(d, order, axis=1))


def test_sort_topk_orders_by_value_then_index(knn):
    rs = np.random.RandomState(9)
    vals = rs.randint(0, 6, size=(37, 300)).astype(np.float32)       # many ties
    idx = np.stack([rs.permutation(5000)[:300] for _ in range(37)]).astype(np.int64)
    for largest in (True, False):
        v, i = knn.fusion.sort_topk(dev(vals), dev(idx), largest)
        key = np.lexsort((idx, -vals if largest else vals), axis=1)
        assert np.array_equal(host(i), np.take_along_axis(idx, key, axis=1))
        assert np.array_equal(host(v), np.take_along_axis(vals, key, axis=1))


def test_reranker_protocol_default():
    import b200knn

    items = [{"id": 3}, {"id": 1}]
    assert b200knn.fusion.IdentityReranker().rerank({"q": 0}, iter(items)) == items
