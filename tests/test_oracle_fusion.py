"""CPU: the late-fusion / re-ranking oracle (oracle/reference_fusion.py) against the golden vectors the REAL
reference produced (oracle/make_golden_fusion.py -> tests/golden/golden_fusion*)."""
import json
import os

import numpy as np
import pytest

from util import tie_aware_mismatches

from oracle import make_golden_fusion as mg
from oracle import reference_fusion as rf
from oracle import reference_metrics as rm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gf():
    with open(os.path.join(GOLDEN, "golden_fusion.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def ga():
    return dict(np.load(os.path.join(GOLDEN, "golden_fusion_arrays.npz")))


def _metrics_from_similarity(sim, lab, ks=(1, 5, 10)):
    _, idx = rf.stable_self_ranking(sim)
    return rm.fusion_evaluate_retrieval_metrics(idx, [f"class{v}" for v in lab], ks)


@pytest.mark.parametrize("mode", ["zscore", "minmax"])
def test_normalised_similarity_and_fused_topk(ga, mode):
    conv, dino, _ = mg.fusion_inputs(mg.CASES["fusion"])
    cs = rf.normalize_similarity_matrix(rf.compute_similarity_matrix(conv), mode)
    # numpy's sgemm is not bit-reproducible across BLAS thread counts: values at a few ulp, ranking tie-aware
    assert np.allclose(cs[:4], ga[f"fusion_{mode}_conv_rows_0_4"], rtol=0, atol=5e-6)
    fused = rf.score_fusion_similarity(conv, dino, 0.5, mode)
    val, idx = rf.stable_self_ranking(fused)
    assert np.allclose(val[:, :10], ga[f"fusion_{mode}_a0.5_top10_val"], rtol=0, atol=5e-6)
    _, bad = tie_aware_mismatches(val[:, :10], idx[:, :10], ga[f"fusion_{mode}_a0.5_top10_val"],
                                  ga[f"fusion_{mode}_a0.5_top10_idx"], tol=1e-5)
    assert bad == 0


def test_confidence_fusion(ga, gf):
    conv, dino, _ = mg.fusion_inputs(mg.CASES["fusion"])
    conf = rf.confidence_based_fusion(rf.compute_similarity_matrix(conv), rf.compute_similarity_matrix(dino))
    val, idx = rf.stable_self_ranking(conf["similarity"])
    assert np.allclose(val[:, :10], ga["fusion_conf_top10_val"], rtol=0, atol=5e-6)
    _, bad = tie_aware_mismatches(val[:, :10], idx[:, :10], ga["fusion_conf_top10_val"], ga["fusion_conf_top10_idx"],
                                  tol=1e-5)
    assert bad == 0
    assert conf["conv_selected_queries"] == gf["fusion_conf"]["conv_selected_queries"]
    assert conf["dino_selected_queries"] == gf["fusion_conf"]["dino_selected_queries"]


@pytest.mark.parametrize("case", ["fusion", "fusion_same_dim"])
@pytest.mark.parametrize("mode", ["none", "zscore", "minmax"])
def test_experiment_loop_metrics(gf, case, mode):
    conv, dino, lab = mg.fusion_inputs(mg.CASES[case])
    got = {}
    for name, emb in (("convnext_baseline", rf.l2_normalize(conv)), ("dino_baseline", rf.l2_normalize(dino)),
                      ("concat_fusion", rf.concat_fusion(conv, dino))):
        got[name] = _metrics_from_similarity(rf.compute_similarity_matrix(emb), lab)
    for a in (0.2, 0.5, 0.8):
        got[f"score_fusion_alpha_{a:.1f}"] = _metrics_from_similarity(rf.score_fusion_similarity(conv, dino, a, mode), lab)
        ws = rf.weighted_sum_fusion(conv, dino, a)
        got[f"weighted_sum_alpha_{a:.1f}"] = None if ws is None else _metrics_from_similarity(
            rf.compute_similarity_matrix(ws), lab)
    conf = rf.confidence_based_fusion(rf.normalize_similarity_matrix(rf.compute_similarity_matrix(conv), mode),
                                      rf.normalize_similarity_matrix(rf.compute_similarity_matrix(dino), mode))
    got["confidence_fusion_top12_margin"] = _metrics_from_similarity(conf["similarity"], lab)
    for want in gf[f"{case}_{mode}"]:
        g = got[want["experiment_name"]]
        if want["skipped"]:
            assert g is None and "dimension_mismatch" in want["skipped_reason"]
            continue
        for key, v in want["metrics"].items():
            if key.endswith("_selected_queries"):
                continue
            # full-ranking metrics see the last-ulp differences of the similarity matrices as rare rank swaps
            assert abs(g[key] - v) <= 2e-3, (want["experiment_name"], key, g[key], v)


def test_rerank_restatement(ga):
    c = mg.CASES["rerank"]
    x, lab, text = mg.rerank_inputs(c)
    e = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    img_sim = e @ e.T
    table = ga["rerank_table"]
    d = rf.rerank_rows(img_sim, table, lab, c["rerank_k"], c["text_weight"], 1.0 - c["text_weight"])
    # numpy's and torch's sgemm / norm may differ in the last ulp: compare at 1e-6, the ranking tie-aware
    assert np.allclose(d[:4][np.isfinite(d[:4])], ga["rerank_dists_rows_0_4"][np.isfinite(ga["rerank_dists_rows_0_4"])],
                       rtol=0, atol=2e-6)
    order = np.argsort(-d, axis=1, kind="stable")[:, :10]
    same = np.mean(order == ga["rerank_top10_idx"])
    assert same >= 0.995, same
