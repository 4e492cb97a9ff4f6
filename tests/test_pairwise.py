"""SURVEY 8(f)-4 tail: batch-hard / batch-all triplet mining (loss.py:60-112), the Jaccard similarity matrix
(loss.py:237-242) and the nearest-centroid anomaly score (anomaly/test_anomaly.py:31-48).

CPU: the restatements of oracle/reference_metrics.py against goldens of the REAL reference functions
(oracle/make_golden_pairwise.py).  GPU: the library against the goldens and the restatements."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import make_golden_pairwise as mgp
from oracle import reference_metrics as RM
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gp():
    with open(os.path.join(GOLDEN, "golden_pairwise.json")) as fh:
        g = json.load(fh)
    g["arrays"] = dict(np.load(os.path.join(GOLDEN, "golden_pairwise_arrays.npz")))
    return g


def test_oracle_restatements_match_the_reference(gp):
    for name in mgp.CASES:
        x, lab = mgp.triplet_inputs(name)
        dist = oracle.scores(x, x, "l2")                      # the engine's fp32 GEMM-form distances
        for m in mgp.MARGINS:
            want = gp[f"{name}_m{m}"]
            assert RM.triplet_batch_hard(dist, lab, m) == pytest.approx(want["hard"], rel=2e-6)
            a, f = RM.triplet_batch_all(dist, lab, m)
            assert a == pytest.approx(want["all"], rel=1e-5) and f == pytest.approx(want["fraction"], rel=1e-4)
    ml = synth.multihot(gp["jaccard_case"]["n"], seed=gp["jaccard_case"]["seed"])
    assert np.array_equal(RM.jaccard_sim_matrix(ml), gp["arrays"]["jaccard"])
    xtr, ltr, xte, _ = mgp.anomaly_inputs()
    got = RM.nearest_centroid_scores(xtr, ltr, xte)
    assert np.allclose(got, gp["arrays"]["anomaly_scores"], rtol=1e-6, atol=0)   # numpy's fp32 mean order may differ


@pytest.mark.gpu
def test_library_matches_reference_and_oracle(gp):
    import torch

    import b200knn
    from b200knn import pairwise as PW

    for name in mgp.CASES:
        x, lab = mgp.triplet_inputs(name)
        xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
        dist = b200knn.scores_dense(xd, xd, "l2").cpu().numpy()
        assert np.array_equal(dist, oracle.scores(x, x, "l2"))
        for m in mgp.MARGINS:
            want = gp[f"{name}_m{m}"]
            h, minus1 = PW.batch_hard_triplet_loss(ld, xd, m)
            assert minus1 == -1 and h == pytest.approx(want["hard"], rel=2e-6)
            assert h == RM.triplet_batch_hard(dist, lab, m)                      # bit-exact against the restatement
            a, f = PW.batch_all_triplet_loss(ld, xd, m)
            assert a == pytest.approx(want["all"], rel=1e-5) and f == pytest.approx(want["fraction"], rel=1e-4)
            oa, of = RM.triplet_batch_all(dist, lab, m)
            assert a == pytest.approx(oa, rel=1e-12) and f == of
    with pytest.raises(ValueError):
        PW.batch_hard_triplet_loss(ld, xd, 1.0, p=1.0)
    ml = synth.multihot(gp["jaccard_case"]["n"], seed=gp["jaccard_case"]["seed"])
    assert np.array_equal(PW.compute_jaccard_sim(torch.from_numpy(ml).cuda()).cpu().numpy(), gp["arrays"]["jaccard"])
    xtr, ltr, xte, _ = mgp.anomaly_inputs()
    cm = PW.class_means(torch.from_numpy(xtr).cuda(), torch.from_numpy(ltr).cuda(), (0, 1)).cpu().numpy()
    assert np.allclose(cm, gp["arrays"]["class_means"], rtol=2e-6, atol=1e-7)
    got = PW.nearest_centroid_scores(torch.from_numpy(xtr).cuda(), torch.from_numpy(ltr).cuda(), torch.from_numpy(xte).cuda())
    assert np.allclose(got, gp["arrays"]["anomaly_scores"], rtol=1e-6, atol=0)
    assert np.array_equal(got, RM.nearest_centroid_scores(xtr, ltr, xte))         # bit-exact against the restatement
