"""CPU: the oracle restatements of the D12 / D10' / D14 metric rows and of the lesion re-ranking against golden outputs
of the REAL reference functions (oracle/make_golden_metrics2.py, oracle/make_golden_lesion.py)."""
import json
import os

import numpy as np

import oracle
from oracle import reference_metrics as rm
from oracle import synth
from oracle.make_golden_lesion import CASE as LESION_CASE
from oracle.make_golden_lesion import inputs as lesion_inputs
from oracle.make_golden_metrics2 import d10_codes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _G(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def _stable_desc(scores):
    return np.argsort(-scores, axis=1, kind="stable")


def test_d12_evaluate_retrieval():
    G = _G("golden_metrics2.json")
    c = G["cases"]["d12"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    e = oracle.normalize(x)
    sim = oracle.scores(e, e, "ip", "minus1", 0)
    got = rm.medsiglip_evaluate_retrieval(_stable_desc(sim)[:, : max(c["topk"])], lab, c["topk"])
    assert set(got) == set(G["d12"])
    for k, v in G["d12"].items():
        assert abs(got[k] - v) <= 1e-12 * max(1.0, abs(v)), k


def test_d10_train_ath_metrics_l2_and_hamming():
    G = _G("golden_metrics2.json")
    c = G["cases"]["d10"]
    x, lab = synth.clustered(c["nq"] + c["ng"], c["d"], c["classes"], c["seed"], c["noise"])
    x = oracle.normalize(x)
    nq = c["nq"]
    dist = oracle.scores(x[:nq], x[nq:], "l2", "keep", 0)
    got = rm.ath_train_retrieval_metrics(np.argsort(dist, axis=1, kind="stable"), lab[:nq], lab[nq:], c["topk"])
    for k in c["topk"]:
        for m, v in G["d10_l2"][str(k)].items():
            assert abs(got[k][m] - v) <= 1e-12, (k, m)
    codes, clab = d10_codes(c)
    ham = (codes[:nq, None, :] != codes[None, nq:, :]).sum(2)
    got = rm.ath_train_retrieval_metrics(np.argsort(ham, axis=1, kind="stable"), clab[:nq], clab[nq:], c["topk"])
    for k in c["topk"]:
        for m, v in G["d10_hamming"][str(k)].items():
            assert abs(got[k][m] - v) <= 1e-12, (k, m)


def test_d14_metrics_from_ranks():
    G = _G("golden_metrics2.json")
    c = G["cases"]["d14"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    e = oracle.normalize(x)
    sim = (e @ e.T).astype(np.float32)          # the golden run's own similarity matrix (numpy / BLAS order)
    np.fill_diagonal(sim, -np.inf)
    ranks = np.argsort(-sim, axis=0, kind="stable").T
    names = np.array(["normal", "pneumonia", "covid", "tb"], dtype=object)[lab]
    assert np.array_equal(rm.chestmir_accuracy_from_ranks(ranks, names, c["topk"]), np.array(G["d14_acc"]))
    got = rm.chestmir_classification_from_ranks(names, ranks, c["k_values"])
    for k in c["k_values"]:
        assert got[k] == G["d14_cls"][str(k)]


def test_lesion_rerank_restatement():
    G = _G("golden_lesion.json")
    A = dict(np.load(os.path.join(GOLDEN, "golden_lesion_arrays.npz")))
    c = G["case"]
    assert c == json.loads(json.dumps(LESION_CASE))
    g, _, maps = lesion_inputs(c)
    sim = (g @ g.T).astype(np.float32)
    np.fill_diagonal(sim, -np.inf)
    base_idx = np.argsort(-sim, axis=1, kind="stable")[:, : c["keep"]]
    base_val = np.take_along_axis(sim, base_idx, axis=1)
    assert np.array_equal(base_idx, A["base"])
    for name in c["lesions"]:
        choice = [(name, m[name][0]) if len(m.get(name, [])) else None for m in maps]
        got = rm.lesion_rerank(base_val, base_idx, maps, choice, c["rerank_topk"], c["global_weight"])
        assert np.array_equal(got, A[f"specific_{name.replace(' ', '_')}"]), name
    choice = []
    for m in maps:
        best, cnt = None, -1
        for name in c["lesions"]:
            if len(m.get(name, [])) > cnt and len(m.get(name, [])):
                cnt, best = len(m[name]), name
        choice.append((best, m[best][0]) if best is not None else None)
    got = rm.lesion_rerank(base_val, base_idx, maps, choice, c["rerank_topk"], c["global_weight"])
    assert np.array_equal(got, A["adaptive"])


def test_split_filter_error_bound_holds_in_exact_arithmetic():
    """The mathematical part of the exact engine's error bound (include/b200knn.h, knn_filter_error_bound): the three
    kept products of the bf16x3 split differ from q.g by at most 8.04 * 2^-18 * |q||g| -- checked in float64, where the
    split rows' inner product is exact.  Random rows stay far inside; rows whose every element sits just below a bf16
    rounding midpoint (|lo| ~ 2^-8 |x|) reach 4 * 2^-18 and break the 3.02 * 2^-18 round 1 claimed (u = 2^-9 instead of
    the 2^-8 of an 8-bit significand)."""
    rs = np.random.RandomState(3)
    for d, scale in ((64, 1.0), (1024, 1.0), (100, 1e3), (36, 1e-3)):
        q = (rs.standard_normal((50, d)) * scale).astype(np.float32)
        g = (rs.standard_normal((400, d)) * np.exp(rs.uniform(-3, 3, (400, 1)))).astype(np.float32)
        q3, g3 = oracle.split_bf16x3(q, "queries"), oracle.split_bf16x3(g, "gallery")
        assert q3.shape == (50, 3 * ((d + 7) // 8 * 8)) and np.array_equal(oracle.bf16_round(q3), q3)
        approx = q3.astype(np.float64) @ g3.astype(np.float64).T
        exact = q.astype(np.float64) @ g.astype(np.float64).T
        qn = np.linalg.norm(q.astype(np.float64), axis=1)[:, None]
        gn = np.linalg.norm(g.astype(np.float64), axis=1)[None, :]
        bound = 8.04 * 2.0 ** -18 * qn * gn
        assert np.all(np.abs(approx - exact) <= bound)
    # adversarial rows: x = 1 + 2^-8 - 2^-16 rounds DOWN to 1 and lo = x - 1 is exact: q_lo . g_lo ~ 2^-16 |q||g|
    x = np.full((4, 64), 1.0 + 2.0 ** -8 - 2.0 ** -16, dtype=np.float32)
    q3, g3 = oracle.split_bf16x3(x, "queries"), oracle.split_bf16x3(x, "gallery")
    err = np.abs(q3.astype(np.float64) @ g3.astype(np.float64).T - x.astype(np.float64) @ x.astype(np.float64).T)
    nn = np.linalg.norm(x.astype(np.float64), axis=1)[:, None] * np.linalg.norm(x.astype(np.float64), axis=1)[None, :]
    assert np.all(err > 3.02 * 2.0 ** -18 * nn) and np.all(err <= 8.04 * 2.0 ** -18 * nn)
