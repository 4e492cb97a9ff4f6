"""Golden vectors for the remaining metric rows of SURVEY 8(a) from the REAL reference:

  D12  evaluate_medsiglip.evaluate_retrieval                      (evaluate_medsiglip.py:142-163)
  D10' train_ath.compute_retrieval_metrics (L2 and Hamming)        (train_ath.py:171-218)
  D14  chestmir_eval.retrieval_accuracy_from_ranks /
       compute_classification_metrics_from_ranks                   (ChestMIR/chestmir_eval.py:191-272)
  D9   evaluate_nih_zilliz.jaccard_score / precision_at_k / recall_at_k  (evaluate_nih_zilliz.py:12-31)

    python -m oracle.make_golden_metrics2          (build container only: needs /root/reference)
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np
import torch

from . import ref_shim, synth
from . import normalize as oracle_normalize

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {
    "d12": dict(n=500, d=96, classes=3, seed=61, noise=5.0, topk=[1, 5, 10, 20]),
    "d10": dict(nq=180, ng=1200, d=48, bits=36, classes=4, seed=62, noise=3.0, flip=0.25, topk=[1, 5, 10]),
    "d14": dict(n=300, d=64, classes=4, seed=63, noise=3.5, topk=[1, 5, 10], k_values=[1, 5, 10, 15, 20]),
}


def d10_codes(c):
    rs = np.random.RandomState(c["seed"])
    proto = rs.randint(0, 2, size=(c["classes"], c["bits"]))
    lab = rs.randint(0, c["classes"], size=c["nq"] + c["ng"])
    flips = rs.random_sample((len(lab), c["bits"])) < c["flip"]
    return (proto[lab] ^ flips).astype(np.float32), lab.astype(np.int64)


def main():
    warnings.filterwarnings("ignore")
    G = {"cases": CASES}
    # ---- D12 -------------------------------------------------------------------------------------------------
    c = CASES["d12"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    e = oracle_normalize(x)
    med = ref_shim.module("evaluate_medsiglip")
    out = med.evaluate_retrieval(torch.from_numpy(e.copy()), torch.from_numpy(lab), c["topk"])
    G["d12"] = {k: float(v) for k, v in out.items()}
    # ---- D10' ------------------------------------------------------------------------------------------------
    c = CASES["d10"]
    ath = ref_shim.module("train_ath")
    x, lab = synth.clustered(c["nq"] + c["ng"], c["d"], c["classes"], c["seed"], c["noise"])
    x = oracle_normalize(x)
    q, g, ql, gl = x[: c["nq"]], x[c["nq"]:], lab[: c["nq"]], lab[c["nq"]:]
    out = ath.compute_retrieval_metrics(torch.from_numpy(q), torch.from_numpy(ql), torch.from_numpy(g),
                                        torch.from_numpy(gl), c["topk"], False)
    G["d10_l2"] = {str(k): {m: float(v) for m, v in dd.items()} for k, dd in out.items()}
    codes, clab = d10_codes(c)
    # Hamming distances tie massively and the reference ranks them with an UNSTABLE torch.argsort (train_ath.py:175):
    # its tie order is unspecified.  SURVEY 8.1-Q1: the oracle form is the same function with the sort made stable
    # (ties -> ascending gallery row), which is what the engine implements.
    orig_argsort = torch.argsort
    torch.argsort = lambda *a, **k: orig_argsort(*a, **{**k, "stable": True})
    try:
        out = ath.compute_retrieval_metrics(torch.from_numpy(codes[: c["nq"]]), torch.from_numpy(clab[: c["nq"]]),
                                            torch.from_numpy(codes[c["nq"]:]), torch.from_numpy(clab[c["nq"]:]),
                                            c["topk"], True)
    finally:
        torch.argsort = orig_argsort
    G["d10_hamming"] = {str(k): {m: float(v) for m, v in dd.items()} for k, dd in out.items()}
    # ---- D14 -------------------------------------------------------------------------------------------------
    c = CASES["d14"]
    cm = ref_shim.module("ChestMIR.chestmir_eval")
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    e = oracle_normalize(x)
    sim = (e @ e.T).astype(np.float32)
    np.fill_diagonal(sim, -np.inf)
    # the reference ranks columns with an unstable argsort (chestmir_eval.py:431); the deterministic form of it
    ranks = np.argsort(-sim, axis=0, kind="stable")
    names = np.array(["normal", "pneumonia", "covid", "tb"], dtype=object)[lab]     # string labels, as in ChestMIR
    G["d14_acc"] = [float(v) for v in cm.retrieval_accuracy_from_ranks(ranks, names, c["topk"])]
    out = cm.compute_classification_metrics_from_ranks(names, ranks, c["k_values"])
    G["d14_cls"] = {str(k): {m: float(v) for m, v in dd.items()} for k, dd in out.items()}
    # ---- D9 helpers ------------------------------------------------------------------------------------------
    nih = ref_shim.module("evaluate_nih_zilliz")
    rs = np.random.RandomState(64)
    a = synth.multihot(40, 65)
    b = np.clip(a * (rs.random_sample(a.shape) < 0.7) + synth.multihot(40, 66) * (rs.random_sample(a.shape) < 0.5), 0, 1)
    b = b.astype(np.float32)
    G["d9_b_seedcheck"] = float(b.sum())
    G["d9_jaccard"] = [float(nih.jaccard_score(list(map(float, a[i])), list(map(float, b[i])))) for i in range(40)]
    rel = [float(v) for v in (rs.random_sample(30) < 0.3)]
    G["d9_rel"] = rel
    G["d9_p_at"] = {str(k): nih.precision_at_k(rel, k) for k in (1, 5, 10, 50)}
    G["d9_r_at"] = {str(k): nih.recall_at_k(rel, int(sum(rel)), k) for k in (1, 5, 10, 50)}
    with open(os.path.join(OUT, "golden_metrics2.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    print({k: G[k] for k in ("d12", "d14_acc")})


if __name__ == "__main__":
    main()
