"""Golden outputs of the REAL ``evaluate_results`` (evaluate_nih_zilliz.py:34-64) on hits-JSON items, for the
reference-signature entry ``metrics.evaluate_results`` (tests/test_nih_eval_items.py).

    python -m oracle.make_golden_nih_eval          (build container only: needs /root/reference)
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def items(nq=12, k=8, seed=23):
    """Hits with descending scores, a few exact score ties, queries without any relevant hit."""
    lab = synth.multihot(nq * (k + 1), seed=seed)
    rs = np.random.RandomState(seed)
    out = []
    for q in range(nq):
        scores = np.sort(rs.uniform(0.2, 0.99, k).astype(np.float32))[::-1].copy()
        if q % 3 == 0:
            scores[2] = scores[1]                                   # a tie inside the list
        hits = [{"id": int(q * k + j), "score": float(scores[j]), "label_vector": lab[nq + q * k + j].astype(int).tolist()}
                for j in range(k)]
        qv = lab[q].astype(int).tolist()
        if q == 5:
            qv = [0] * len(qv)                                      # a query without labels: nothing is relevant
        out.append({"query_image_path": f"q/{q}.npy", "query_label_vector": qv, "results": hits})
    return out


def main():
    ez = ref_shim.module("evaluate_nih_zilliz")
    it = items()
    gold = {"items": it, "ks": [1, 5, 10], "threshold": 0.4,
            "metrics": {k: float(v) for k, v in ez.evaluate_results(it, 0.4, [1, 5, 10]).items()},
            "metrics_thr_0": {k: float(v) for k, v in ez.evaluate_results(it, 0.0, [1, 3]).items()},
            "empty": ez.evaluate_results([], 0.4, [1, 5]),
            "no_hits": ez.evaluate_results([{"query_label_vector": [1, 0], "results": []},
                                            {"query_label_vector": [0, 1], "results": []}], 0.4, [1, 5])}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "golden_nih_eval.json"), "w", encoding="utf-8") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote golden_nih_eval.json", gold["metrics"])


if __name__ == "__main__":
    main()
