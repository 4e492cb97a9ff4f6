"""Golden outputs for the full-ranking metrics on inputs the other goldens do not cover, by the REAL reference functions.

    python -m oracle.make_golden_fullrank          (build container only: needs /root/reference)

* ``fusion_eval.metrics.evaluate_retrieval_metrics`` (fusion_eval/metrics.py:26-94) with image paths that are NOT unique:
  the reference drops every gallery row sharing the query's path from the ranking but still counts it in
  ``relevant_count`` (metrics.py:67-69).
Output: tests/golden/golden_fullrank.json.
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np

from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DUP = dict(n=180, d=48, classes=4, seed=61, noise=2.5, n_dup=14)


def dup_inputs():
    x, lab = synth.clustered(DUP["n"], DUP["d"], DUP["classes"], DUP["seed"], DUP["noise"])
    paths = [f"scan/img_{i:04d}.png" for i in range(DUP["n"])]
    rs = np.random.RandomState(DUP["seed"] + 1)
    src = rs.choice(DUP["n"], size=DUP["n_dup"], replace=False)
    for s in src:            # a second row stored under the same image path (another crop of the same scan)
        t = int(rs.randint(0, DUP["n"]))
        paths[t] = paths[int(s)]
    return x, [f"class{v}" for v in lab], paths


def main():
    warnings.filterwarnings("ignore")
    fm = ref_shim.module("fusion_eval.metrics")
    x, labels, paths = dup_inputs()
    G = {"dup": DUP, "n_unique_paths": len(set(paths)),
         "dup_fusion": {k: float(v) for k, v in fm.evaluate_retrieval_metrics(x, labels, paths, (1, 3, 5, 10, 20)).items()}}
    with open(os.path.join(OUT, "golden_fullrank.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    print(G)


if __name__ == "__main__":
    main()
