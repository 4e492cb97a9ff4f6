"""Golden vectors for the Hamming row from the REAL reference (test_ath.py).

    python -m oracle.make_golden_hamming          (build container only: needs /root/reference)
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np
import torch

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(nq=200, ng=1500, bits=36, classes=4, seed=51, flip=0.22)
CASE_WIDE = dict(nq=70, ng=900, bits=200, classes=3, seed=53, flip=0.3)


def codes(c):
    """class prototypes with independent bit flips -> 0/1 float codes like `(codes >= 0).float()` (test_ath.py:68)"""
    rs = np.random.RandomState(c["seed"])
    proto = rs.randint(0, 2, size=(c["classes"], c["bits"]))
    lab = rs.randint(0, c["classes"], size=c["nq"] + c["ng"])
    flips = rs.random_sample((len(lab), c["bits"])) < c["flip"]
    x = (proto[lab] ^ flips).astype(np.float32)
    return x[: c["nq"]], lab[: c["nq"]].astype(np.int64), x[c["nq"]:], lab[c["nq"]:].astype(np.int64)


def main():
    warnings.filterwarnings("ignore")
    ath = ref_shim.module("test_ath")
    G, A = {"case": CASE, "case_wide": CASE_WIDE}, {}
    for name, c in (("h36", CASE), ("h200", CASE_WIDE)):
        q, ql, g, gl = codes(c)
        tq, tg = torch.from_numpy(q), torch.from_numpy(g)
        d = ath.pairwise_distance(tq, tg, True).numpy()
        order = np.argsort(d, axis=1, kind="stable")[:, :10]
        A[f"{name}_top10_dist"] = np.take_along_axis(d, order, axis=1)
        A[f"{name}_top10_idx"] = order
        out = ath.compute_metrics(tq, torch.from_numpy(ql), tg, torch.from_numpy(gl), torch.zeros((c["nq"], c["classes"])),
                                  (1, 5, 10), True)
        G[name] = {str(k): {m: float(v) for m, v in dd.items()} for k, dd in out["retrieval"].items()}
    with open(os.path.join(OUT, "golden_hamming.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "golden_hamming_arrays.npz"), **A)
    print(G["h36"])


if __name__ == "__main__":
    main()
