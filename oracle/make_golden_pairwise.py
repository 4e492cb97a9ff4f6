"""Golden outputs for the training-time pairwise row (SURVEY 8(f)-4 tail), produced by the REAL reference functions.

    python -m oracle.make_golden_pairwise          (build container only: needs /root/reference)

``loss.batch_hard_triplet_loss`` / ``loss.batch_all_triplet_loss`` (loss.py:60-112) and
``loss.JaccardSupConLoss.compute_jaccard_sim`` (loss.py:237-242) are imported and run unmodified; the nearest-centroid
score lives inside ``main()`` of anomaly/test_anomaly.py, so its three statements (lines 31-32 and 46-48: class means,
``scipy.spatial.distance.cdist(...).min(axis=1)``, division by the maximum) are executed here on the same arrays.
Output: tests/golden/golden_pairwise.json + golden_pairwise_arrays.npz.
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np

from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {"small": dict(n=24, d=32, classes=3, seed=71, noise=1.2), "batch": dict(n=160, d=128, classes=5, seed=72, noise=2.0)}
ANOM = dict(n_train=300, n_test=200, d=64, seed=73, noise=1.5)
MARGINS = (0.3, 1.0)


def triplet_inputs(name):
    c = CASES[name]
    return synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])


def anomaly_inputs():
    xtr, ltr = synth.clustered(ANOM["n_train"], ANOM["d"], 2, ANOM["seed"], ANOM["noise"])
    xte, lte = synth.clustered(ANOM["n_test"], ANOM["d"], 3, ANOM["seed"], ANOM["noise"])   # same centres + a third class
    return xtr, ltr, xte, lte


def main():
    import torch
    from scipy.spatial.distance import cdist

    warnings.filterwarnings("ignore")
    loss = ref_shim.module("loss")
    G, A = {"cases": CASES, "anomaly": ANOM, "margins": list(MARGINS)}, {}
    for name in CASES:
        x, lab = triplet_inputs(name)
        xt, lt = torch.from_numpy(x), torch.from_numpy(lab)
        for m in MARGINS:
            h, _ = loss.batch_hard_triplet_loss(lt, xt, m, 2.0)
            a, frac = loss.batch_all_triplet_loss(lt, xt, m, 2.0)
            G[f"{name}_m{m}"] = {"hard": float(h), "all": float(a), "fraction": float(frac)}
    ml = synth.multihot(90, seed=74)
    A["jaccard"] = loss.JaccardSupConLoss().compute_jaccard_sim(torch.from_numpy(ml)).numpy()
    G["jaccard_case"] = dict(n=90, seed=74)
    xtr, ltr, xte, lte = anomaly_inputs()
    class_0 = xtr[ltr == 0].mean(axis=0)                          # anomaly/test_anomaly.py:31
    class_1 = xtr[ltr == 1].mean(axis=0)                          # :32
    dists = cdist(xte, np.stack((class_0, class_1)))              # :46
    dists = dists.min(axis=1)                                     # :47
    dists /= dists.max()                                          # :48
    A["anomaly_scores"], A["class_means"] = dists, np.stack((class_0, class_1))
    with open(os.path.join(OUT, "golden_pairwise.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "golden_pairwise_arrays.npz"), **A)
    print({k: v for k, v in G.items() if "_m" in k})


if __name__ == "__main__":
    main()
