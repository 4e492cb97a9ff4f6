"""Golden vectors for the lesion re-ranking row from the REAL reference
(ChestMIR/chestmir_eval.py: rerank_with_specific_lesion, rerank_with_adaptive_lesion).

    python -m oracle.make_golden_lesion          (build container only: needs /root/reference)
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np

from . import normalize as oracle_normalize
from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(n=260, d=64, dl=24, classes=4, seed=71, noise=3.0, lesions=["consolidation", "nodule mass", "edema"],
            presence=[0.55, 0.35, 0.2], max_regions=3, rerank_topk=40, global_weight=0.6, keep=60)


def inputs(c):
    """Global vectors (class-clustered) and ragged region vectors per lesion; a lesion's vectors cluster by class so that
    the re-ranking actually moves candidates."""
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    g = oracle_normalize(x)
    rs = np.random.RandomState(c["seed"] + 1)
    centers = rs.standard_normal((len(c["lesions"]), c["classes"], c["dl"])).astype(np.float32)
    maps = []
    for i in range(c["n"]):
        m = {}
        for li, name in enumerate(c["lesions"]):
            if rs.random_sample() < c["presence"][li]:
                cnt = rs.randint(1, c["max_regions"] + 1)
                vecs = centers[li, lab[i]][None, :] + 1.5 * rs.standard_normal((cnt, c["dl"])).astype(np.float32)
                vecs = (vecs / np.linalg.norm(vecs, axis=1, keepdims=True)).astype(np.float32)
                m[name] = [v for v in vecs]
        maps.append(m)
    return g, lab, maps


def main():
    warnings.filterwarnings("ignore")
    cm = ref_shim.module("ChestMIR.chestmir_eval")
    c = CASE
    g, lab, maps = inputs(c)
    sim = (g @ g.T).astype(np.float32)
    np.fill_diagonal(sim, -np.inf)
    # similarity_to_ranks is an unstable argsort (chestmir_eval.py:429); the oracle form sorts stably
    cm.similarity_to_ranks = lambda s: np.argsort(-s, axis=0, kind="stable")
    G, A = {"case": c}, {}
    for name in c["lesions"]:
        ranks, stats = cm.rerank_with_specific_lesion(sim, maps, name, c["rerank_topk"], c["global_weight"])
        A[f"specific_{name.replace(' ', '_')}"] = ranks[: c["keep"]].T.copy()
        G[f"specific_{name}"] = {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item())
                                 for k, v in stats.items()}
    ranks, stats = cm.rerank_with_adaptive_lesion(sim, maps, c["lesions"], c["rerank_topk"], c["global_weight"])
    A["adaptive"] = ranks[: c["keep"]].T.copy()
    G["adaptive"] = {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item()) for k, v in stats.items()}
    A["base"] = np.argsort(-sim, axis=0, kind="stable")[: c["keep"]].T.copy()
    with open(os.path.join(OUT, "golden_lesion.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "golden_lesion_arrays.npz"), **A)
    print(G["adaptive"], {k: v["queries_reranked"] for k, v in G.items() if k.startswith("specific")},
          "moved rows:", int((A["adaptive"] != A["base"]).any(1).sum()))


if __name__ == "__main__":
    main()
