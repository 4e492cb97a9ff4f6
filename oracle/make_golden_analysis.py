"""Golden outputs for SURVEY 8(a) D13 and its caller, produced by the REAL reference functions.

    python -m oracle.make_golden_analysis          (build container only: needs /root/reference)

``is_retrieval_correct`` / ``CorrectnessConfig`` (retrieval_analysis/evaluator.py:12-26), ``assign_group``
(retrieval_analysis/comparison.py:236-244) and ``compare_models`` (comparison.py:85-234) are run UNMODIFIED; the two
``MilvusCollectionAdapter`` objects talk to a stand-in client (exact numpy cosine search with the deterministic tie
order + the three filter expressions the adapter builds).  Output: tests/golden/golden_analysis.json.
"""
from __future__ import annotations

import json
import os
import re
import warnings

import numpy as np

from . import ref_shim, synth
from .make_golden_collection import _NumpyClient

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(n=240, classes=4, seed_conv=51, seed_dino=52, d_conv=64, d_dino=48, noise_conv=3.5, noise_dino=2.5,
            top_k=5, n_queries=60, missing_in_dino=7)

# (query_label, labels of the hits in order, top_k) -- exercises the falsy-label and empty-result rules
CORRECT_CASES = [
    ("a", ["a", "b"], 1), ("a", ["b", "a"], 1), ("a", ["b", "a"], 2), ("a", [], 3), (None, ["a"], 1), ("", [""], 1),
    ("a", ["b", "c", "d"], 5), ("a", ["b", "c", "a"], 2), ("a", ["b", "c", "a"], 3), ("x", [None, "x"], 2),
]


def inputs():
    """Two embedding sets over the same images (different dims / noise); the last images are absent from 'dino'."""
    lab = np.arange(CASE["n"]) % CASE["classes"]
    np.random.RandomState(7).shuffle(lab)
    out = {}
    for side in ("conv", "dino"):
        rs = np.random.RandomState(CASE[f"seed_{side}"])
        mu = rs.standard_normal((CASE["classes"], CASE[f"d_{side}"]))
        out[side] = (mu[lab] + CASE[f"noise_{side}"] * rs.standard_normal((CASE["n"], CASE[f"d_{side}"]))).astype(np.float32)
    paths = [f"covid/img_{i:04d}.png" for i in range(CASE["n"])]
    labels = [f"class{v}" for v in lab]
    return out["conv"], out["dino"], paths, labels


def query_records(paths, labels, record_cls):
    step = CASE["n"] // CASE["n_queries"]
    qi = list(range(0, CASE["n"], step))[: CASE["n_queries"]]
    qs = [record_cls(image_path=paths[i], label=labels[i] if i % 3 else None) for i in qi]   # some labels come from the store
    qs.append(record_cls(image_path="covid/not_ingested.png", label="class0"))
    return qs


class _Client(_NumpyClient):
    """adds ``query`` for `field == "v"`, `field != "v"`, `field in [..]` (milvus_adapter.py:94-175)"""

    def __init__(self, x, columns):
        super().__init__(x, columns)
        self.raw = x

    def query(self, collection_name=None, filter="", output_fields=None, limit=None, offset=0):
        n = len(self.columns["image_path"])
        m = re.match(r'^\s*(\w+)\s+in\s+\[(.*)\]\s*$', filter, re.S)
        if m:
            vals = set(re.findall(r'"((?:[^"\\]|\\.)*)"', m.group(2)))
            rows = [i for i in range(n) if self.columns[m.group(1)][i] in vals]
        else:
            m = re.match(r'^\s*(\w+)\s*(==|!=)\s*"(.*)"\s*$', filter)
            col, op, val = m.group(1), m.group(2), m.group(3)
            rows = [i for i in range(n) if (self.columns[col][i] == val) == (op == "==")]
        rows = rows[offset:][: limit]
        out = []
        for i in rows:
            ent = {f: self.columns[f][i] for f in (output_fields or self.columns) if f in self.columns}
            if output_fields and "embedding" in output_fields:
                ent["embedding"] = self.raw[i].tolist()
            out.append(ent)
        return out


def main():
    warnings.filterwarnings("ignore")
    ma = ref_shim.module("retrieval_analysis.milvus_adapter")
    ev = ref_shim.module("retrieval_analysis.evaluator")
    cmp_ = ref_shim.module("retrieval_analysis.comparison")
    G = {"case": CASE}

    def items(labels):
        return [ma.RetrievedItem(id=j, image_path=f"p{j}", label=lab, score=1.0 - 0.1 * j, distance=1.0 - 0.1 * j)
                for j, lab in enumerate(labels)]

    G["correct"] = [bool(ev.is_retrieval_correct(q, items(labs), ev.CorrectnessConfig(top_k=k)))
                    for q, labs, k in CORRECT_CASES]
    G["correct_default_top_k"] = ev.CorrectnessConfig().top_k
    G["groups"] = {f"{int(c)}{int(d)}": cmp_.assign_group(bool(c), bool(d)) for c in (0, 1) for d in (0, 1)}

    xc, xd, paths, labels = inputs()
    nd = CASE["n"] - CASE["missing_in_dino"]
    cols = lambda m: {"id": list(range(m)), "image_path": paths[:m], "label": labels[:m]}   # noqa: E731
    adapters = []
    for name, x, m in (("conv", xc, CASE["n"]), ("dino", xd, nd)):
        ad = ma.MilvusCollectionAdapter(ma.MilvusCollectionConfig(name=name, collection_name=name, uri="http://stub"))
        ad.client = _Client(x[:m], cols(m))
        adapters.append(ad)
    for ck in (1, 3):
        cfg = cmp_.ComparisonConfig(top_k=CASE["top_k"], correctness=ev.CorrectnessConfig(top_k=ck), search_batch_size=16)
        res = cmp_.compare_models(adapters[0], adapters[1], query_records(paths, labels, ma.QueryRecord), cfg)
        G[f"compare_top{ck}"] = {
            "summary": res["summary"], "coverage_counts": {k: len(v) for k, v in res["coverage"].items()},
            "missing_queries": res["missing_queries"], "errors": res["errors"],
            "rows": [{"query": r["query_image_path"], "label": r["query_label"], "group": r["assigned_group"],
                      "conv_correct": r["conv_correct"], "dino_correct": r["dino_correct"],
                      "conv_paths": r["conv"]["image_paths"], "dino_paths": r["dino"]["image_paths"],
                      "conv_labels": r["conv"]["labels"]} for r in res["results"]],
        }
    with open(os.path.join(OUT, "golden_analysis.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    print(G["correct"], G["groups"])
    print(G["compare_top1"]["summary"], G["compare_top3"]["summary"], G["compare_top1"]["coverage_counts"])


if __name__ == "__main__":
    main()
