"""Golden outputs for the collection drop-in row, produced by the REAL reference classes.

    python -m oracle.make_golden_collection          (build container only: needs /root/reference)

The reference's ``MilvusCollectionAdapter`` (retrieval_analysis/milvus_adapter.py) and ``search_collection``
(nih_zilliz_utils.py) are run UNMODIFIED against a stand-in client whose ``search`` is an exact numpy cosine search
with the deterministic tie order; what they return for seeded inputs is recorded in tests/golden/golden_collection.json.
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np

from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(n=300, d=96, classes=5, seed=41, noise=3.0, top_k=7, n_queries=40)
NIH = dict(n=220, nq=25, d=64, seed=43, noise=1.0, top_k=20)


def inputs():
    x, lab = synth.clustered(CASE["n"], CASE["d"], CASE["classes"], CASE["seed"], CASE["noise"])
    paths = [f"data/img_{i:04d}.png" for i in range(CASE["n"])]
    labels = [f"class{v}" for v in lab]
    return x, paths, labels


def nih_inputs():
    lab = synth.multihot(NIH["n"] + NIH["nq"], NIH["seed"])
    emb = synth.labelset_clustered(lab, NIH["d"], NIH["seed"] + 100, NIH["noise"])
    return emb, lab


class _NumpyClient:
    """exact cosine search, best first, ties by ascending row; hit dicts shaped like MilvusClient.search output"""

    def __init__(self, x, columns):
        n = np.linalg.norm(x, axis=1, keepdims=True)
        self.e = (x / np.maximum(n, 1e-12)).astype(np.float32)
        self.columns = columns

    def search(self, collection_name=None, data=None, anns_field=None, search_params=None, limit=10, output_fields=None,
               param=None):
        q = np.asarray(data, dtype=np.float32)
        q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
        s = (q.astype(np.float64) @ self.e.astype(np.float64).T)
        order = np.argsort(-s, axis=1, kind="stable")[:, :limit]
        out = []
        for r in range(q.shape[0]):
            out.append([{"id": int(j), "distance": float(s[r, j]), "score": float(s[r, j]),
                         "entity": {f: self.columns[f][j] for f in (output_fields or self.columns)}} for j in order[r]])
        return out


class _HitObj:
    def __init__(self, d):
        self.id, self.distance, self.entity = d["id"], d["distance"], d["entity"]


class _PymilvusCollection:
    def __init__(self, client):
        self.client = client

    def search(self, data, anns_field, param, limit, output_fields):
        return [[_HitObj(h) for h in hits] for hits in self.client.search(data=data, limit=limit, output_fields=output_fields)]


def main():
    warnings.filterwarnings("ignore")
    ma = ref_shim.module("retrieval_analysis.milvus_adapter")
    nz = ref_shim.module("nih_zilliz_utils")
    ez = ref_shim.module("evaluate_nih_zilliz")
    G = {"case": CASE, "nih": NIH}

    x, paths, labels = inputs()
    cfg = ma.MilvusCollectionConfig(name="conv", collection_name="c", uri="http://stub")
    adapter = ma.MilvusCollectionAdapter(cfg)
    adapter.client = _NumpyClient(x, {"id": list(range(len(paths))), "image_path": paths, "label": labels})
    qi = list(range(0, CASE["n"], CASE["n"] // CASE["n_queries"]))[: CASE["n_queries"]]
    queries = [ma.QueryRecord(image_path=paths[i], label=labels[i]) for i in qi]
    for excl in (True, False):
        res = adapter.search_by_embeddings(queries, [x[i].tolist() for i in qi], CASE["top_k"], exclude_self=excl)
        G[f"adapter_exclude_self_{excl}"] = [
            {"query": r.query.image_path, "source": r.query_source,
             "image_paths": [it.image_path for it in r.retrieved], "labels": [it.label for it in r.retrieved],
             "ids": [it.id for it in r.retrieved], "scores": [it.score for it in r.retrieved]} for r in res]

    emb, lab = nih_inputs()
    g, q = emb[: NIH["n"]], emb[NIH["n"]:]
    cols = {"image_path": [f"nih/{i}.npy" for i in range(NIH["n"])], "image_name": [f"{i}.npy" for i in range(NIH["n"])],
            "label_text": ["|".join(str(j) for j in np.flatnonzero(lab[i])) for i in range(NIH["n"])],
            "label_vector_json": [json.dumps(lab[i].astype(int).tolist()) for i in range(NIH["n"])]}
    coll = _PymilvusCollection(_NumpyClient(g, cols))
    items = []
    for r in range(NIH["nq"]):
        hits = nz.search_collection(coll, q[r].tolist(), NIH["top_k"])
        items.append({"query_label_vector": lab[NIH["n"] + r].astype(int).tolist(), "results": hits})
    G["nih_hits_first_query"] = items[0]["results"][:5]
    G["nih_hit_ids"] = [[h["id"] for h in it["results"]] for it in items]
    G["nih_metrics"] = {k: float(v) for k, v in ez.evaluate_results(items, 0.4, [1, 5, 10, 20]).items()}

    with open(os.path.join(OUT, "golden_collection.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    print(json.dumps(G["adapter_exclude_self_True"][0], indent=1)[:800])
    print(G["nih_metrics"])


if __name__ == "__main__":
    main()
