"""CPU: ``collection.PathMapper`` against the REAL reference class (tests/golden/golden_paths.json,
oracle/make_golden_paths.py; milvus/path_mapper.py:10-107) and the path re-rooting of ``LocalRetrieverPatched``
(milvus/milvus_retrieval_patched.py:31-42, 96-121)."""
import importlib
import json
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_paths.json")


def test_path_mapper_matches_the_reference():
    C = importlib.import_module("b200knn.collection")
    with open(GOLDEN) as fh:
        gold = json.load(fh)
    paths = gold["paths"]
    for case in gold["cases"]:
        m = C.PathMapper(local_base_path=case["base"])
        assert [m.extract_filename(p) for p in paths] == case["filename"]
        assert [m.extract_relative_path(p) for p in paths] == case["relative"]
        assert [m.remap_path(p) for p in paths] == case["remap"] == m.batch_remap(paths) == case["batch"]
        assert [m.remap_path(p, "/other") for p in paths] == case["override"]
        assert [list(m.verify_path(p)) for p in paths[:3]] == case["verify"]
    with pytest.raises(Exception) as info:
        C.PathMapper().remap_path(paths[0])
    assert type(info.value).__name__ == gold["no_base"]["type"] and str(info.value) == gold["no_base"]["message"]


def test_patched_retriever_re_roots_only_kaggle_paths(tmp_path):
    C = importlib.import_module("b200knn.collection")
    hits = lambda: [{"id": 1, "image_path": "/kaggle/input/ds/train/a.png", "label": "x", "distance": 0.9, "similarity": 0.9},  # noqa: E731
                    {"id": 2, "image_path": "/data/local/b.png", "label": "y", "distance": 0.8, "similarity": 0.8}]
    r = C.LocalRetrieverPatched(None, local_data_base_path=str(tmp_path))
    out = r._remap_results(hits())
    assert [h["image_path"] for h in out] == [os.path.join(str(tmp_path), "a.png"), "/data/local/b.png"]
    assert out[0]["id"] == 1 and out[0]["similarity"] == 0.9                      # everything else untouched
    for off in (C.LocalRetrieverPatched(None, local_data_base_path=str(tmp_path), enable_path_mapping=False),
                C.LocalRetrieverPatched(None)):
        assert off.path_mapper is None
        assert [h["image_path"] for h in off._remap_results(hits())] == [h["image_path"] for h in hits()]
    assert issubclass(C.LocalRetrieverPatched, C.LocalRetriever)


@pytest.mark.gpu
def test_patched_retriever_search_on_the_device():
    """Same hits as the plain retriever, stored ``/kaggle/`` paths re-rooted under the local base (verified on a B200)."""
    import numpy as np

    C = importlib.import_module("b200knn.collection")
    rs = np.random.RandomState(0)
    x = rs.standard_normal((50, 16)).astype(np.float32)
    paths = [f"/kaggle/input/ds/train/img_{i}.png" if i % 2 == 0 else f"/data/local/img_{i}.png" for i in range(50)]
    coll = C.LocalCollection("c", 16, "COSINE")
    coll.insert([{"image_path": p, "label": f"l{i % 3}", "embedding": v} for i, (p, v) in enumerate(zip(paths, x))])
    plain, _ = C.LocalRetriever(coll).search(x[4], top_k=6)
    patched, _ = C.LocalRetrieverPatched(coll, local_data_base_path="/mnt/data").search(x[4], top_k=6)
    assert [h["id"] for h in plain] == [h["id"] for h in patched] and plain[0]["id"] == 4
    for a, b in zip(plain, patched):
        want = "/mnt/data/" + a["image_path"].rsplit("/", 1)[1] if a["image_path"].startswith("/kaggle/") else a["image_path"]
        assert b["image_path"] == want and b["similarity"] == a["similarity"] and b["label"] == a["label"]
    off, _ = C.LocalRetrieverPatched(coll, local_data_base_path="/mnt/data", enable_path_mapping=False).search(x[4], top_k=6)
    assert off == plain
