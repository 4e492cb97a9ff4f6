"""SURVEY 8(a) D13 (`is_retrieval_correct`, retrieval_analysis/evaluator.py:12-26) and the dual-collection comparison
around it (retrieval_analysis/comparison.py:85-244) against golden outputs of the REAL reference functions
(oracle/make_golden_analysis.py)."""
import json
import os

import pytest

from oracle import make_golden_analysis as mga

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ga():
    with open(os.path.join(GOLDEN, "golden_analysis.json")) as fh:
        return json.load(fh)


def _items(labels):
    from b200knn.collection import RetrievedItem

    return [RetrievedItem(id=j, image_path=f"p{j}", label=lab, score=1.0 - 0.1 * j, distance=1.0 - 0.1 * j)
            for j, lab in enumerate(labels)]


def test_is_retrieval_correct_matches_the_reference(ga):
    from b200knn import metrics
    from b200knn.analysis import CorrectnessConfig, is_retrieval_correct

    assert CorrectnessConfig().top_k == ga["correct_default_top_k"]
    got = [is_retrieval_correct(q, _items(labs), CorrectnessConfig(top_k=k)) for q, labs, k in mga.CORRECT_CASES]
    assert got == ga["correct"]
    # the falsy-label and empty-result rules, spelled out (evaluator.py:24-25)
    assert is_retrieval_correct(None, _items(["a"]), CorrectnessConfig()) is False
    assert is_retrieval_correct("", _items([""]), CorrectnessConfig()) is False
    assert is_retrieval_correct("a", [], CorrectnessConfig(top_k=3)) is False
    # the metric-row re-export has the reference signature too
    assert metrics.is_retrieval_correct("a", _items(["b", "a"]), CorrectnessConfig(top_k=2)) is True
    assert metrics.is_retrieval_correct("a", _items(["b", "a"])) is False


def test_assign_group_matches_the_reference(ga):
    from b200knn.analysis import assign_group

    for c in (0, 1):
        for d in (0, 1):
            assert assign_group(bool(c), bool(d)) == ga["groups"][f"{c}{d}"]


@pytest.mark.gpu
@pytest.mark.parametrize("ck", [1, 3])
def test_compare_models_matches_the_reference(ga, ck):
    from b200knn.analysis import ComparisonConfig, CorrectnessConfig, compare_models
    from b200knn.collection import CollectionConfig, LocalCollection, LocalCollectionAdapter, QueryRecord

    xc, xd, paths, labels = mga.inputs()
    nd = mga.CASE["n"] - mga.CASE["missing_in_dino"]
    adapters = []
    for name, x, m in (("conv", xc, mga.CASE["n"]), ("dino", xd, nd)):
        coll = LocalCollection(name, x.shape[1], "COSINE")
        coll.insert([{"image_path": p, "label": l, "embedding": v} for p, l, v in zip(paths[:m], labels[:m], x[:m])])
        adapters.append(LocalCollectionAdapter(CollectionConfig(name, name), coll))
    cfg = ComparisonConfig(top_k=mga.CASE["top_k"], correctness=CorrectnessConfig(top_k=ck), search_batch_size=16)
    res = compare_models(adapters[0], adapters[1], mga.query_records(paths, labels, QueryRecord), cfg)
    want = ga[f"compare_top{ck}"]
    assert res["summary"] == want["summary"]
    assert {k: len(v) for k, v in res["coverage"].items()} == want["coverage_counts"]
    assert res["missing_queries"] == want["missing_queries"] and res["errors"] == want["errors"]
    assert len(res["results"]) == len(want["rows"])
    for r, w in zip(res["results"], want["rows"]):
        assert (r["query_image_path"], r["query_label"], r["assigned_group"]) == (w["query"], w["label"], w["group"])
        assert (r["conv_correct"], r["dino_correct"]) == (w["conv_correct"], w["dino_correct"])
        assert r["conv"]["image_paths"] == w["conv_paths"] and r["dino"]["image_paths"] == w["dino_paths"]
        assert r["conv"]["labels"] == w["conv_labels"]
