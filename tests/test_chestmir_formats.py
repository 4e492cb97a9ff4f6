"""CPU: the ChestMIR data-format helpers (b200knn.chestmir) against the REAL reference functions
(tests/golden/golden_chestmir_formats.json, oracle/make_golden_chestmir_formats.py; ChestMIR/chestmir_eval.py:46-121,
275-321, 434-448, 653-667)."""
import importlib
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_chestmir_formats.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def CM():
    return importlib.import_module("b200knn.chestmir")


def test_alias_table_and_names(CM, gold):
    assert CM.LESION_ALIAS_GROUPS == gold["alias_groups"] and CM.LESION_ALIAS_TO_CANON == gold["alias_to_canon"]
    for name, want in gold["names"]:
        assert CM.canonical_lesion_name(name) == want


def test_parse_json_list(CM, gold):
    for raw, want in gold["json_cases"]:
        assert CM.parse_json_list(raw) == want, raw


def test_build_lesion_vector_map(CM, gold):
    for case in gold["maps"]:
        got = CM.build_lesion_vector_map(case["labels_json"], case["vectors_json"])
        assert list(got) == list(case["map"])                                    # lesion order = first appearance
        for name, vecs in case["map"].items():
            assert len(got[name]) == len(vecs)
            for g, w in zip(got[name], vecs):
                assert g.dtype == np.float32 and np.array_equal(g, np.asarray(w, dtype=np.float32))


def test_normalize_rows(CM, gold):
    c = gold["normalize_rows"]
    x = np.asarray(c["x"], dtype=np.float32)
    y = CM.normalize_rows(x)
    assert str(y.dtype) == c["dtype"] and np.array_equal(y, np.asarray(c["y"], dtype=y.dtype))
    assert np.array_equal(CM.normalize_rows(x.astype(np.float64), eps=1e-6), np.asarray(c["y64"]))


def _report(c):
    return {"R@K": {k: v for k, v in c["R@K"]}, "mAP": c["mAP"], "mP@K": {k: v for k, v in c["mP@K"]},
            "classification": {k: v for k, v in c["classification"]}}


def test_stage_report_text(CM, gold, capsys):
    c = gold["report"]
    assert CM.stage_report_text(c["title"], _report(c), c["kappas"], c["cls_k_values"]) == c["text"]
    CM.print_stage_report(c["title"], _report(c), c["kappas"], c["cls_k_values"])
    assert capsys.readouterr().out == c["text"]


def test_evaluate_rankings_shapes_the_metric_bundle(CM, monkeypatch):
    """chestmir_eval.py:434-448: the three metric functions (GPU-tested against their goldens elsewhere) are replaced by
    stand-ins; what is checked here is the bundle: keys, percent scaling, string labels reaching compute_map as codes."""
    M = importlib.import_module("b200knn.metrics")
    seen = {}
    monkeypatch.setattr(M, "retrieval_accuracy_from_ranks", lambda r, lab, ks, device=None: np.array([50.0, 75.0]))

    def fake_map(ranks, gnd, kappas):
        seen["gnd"] = np.asarray(gnd)
        return 0.4321, None, np.array([0.5, 0.25]), None

    monkeypatch.setattr(M, "compute_map", fake_map)
    monkeypatch.setattr(M, "compute_classification_metrics_from_ranks", lambda lab, r, ks, device=None: {1: {"accuracy": 1.0}})
    labels = np.array(["b", "a", "b", "c"], dtype=object)
    out = CM.evaluate_rankings(np.zeros((4, 4), dtype=np.int64), labels, [1, 5], [1])
    assert out == {"R@K": {1: 50.0, 5: 75.0}, "mAP": pytest.approx(43.21), "mP@K": {1: 50.0, 5: 25.0},
                   "classification": {1: {"accuracy": 1.0}}}
    assert seen["gnd"].tolist() == [1, 0, 1, 2] and np.issubdtype(seen["gnd"].dtype, np.integer)
