"""CPU: the NIH label / path helpers (b200knn.nih) against the REAL reference functions' outputs
(tests/golden/golden_nih_labels.json, oracle/make_golden_nih_labels.py; nih_zilliz_utils.py:25-133)."""
import importlib
import json
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_nih_labels.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def N():
    return importlib.import_module("b200knn.nih")


def _same(fn, want):
    if "ok" in want:
        got = fn()
        assert (list(got) if isinstance(got, tuple) else got) == want["ok"]
    else:
        with pytest.raises(Exception) as info:
            fn()
        assert type(info.value).__name__ == want["type"] and str(info.value) == want["message"]


def test_pathologies_and_normalisation(N, gold):
    assert N.NIH_RETRIEVAL_PATHOLOGIES == gold["pathologies"] and len(gold["pathologies"]) == 14
    for raw, want in gold["normalize"].items():
        assert N.normalize_nih_label(raw) == want
    assert N.build_collection_name("dinov2", "gallery") == gold["collection_name"]


def test_labels_from_file_names(N, gold):
    for name, want in gold["parse"].items():
        _same(lambda: N.parse_nih_labels_from_path(name), want)
    names = list(gold["parse"])
    _same(lambda: N.parse_nih_labels_from_path(names[1], ["Edema", "Mass"]), gold["parse_subset"])
    _same(lambda: N.parse_nih_labels_from_path(names[0], ["Mass", "Edema"]), gold["parse_subset_ok"])


def test_resolve_npy_paths(N, gold, tmp_path):
    for rel in gold["tree"]:
        (tmp_path / rel).parent.mkdir(parents=True, exist_ok=True)
        (tmp_path / rel).touch()
    (tmp_path / "list.csv").write_text(gold["manifest"], encoding="utf-8")
    strip = lambda ps: [p.replace(str(tmp_path) + os.sep, "<tmp>/") for p in ps]  # noqa: E731
    assert strip(N.resolve_npy_paths(str(tmp_path), str(tmp_path / "list.csv"))) == gold["resolve_manifest"]
    assert strip(N.resolve_npy_paths(str(tmp_path))) == gold["resolve_tree"]
    empty = tmp_path / "z" / "y" / "none"
    empty.mkdir(parents=True)
    _same(lambda: N.resolve_npy_paths(str(empty)), gold["resolve_empty"])
