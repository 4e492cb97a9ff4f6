"""CPU: the output side of the dual-collection comparison -- ``formats.export_analysis`` / ``comparison_summary_text`` /
``analysis.compare_collection_coverage`` -- byte for byte against the files and the console text the REAL reference
functions produced for the same payload (tests/golden/golden_export.json, oracle/make_golden_export.py;
retrieval_analysis/run_analysis.py:67-108, milvus_adapter.py:309-320)."""
import importlib
import json
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_export.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as fh:
        return json.load(fh)


def test_export_analysis_writes_the_reference_files(gold, tmp_path):
    F = importlib.import_module("b200knn.formats")
    out = tmp_path / "nested" / "out"
    written = F.export_analysis(gold["payload"], out)
    assert sorted(os.path.basename(p) for p in written) == sorted(gold["files"]) == sorted(os.listdir(out))
    for name, want in gold["files"].items():
        with open(out / name, "r", encoding="utf-8", newline="") as fh:
            assert fh.read() == want, name
    assert gold["files"]["group_conv_correct_dino_wrong.csv"] == "\r\n"        # an empty group still gets its file


def test_summary_text_and_print(gold, capsys):
    F = importlib.import_module("b200knn.formats")
    assert F.comparison_summary_text(gold["payload"]) == gold["summary_text"]
    F.print_summary(gold["payload"])
    assert capsys.readouterr().out == gold["summary_text"]


def test_collection_coverage(gold):
    A = importlib.import_module("b200knn.analysis")

    class Adapter:
        def __init__(self, paths):
            self.paths = paths

        def list_image_paths(self, batch_size=1000):
            return list(self.paths)

    for case in gold["coverage"]:
        assert A.compare_collection_coverage(Adapter(case["conv"]), Adapter(case["dino"])) == case["result"]


def test_write_csv_accepts_a_generator(tmp_path):
    F = importlib.import_module("b200knn.formats")
    p = F.write_csv(str(tmp_path / "g.csv"), ({"a": i, "b": 2 * i} for i in range(3)))
    assert open(p).read().splitlines() == ["a,b", "0,0", "1,2", "2,4"]
