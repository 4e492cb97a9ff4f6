"""Golden outputs for the OUTPUT side of the dual-collection comparison, by the REAL reference functions.

    python -m oracle.make_golden_export          (build container only: needs /root/reference)

Run UNMODIFIED on a hand-built ``compare_models`` payload: ``export_analysis`` and ``print_summary``
(retrieval_analysis/run_analysis.py:67-108) -- the text of every file they write and of the console summary -- and
``compare_collection_coverage`` (retrieval_analysis/milvus_adapter.py:309-320) over two stand-in adapters.
Output: tests/golden/golden_export.json.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import tempfile

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _side(paths, labels, scores):
    return {"image_paths": paths, "labels": labels, "scores": scores, "distances": scores,
            "hits": [{"id": i, "image_path": p, "label": lab, "score": s, "distance": s}
                     for i, (p, lab, s) in enumerate(zip(paths, labels, scores))]}


def payload():
    """Five analysed queries over three of the four groups (one group stays EMPTY: its CSV has no header), labels with a
    comma / quote / non-ASCII character (CSV quoting, JSON escaping), a missing query and an error entry."""
    rows = []
    spec = [("covid/a.png", "covid", True, True, "both_correct"), ("covid/b,1.png", "normal", False, False, "both_wrong"),
            ("covid/c.png", 'pneu"monia', False, True, "dino_correct_conv_wrong"),
            ("covid/d.png", None, False, False, "both_wrong"), ("covid/é.png", "covid", True, True, "both_correct")]
    for n, (path, label, cc, dc, group) in enumerate(spec):
        rows.append({"query_image_path": path, "query_label": label,
                     "conv": _side([f"g/{n}_{j}.png" for j in range(3)], ["covid", "normal", label], [0.9 - 0.1 * j for j in range(3)]),
                     "dino": _side([f"h/{n}_{j}.png" for j in range(2)], [label, "normal"], [0.75, 0.5]),
                     "conv_correct": cc, "dino_correct": dc, "assigned_group": group})
    return {"coverage": {"present_in_conv_only": ["x/1.png"], "present_in_dino_only": [], "present_in_both": ["covid/a.png", "covid/c.png"]},
            "summary": {"both_correct": 2, "both_wrong": 2, "dino_correct_conv_wrong": 1, "conv_correct_dino_wrong": 0,
                        "evaluated_queries": 5},
            "results": rows, "missing_queries": [{"image_path": "covid/zz.png", "label": "covid"}],
            "errors": [{"image_path": "covid/bad.png", "error": "no embedding"}]}


class _Adapter:
    def __init__(self, paths):
        self._paths = paths

    def list_image_paths(self, batch_size=1000):
        return list(self._paths)


COVERAGE_CASES = [(["b", "a", "c", "a"], ["c", "d"]), ([], ["x"]), (["p"], ["p"])]


def main():
    R = ref_shim.module("retrieval_analysis.run_analysis")
    M = ref_shim.module("retrieval_analysis.milvus_adapter")
    pay = payload()
    gold = {"payload": pay, "files": {}}
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "nested", "out")
        R.export_analysis(pay, out)
        for name in sorted(os.listdir(out)):
            with open(os.path.join(out, name), "r", encoding="utf-8", newline="") as fh:
                gold["files"][name] = fh.read()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        R.print_summary(pay)
    gold["summary_text"] = buf.getvalue()
    gold["coverage"] = [{"conv": c, "dino": d, "result": M.compare_collection_coverage(_Adapter(c), _Adapter(d))}
                        for c, d in COVERAGE_CASES]
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "golden_export.json"), "w", encoding="utf-8") as fh:
        json.dump(gold, fh, indent=1)      # key order preserved: the payload's own order is what the files show
    print("wrote", os.path.join(OUT, "golden_export.json"), sorted(gold["files"]))


if __name__ == "__main__":
    main()
