"""Golden outputs of the REAL ChestMIR format helpers (ChestMIR/chestmir_eval.py:46-121, 275-321, 653-667) for
tests/test_chestmir_formats.py.

    python -m oracle.make_golden_chestmir_formats          (build container only: needs /root/reference)
"""
from __future__ import annotations

import contextlib
import io
import json
import os

import numpy as np

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

JSON_CASES = [None, "", "[]", "[1, 2]", '["a", null]', "{\"a\": 1}", "3", "not json", "[1, 2", '"str"', "null"]
MAP_CASES = [
    (json.dumps(["Nodule/Mass", "effusion", "Mass", "Lung_Opacity", "unknown thing", "ILD"]),
     json.dumps([[3.0, 4.0], [0.0, 2.0], [1.0, 0.0], [1.0, 1.0], [0.5, 0.5], [0.0, 0.0]])),
    (json.dumps(["edema", "edema", "Interstitial lung disease"]), json.dumps([[1, 0, 0], [], [0, 0, 2]])),   # empty vector
    (json.dumps(["cyst", "cavity"]), json.dumps([[1.0, 2.0, 2.0]])),                                         # ragged lists
    (json.dumps(["Pleural-Thickening", "x"]), json.dumps(["not a list", 5])),
    ("", json.dumps([[1.0]])), (None, None), ("broken", "[[1.0]]"),
]
NAMES = ["Nodule/Mass", " Pleural_Effusion ", "Plural  effusion", "INFILTRATES", "lung-opacity", "Enlarged PA", "other", 7]
REPORT = {"R@K": {1: 81.256, 5: 93.5, 10: 100.0}, "mAP": 45.6789, "mP@K": {1: 81.256, 5: 60.0049, 10: 41.005},
          "classification": {1: {"accuracy": 80.0, "precision_macro": 79.995, "recall_macro": 70.0, "f1_macro": 74.444},
                             5: {"accuracy": 85.5, "precision_macro": 84.0, "recall_macro": 77.125, "f1_macro": 80.0}}}


def main():
    cm = ref_shim.module("ChestMIR.chestmir_eval")
    gold = {"alias_groups": cm.LESION_ALIAS_GROUPS, "alias_to_canon": cm.LESION_ALIAS_TO_CANON,
            "json_cases": [[c, cm.parse_json_list(c)] for c in JSON_CASES],
            "names": [[n, cm.canonical_lesion_name(n)] for n in NAMES], "maps": []}
    for labels_json, vectors_json in MAP_CASES:
        m = cm.build_lesion_vector_map(labels_json, vectors_json)
        gold["maps"].append({"labels_json": labels_json, "vectors_json": vectors_json,
                             "map": {k: [v.tolist() for v in vs] for k, vs in m.items()},
                             "dtypes": sorted({str(v.dtype) for vs in m.values() for v in vs})})
    rs = np.random.RandomState(5)
    x = rs.standard_normal((6, 5)).astype(np.float32)
    x[2] = 0.0
    gold["normalize_rows"] = {"x": x.tolist(), "y": cm.normalize_rows(x).tolist(), "dtype": str(cm.normalize_rows(x).dtype),
                              "y64": cm.normalize_rows(x.astype(np.float64), eps=1e-6).tolist()}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        cm.print_stage_report("Stage 1: global", REPORT, [1, 5, 10], [1, 5])
    gold["report"] = {"title": "Stage 1: global", "kappas": [1, 5, 10], "cls_k_values": [1, 5], "text": buf.getvalue(),
                      "R@K": [[k, v] for k, v in REPORT["R@K"].items()], "mAP": REPORT["mAP"],
                      "mP@K": [[k, v] for k, v in REPORT["mP@K"].items()],
                      "classification": [[k, v] for k, v in REPORT["classification"].items()]}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "golden_chestmir_formats.json"), "w", encoding="utf-8") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote golden_chestmir_formats.json")


if __name__ == "__main__":
    main()
