"""Golden outputs of the REAL NIH label helpers (nih_zilliz_utils.py:25-133) for tests/test_nih_labels.py.

    python -m oracle.make_golden_nih_labels          (build container only: needs /root/reference)
"""
from __future__ import annotations

import json
import os
import tempfile

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

NAMES = [
    "/data/nih/00001_Chest_X-ray_Edema_12.npy", "q/0002_Chest_X-ray_Pleural_Thickening%7CEdema_3.npy",
    "0003_Chest_X-ray_pleural-thickening|MASS_0.npy", "0004_Chest_X-ray_PleuralThickening_9.npy",
    "a/b/Chest_X-ray_Effusion%7CInfiltration%7CAtelectasis_123.npy", "x_Chest_X-ray_ Hernia | Nodule _7.npy",
    "x_Chest_X-ray_Cardiomegaly|Cardiomegaly_1.npy", "x_Chest_X-ray_Edema%20_5.npy",
    "x_Chest_X-ray_No Finding_5.npy", "x_Chest_X-ray_Edema|Covid_5.npy", "plain_name_5.npy", "x_Chest_X-ray_Edema.npy",
    "deep/Chest_X-ray_Pneumonia_1_Chest_X-ray_Mass_2.npy",
]
LABELS = ["  Pleural_Thickening ", "pleural-thickening", "No%20Finding", "MASS", "Pleural  Thickening"]
MANIFEST = "a/one.npy,0\n\n  /abs/two.npy , 1, extra\nthree.npy\n"
TREE = ["z/b.npy", "a.npy", "z/a.npy", "note.txt", "z/y/c.npy"]


def _try(fn):
    try:
        return {"ok": fn()}
    except Exception as exc:  # noqa: BLE001 - type and message are the golden
        return {"type": type(exc).__name__, "message": str(exc)}


def main():
    U = ref_shim.module("nih_zilliz_utils")
    gold = {"pathologies": list(U.NIH_RETRIEVAL_PATHOLOGIES),
            "normalize": {x: U.normalize_nih_label(x) for x in LABELS},
            "parse": {n: _try(lambda n=n: list(U.parse_nih_labels_from_path(n))) for n in NAMES},
            "parse_subset": _try(lambda: list(U.parse_nih_labels_from_path(NAMES[1], ["Edema", "Mass"]))),
            "parse_subset_ok": _try(lambda: list(U.parse_nih_labels_from_path(NAMES[0], ["Mass", "Edema"]))),
            "collection_name": U.build_collection_name("dinov2", "gallery"),
            "manifest": MANIFEST, "tree": TREE}
    with tempfile.TemporaryDirectory() as tmp:
        for rel in TREE:
            os.makedirs(os.path.dirname(os.path.join(tmp, rel)), exist_ok=True)
            open(os.path.join(tmp, rel), "w").close()
        with open(os.path.join(tmp, "list.csv"), "w", encoding="utf-8") as fh:
            fh.write(MANIFEST)
        strip = lambda ps: [p.replace(tmp + os.sep, "<tmp>/") for p in ps]  # noqa: E731
        gold["resolve_manifest"] = strip(U.resolve_npy_paths(tmp, os.path.join(tmp, "list.csv")))
        gold["resolve_tree"] = strip(U.resolve_npy_paths(tmp))
        empty = os.path.join(tmp, "z", "y", "none")
        os.makedirs(empty)
        gold["resolve_empty"] = _try(lambda: U.resolve_npy_paths(empty))
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "golden_nih_labels.json"), "w", encoding="utf-8") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote golden_nih_labels.json")


if __name__ == "__main__":
    main()
