"""Golden outputs of the REAL ``PathMapper`` (milvus/path_mapper.py:10-107) for tests/test_paths.py.

    python -m oracle.make_golden_paths          (build container only: needs /root/reference)
"""
from __future__ import annotations

import json
import os

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PATHS = ["/kaggle/input/rsna-png/data/train/SARS-10.1148rg.242035193-g04mr34g0-Fig8a-day0.jpeg",
         "/kaggle/input/ds/x.png", "/kaggle/input/ds", "/kaggle/working/out/y.png", "/data/local/z.png", "plain.png",
         "input/a/b/c.png", "/kaggle/input/ds/dir/", ""]
BASES = ["/media/user/datasets/covidx-cxr/data/train", "relative/base", "/trailing/"]


def main():
    P = ref_shim.module("milvus.path_mapper")
    gold = {"paths": PATHS, "bases": BASES, "cases": []}
    for base in BASES:
        m = P.PathMapper(local_base_path=base)
        gold["cases"].append({"base": base, "filename": [m.extract_filename(p) for p in PATHS],
                              "relative": [m.extract_relative_path(p) for p in PATHS],
                              "remap": [m.remap_path(p) for p in PATHS], "batch": m.batch_remap(PATHS),
                              "override": [m.remap_path(p, "/other") for p in PATHS],
                              "verify": [list(m.verify_path(p)) for p in PATHS[:3]]})
    try:
        P.PathMapper().remap_path(PATHS[0])
        gold["no_base"] = None
    except Exception as exc:  # noqa: BLE001
        gold["no_base"] = {"type": type(exc).__name__, "message": str(exc)}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "golden_paths.json"), "w", encoding="utf-8") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote golden_paths.json")


if __name__ == "__main__":
    main()
