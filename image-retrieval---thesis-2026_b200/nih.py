"""NIH ChestX-ray14 label formats on the INPUT side of the multilabel retrieval path (host-side only).

The file name of an exported NIH image carries its findings (``..._Chest_X-ray_<finding>|<finding>_<n>.npy``, URL-quoted);
the reference turns it into the ``label_names`` / ``multi_hot`` pair every later stage consumes -- the rows
``collection.insert_rows`` stores, the ``query_label_vector`` of the hits JSON (``formats.nih_query_results``), the
multi-hot label tensors of the metric kernels.  Same names and behaviour as nih_zilliz_utils.py:25-133:
``normalize_nih_label``, ``parse_nih_labels_from_path``, ``resolve_npy_paths``, ``build_collection_name``; the 14 findings
are ``NIH_RETRIEVAL_PATHOLOGIES`` (nih_multilabel_retrieval.py).  Pinned by tests/test_nih_labels.py against the real
functions (tests/golden/golden_nih_labels.json).
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Sequence, Tuple
from urllib.parse import unquote

NIH_RETRIEVAL_PATHOLOGIES = [
    "Atelectasis", "Cardiomegaly", "Effusion", "Infiltration", "Mass", "Nodule", "Pneumonia", "Pneumothorax",
    "Consolidation", "Edema", "Emphysema", "Fibrosis", "Pleural Thickening", "Hernia",
]
_FILE_TOKEN = "Chest_X-ray_"
_PLEURAL_SPELLINGS = ("pleural_thickening", "pleural thickening", "pleuralthickening")


def normalize_nih_label(label_name: str) -> str:
    """Lower case, ``%20`` / ``_`` / ``-`` read as blanks (nih_zilliz_utils.py:25-32)."""
    out = label_name.strip()
    for token in ("%20", "_", "-"):
        out = out.replace(token, " ")
    return out.lower()


def parse_nih_labels_from_path(image_path: str, pathology_names: Optional[Sequence[str]] = None
                               ) -> Tuple[List[str], List[float]]:
    """File name -> (finding names in file order, multi-hot over ``pathology_names``) -- nih_zilliz_utils.py:35-78.  A name
    without the ``Chest_X-ray_`` token, or with a finding outside the list (e.g. "No Finding"), is an error: every row of
    the NIH retrieval sets therefore has at least one label."""
    names = list(pathology_names or NIH_RETRIEVAL_PATHOLOGIES)
    column = {name: i for i, name in enumerate(names)}
    canonical = {spelling: "Pleural Thickening" for spelling in _PLEURAL_SPELLINGS}
    canonical.update({normalize_nih_label(name): name for name in names})
    file = Path(image_path)
    at = file.stem.find(_FILE_TOKEN)
    if at < 0:
        raise ValueError(f"Unsupported NIH file name '{file.name}'. Expected token '{_FILE_TOKEN}'.")
    findings, _serial = file.stem[at + len(_FILE_TOKEN):].rsplit("_", 1)
    label_names: List[str] = []
    multi_hot = [0.0] * len(names)
    unknown = []
    for raw in (part.strip() for part in unquote(findings).split("|")):
        name = canonical.get(normalize_nih_label(raw))
        if name is None or name not in column:
            unknown.append(raw)
            continue
        multi_hot[column[name]] = 1.0
        label_names.append(name)
    if unknown:
        raise ValueError(f"Unknown pathologies in '{file.name}': {unknown}.")
    return label_names, multi_hot


def resolve_npy_paths(data_dir: str, image_list_file: Optional[str] = None) -> List[str]:
    """The ``.npy`` images of a run (nih_zilliz_utils.py:103-121): the first comma-separated field of every non-blank
    manifest line (relative names under ``data_dir``), or every ``*.npy`` below ``data_dir``, sorted."""
    if image_list_file:
        paths = []
        with open(Path(image_list_file), "r", encoding="utf-8") as fh:
            for line in fh:
                if not line.strip():
                    continue
                item = Path(line.strip().split(",")[0].strip())
                paths.append(str(item if item.is_absolute() else Path(data_dir) / item))
    else:
        paths = sorted(str(p) for p in Path(data_dir).rglob("*.npy"))
    if not paths:
        raise ValueError("No .npy files found for NIH ingestion/query.")
    return paths


def build_collection_name(model_name: str, suffix: str) -> str:
    """nih_zilliz_utils.py:132-133."""
    return f"nih_{model_name}_{suffix}"
