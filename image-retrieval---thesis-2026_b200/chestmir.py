"""ChestMIR evaluation: the data formats either side of the lesion re-ranking path (host-side only).

The gallery rows of the ChestMIR collections carry their region vectors as two JSON strings (``region_labels_json``,
``region_vectors_json``); the reference turns them into one ``{canonical lesion -> [unit vectors]}`` map per image, which
is what ``fusion.LesionIndex`` / ``fusion.lesion_rerank_search`` consume, and prints a per-stage report of the metric
dictionaries.  Same names and behaviour as ChestMIR/chestmir_eval.py: ``LESION_ALIAS_GROUPS`` / ``LESION_ALIAS_TO_CANON``
(:46-121), ``normalize_rows`` (:275-278), ``parse_json_list`` (:281-288), ``canonical_lesion_name`` (:298-300),
``build_lesion_vector_map`` (:303-321), ``evaluate_rankings`` (:434-448) and ``print_stage_report`` (:653-667).
Pinned by tests/test_chestmir_formats.py against the real functions (tests/golden/golden_chestmir_formats.json).
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .fusion import normalize_lesion_text

# canonical finding -> its other spellings (after normalize_lesion_text; spellings with `_` or `/` can never match a
# normalised name and are kept only so that the table equals the reference's)
_OTHER_SPELLINGS = {
    "consolidation": (), "lung opacity": ("lung_opacity", "opacity", "opacities"),
    "infiltration": ("infiltrate", "infiltrates"), "atelectasis": ("atelectatic",),
    "pleural effusion": ("pleural_effusion", "effusion", "plural effusion"),
    "nodule mass": ("nodule/mass", "nodule_mass", "mass", "nodule"), "cardiomegaly": (), "edema": (), "pneumothorax": (),
    "pleural thickening": ("pleural_thickening",), "pulmonary fibrosis": ("pulmonary_fibrosis", "fibrosis"),
    "enlarged pa": ("enlarged_pa",), "ild": ("interstitial lung disease",), "calcification": (),
    "lung cavity": ("lung_cavity", "cavity"), "lung cyst": ("lung_cyst", "cyst"),
}
LESION_ALIAS_GROUPS: Dict[str, List[str]] = {canon: [canon, *others] for canon, others in _OTHER_SPELLINGS.items()}
LESION_ALIAS_TO_CANON: Dict[str, str] = {alias: canon for canon, aliases in LESION_ALIAS_GROUPS.items() for alias in aliases}


def normalize_rows(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """numpy ``x / max(|x|, eps)`` row-wise (the device form is ``b200knn.normalize``)."""
    return x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), eps)


def parse_json_list(raw) -> List[Any]:
    """A JSON string holding a list -> that list; anything else (None, "", invalid JSON, another JSON type) -> []."""
    if raw is None or raw == "":
        return []
    try:
        value = json.loads(raw)
    except Exception:  # noqa: BLE001 - the reference swallows every parse error
        return []
    return value if isinstance(value, list) else []


def canonical_lesion_name(name) -> str:
    """Normalised name mapped through the alias table (``fusion.canonical_lesion_name`` with the reference's table)."""
    text = normalize_lesion_text(name)
    return LESION_ALIAS_TO_CANON.get(text, text)


def build_lesion_vector_map(region_labels_json, region_vectors_json) -> Dict[str, List[np.ndarray]]:
    """The two JSON columns of one gallery row -> {canonical lesion -> [fp32 unit vectors, in row order]}; pairs beyond
    the shorter list, empty / non-list vectors and zero vectors are dropped."""
    out: Dict[str, List[np.ndarray]] = {}
    for label, raw in zip(parse_json_list(region_labels_json), parse_json_list(region_vectors_json)):
        if not isinstance(raw, list) or not raw:
            continue
        vec = np.asarray(raw, dtype=np.float32)
        norm = np.linalg.norm(vec)
        if norm <= 0:
            continue
        out.setdefault(canonical_lesion_name(label), []).append(vec / norm)
    return out


def evaluate_rankings(ranks: np.ndarray, labels: np.ndarray, kappas: Sequence[int], cls_k_values: Sequence[int],
                      device=None) -> Dict[str, Any]:
    """The metric bundle of one ranking stage: R@K, mAP, mP@K (percent) and the majority-vote classification table,
    computed by the device metric functions behind their ``*_from_ranks`` names."""
    from . import metrics as M

    acc = M.retrieval_accuracy_from_ranks(ranks, labels, kappas, device=device)
    m_ap, _aps, pr, _prs = M.compute_map(ranks, _codes(labels), list(kappas))
    cls = M.compute_classification_metrics_from_ranks(labels, ranks, cls_k_values, device=device)
    return {"R@K": {k: float(v) for k, v in zip(kappas, acc)}, "mAP": float(m_ap * 100.0),
            "mP@K": {k: float(v * 100.0) for k, v in zip(kappas, pr)}, "classification": cls}


def _codes(labels) -> np.ndarray:
    """compute_map compares labels for equality only: any labels (strings included) -> integer codes."""
    return np.unique(np.asarray(labels), return_inverse=True)[1].reshape(-1)


def stage_report_text(title: str, report: Dict[str, Any], kappas: Sequence[int], cls_k_values: Sequence[int]) -> str:
    """What ``print_stage_report`` prints, as one string."""
    lines = [f"\n=== {title} ===", ", ".join(f"R@{k}: {report['R@K'][k]:.2f}%" for k in kappas),
             f"mAP: {report['mAP']:.2f}%", ", ".join(f"P@{k}: {report['mP@K'][k]:.2f}%" for k in kappas)]
    for k in cls_k_values:
        m = report["classification"][k]
        lines.append(f"Top-{k}: Acc {m['accuracy']:.2f}% | P_macro {m['precision_macro']:.2f}% | "
                     f"R_macro {m['recall_macro']:.2f}% | F1_macro {m['f1_macro']:.2f}%")
    return "\n".join(lines) + "\n"


def print_stage_report(title: str, report: Dict[str, Any], kappas: Sequence[int], cls_k_values: Sequence[int]) -> None:
    print(stage_report_text(title, report, kappas, cls_k_values), end="")
