"""Result / wire formats of the reference, emitted from top-k results.

* the ``np.savez`` bundle of ``test.py:1112-1126`` (``embeds, labels, dists, kappas, acc, mAP, pr,
  classification_k_values, classification_k{k}``) that ``compute_saliency.py:89-100`` reads back -- with the dense
  ``dists`` matrix for small N (compatibility) or a sparse top-k form (``topk_dists``, ``topk_idx``) that never needs
  the N x N matrix, plus the reader-side ``rank_retrieval`` for both;
* the NIH hits JSON of ``query_nih_zilliz.py:58-75`` -> ``evaluate_nih_zilliz.py:34-64``;
* the per-query comparison rows of ``retrieval_analysis/comparison.py:247-284`` and their CSV flattening,
  ``retrieval_analysis/export_utils.py:30-62``;
* the embeddings ``.npz`` of ``fusion_eval/evaluate.py:217-229``.
"""
from __future__ import annotations

import csv
import json
from pathlib import Path
from typing import Any, Dict, List, Mapping, Sequence

import numpy as np


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def save_evaluation_npz(path: str, embeds, labels, kappas, acc, mAP, pr, classification_results: Mapping[int, Mapping],
                        dists=None, topk_dists=None, topk_idx=None) -> str:
    """test.py:1112-1126.  ``dists`` (if given) is the engine's score matrix in the reference's internal convention
    (larger = closer, diagonal -inf); it is stored NEGATED like the reference does (``dists=-dists``: positive
    distances, +inf diagonal).  Without it the top-k pair ``(topk_dists, topk_idx)`` is stored instead, distances
    ascending."""
    payload: Dict[str, Any] = dict(embeds=_np(embeds), labels=_np(labels), kappas=np.asarray(kappas), acc=_np(acc),
                                   mAP=mAP, pr=_np(pr),
                                   classification_k_values=list(classification_results.keys()))
    for k, v in classification_results.items():
        payload[f"classification_k{k}"] = np.array(list(v.values()))
    if dists is not None:
        payload["dists"] = -_np(dists)
    else:
        if topk_dists is None or topk_idx is None:
            raise ValueError("either dists or (topk_dists, topk_idx) is needed")
        payload["topk_dists"], payload["topk_idx"] = _np(topk_dists), _np(topk_idx)
    np.savez(path, **payload)
    return path if str(path).endswith(".npz") else str(path) + ".npz"


def rank_retrieval(results: Mapping[str, np.ndarray], topk: int = 1):
    """compute_saliency.py:19-29 on a loaded bundle: ``(pred labels [N, topk], idx [N, topk])``.  Dense bundles are
    ranked like the reference -- the query's own column dropped (its NaN sorts last), ascending distance, stable tie
    order -- by the library (``knn_rank_rows`` on the device, not numpy); sparse bundles just slice the stored ranking."""
    labels = np.asarray(results["labels"])
    if "dists" in results:
        import torch

        from .search import rank_rows

        d = torch.as_tensor(np.asarray(results["dists"], dtype=np.float32)).cuda()
        d.fill_diagonal_(float("inf"))                      # np.argsort puts the NaN diagonal last
        idx = rank_rows(d, largest_first=False)[:, :topk].cpu().numpy()
    else:
        idx = np.asarray(results["topk_idx"])[:, :topk]
    return labels[idx], idx


def nih_query_results(query_rows: Sequence[Mapping[str, Any]], hits_per_query: Sequence[Sequence[Mapping[str, Any]]]):
    """query_nih_zilliz.py:58-72: one item per query with its hit list (``search_collection`` output)."""
    return [{"query_image_path": row["image_path"], "query_image_name": row["image_name"],
             "query_label_names": row["label_names"], "query_label_vector": row["multi_hot"], "results": list(hits)}
            for row, hits in zip(query_rows, hits_per_query)]


def write_json(path: str, payload, indent: int = 2) -> str:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w", encoding="utf-8") as fh:
        json.dump(payload, fh, indent=indent)
    return path


def serialize_search_result(result) -> Dict[str, Any]:
    """comparison.py:267-284."""
    return {"image_paths": [it.image_path for it in result.retrieved], "labels": [it.label for it in result.retrieved],
            "scores": [it.score for it in result.retrieved], "distances": [it.distance for it in result.retrieved],
            "hits": [{"id": it.id, "image_path": it.image_path, "label": it.label, "score": it.score,
                      "distance": it.distance} for it in result.retrieved]}


def build_query_analysis_row(query, conv_result, dino_result, conv_correct: bool, dino_correct: bool,
                             assigned_group: str) -> Dict[str, Any]:
    """comparison.py:247-264."""
    return {"query_image_path": query.image_path, "query_label": query.label,
            "conv": serialize_search_result(conv_result), "dino": serialize_search_result(dino_result),
            "conv_correct": conv_correct, "dino_correct": dino_correct, "assigned_group": assigned_group}


def flatten_query_result(result: Mapping[str, Any]) -> Dict[str, Any]:
    """export_utils.py:45-62."""
    conv, dino = result.get("conv", {}), result.get("dino", {})
    return {"query_image_path": result.get("query_image_path"), "query_label": result.get("query_label"),
            "group": result.get("assigned_group"), "conv_correct": result.get("conv_correct"),
            "dino_correct": result.get("dino_correct"),
            "conv_topk_image_paths": json.dumps(conv.get("image_paths", [])),
            "conv_topk_labels": json.dumps(conv.get("labels", [])),
            "conv_topk_scores": json.dumps(conv.get("scores", [])),
            "dino_topk_image_paths": json.dumps(dino.get("image_paths", [])),
            "dino_topk_labels": json.dumps(dino.get("labels", [])),
            "dino_topk_scores": json.dumps(dino.get("scores", []))}


def write_csv(path: str, rows: Sequence[Mapping[str, Any]]) -> str:
    """export_utils.py:22-42: header = union of the row keys in first-seen order (no rows: a file with an empty header
    line)."""
    out = Path(path)
    out.parent.mkdir(parents=True, exist_ok=True)
    rows = list(rows)
    fieldnames: List[str] = []
    for row in rows:
        for key in row.keys():
            if key not in fieldnames:
                fieldnames.append(key)
    with out.open("w", encoding="utf-8", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=fieldnames)
        w.writeheader()
        w.writerows(rows)
    return str(out)


def save_fused_embeddings(path: str, image_paths: Sequence[str], labels: Sequence[str], embeddings) -> str:
    """fusion_eval/evaluate.py:217-229."""
    np.savez_compressed(path, image_paths=np.asarray(image_paths), labels=np.asarray(labels),
                        embeddings=np.asarray(_np(embeddings), dtype=np.float32))
    return path


COMPARISON_GROUPS = ("both_correct", "both_wrong", "dino_correct_conv_wrong", "conv_correct_dino_wrong")


def export_analysis(payload: Mapping[str, Any], output_dir) -> List[str]:
    """Files of one dual-collection comparison (retrieval_analysis/run_analysis.py:67-85): the whole ``compare_models``
    payload as ``comparison_results.json``, its per-query rows flattened as ``comparison_results.csv`` and one
    ``group_<name>.csv`` per comparison group (written even when the group is empty).  -> the paths written."""
    root = Path(output_dir)
    written = [write_json(str(root / "comparison_results.json"), payload)]
    flat = [(row["assigned_group"], flatten_query_result(row)) for row in payload["results"]]
    written.append(write_csv(str(root / "comparison_results.csv"), [r for _, r in flat]))
    for group in COMPARISON_GROUPS:
        written.append(write_csv(str(root / f"group_{group}.csv"), [r for g, r in flat if g == group]))
    return written


def comparison_summary_text(payload: Mapping[str, Any]) -> str:
    """What the runner prints after a comparison (run_analysis.py:88-108), as one string."""
    cov, summ = payload["coverage"], payload["summary"]
    lines = ["Coverage:", f"  Present in ConvNeXt only: {len(cov['present_in_conv_only'])}",
             f"  Present in DINO only: {len(cov['present_in_dino_only'])}",
             f"  Present in both: {len(cov['present_in_both'])}", "Summary:"]
    lines += [f"  {key}: {summ[key]}" for key in COMPARISON_GROUPS + ("evaluated_queries",)]
    lines += [f"  missing_queries: {len(payload['missing_queries'])}", f"  errors: {len(payload['errors'])}"]
    return "\n".join(lines) + "\n"


def print_summary(payload: Mapping[str, Any]) -> None:
    print(comparison_summary_text(payload), end="")
