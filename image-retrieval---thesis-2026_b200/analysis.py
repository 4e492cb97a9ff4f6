"""Dual-collection retrieval analysis on local collections (SURVEY 8(a) D13 and the caller around it).

Mirrors the reference's correctness policy and four-way grouping:

* :class:`CorrectnessConfig`, :func:`is_retrieval_correct` -- retrieval_analysis/evaluator.py:12-26
* ``GROUP_*``, :func:`assign_group`                          -- retrieval_analysis/comparison.py:21-24, 236-244
* :class:`ComparisonConfig`, :func:`compare_models`          -- retrieval_analysis/comparison.py:27-38, 85-234

``compare_models`` is the caller on the far side of the hot path: it fetches the stored query embeddings of both
collections by image path, runs the batched exact search of each (:class:`LocalCollectionAdapter`, one fused GPU search
per batch instead of one network round trip), decides correctness and assigns the group.  The search itself is the
CUDA library; nothing here ranks or scores.
"""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

from .collection import LocalCollectionAdapter, QueryRecord, RetrievedItem
from .formats import build_query_analysis_row

GROUP_BOTH_CORRECT = "both_correct"
GROUP_BOTH_WRONG = "both_wrong"
GROUP_DINO_CORRECT_CONV_WRONG = "dino_correct_conv_wrong"
GROUP_CONV_CORRECT_DINO_WRONG = "conv_correct_dino_wrong"


@dataclass(frozen=True)
class CorrectnessConfig:
    """evaluator.py:12-15: how many of the first hits may carry the query's label."""

    top_k: int = 1


def is_retrieval_correct(query_label: Optional[str], results: Sequence[RetrievedItem],
                         config: CorrectnessConfig) -> bool:
    """evaluator.py:18-26: a query without a label (``None`` / empty) or without hits is never correct; otherwise any
    label equality among the first ``config.top_k`` hits."""
    if not query_label or not results:
        return False
    return any(item.label == query_label for item in results[: config.top_k])


def assign_group(conv_correct: bool, dino_correct: bool) -> str:
    """comparison.py:236-244."""
    if conv_correct and dino_correct:
        return GROUP_BOTH_CORRECT
    if not conv_correct and not dino_correct:
        return GROUP_BOTH_WRONG
    return GROUP_DINO_CORRECT_CONV_WRONG if dino_correct else GROUP_CONV_CORRECT_DINO_WRONG


class IdentityReranker:
    """retrieval_analysis/rerank.py:19-25."""

    def rerank(self, query, results):
        return list(results)


@dataclass
class ComparisonConfig:
    """comparison.py:27-38."""

    top_k: int = 5
    conv_search_params: Optional[Dict] = None
    dino_search_params: Optional[Dict] = None
    correctness: CorrectnessConfig = field(default_factory=CorrectnessConfig)
    skip_missing_queries: bool = True
    preload_batch_size: int = 100
    search_batch_size: int = 10


def compare_collection_coverage(conv_adapter, dino_adapter) -> Dict[str, List[str]]:
    """``image_path`` coverage of the two collections (retrieval_analysis/milvus_adapter.py:309-320)."""
    cp, dp = set(conv_adapter.list_image_paths()), set(dino_adapter.list_image_paths())
    return {"conv_only": sorted(cp - dp), "dino_only": sorted(dp - cp), "present_in_both": sorted(cp & dp)}


def filter_present_queries(queries, coverage: Dict[str, List[str]]) -> Dict[str, List[QueryRecord]]:
    """milvus_adapter.py:321-336."""
    present = set(coverage["present_in_both"])
    out: Dict[str, List[QueryRecord]] = {"valid": [], "missing": []}
    for q in queries:
        out["valid" if q.image_path in present else "missing"].append(q)
    return out


def compare_models(conv_adapter: LocalCollectionAdapter, dino_adapter: LocalCollectionAdapter,
                   queries: Sequence[QueryRecord], config: ComparisonConfig, reranker: Optional[Any] = None) -> Dict[str, object]:
    """comparison.py:85-234 with the same output dictionary (``coverage``, ``missing_queries``, ``errors``,
    ``summary``, ``results``)."""
    reranker = reranker or IdentityReranker()
    wanted = [q.image_path for q in queries if q.image_path]
    conv_rec = conv_adapter.fetch_records_by_image_paths(wanted, include_embedding=True,
                                                         batch_size=config.preload_batch_size)
    dino_rec = dino_adapter.fetch_records_by_image_paths(wanted, include_embedding=True,
                                                         batch_size=config.preload_batch_size)
    cp, dp = set(conv_rec), set(dino_rec)
    coverage = {"conv_only": sorted(cp - dp), "dino_only": sorted(dp - cp), "present_in_both": sorted(cp & dp)}
    part = filter_present_queries(queries, coverage)
    valid, missing = part["valid"], part["missing"]
    if missing and not config.skip_missing_queries:
        raise ValueError("Some query image_paths are not present in both collections: "
                         + ", ".join(q.image_path for q in missing[:5]))

    results: List[Dict[str, object]] = []
    summary: Counter = Counter()
    errors: List[Dict[str, str]] = []
    step = config.search_batch_size
    for start in range(0, len(valid), step):
        batch = valid[start:start + step]
        try:
            aligned, conv_emb, dino_emb = [], [], []
            for q in batch:
                c, d = conv_rec.get(q.image_path), dino_rec.get(q.image_path)
                if c is None or d is None:
                    errors.append({"query_image_path": q.image_path, "error": "missing_query_embedding_on_one_side"})
                    continue
                label = q.label or c.get(conv_adapter.config.label_field) or d.get(dino_adapter.config.label_field)
                aligned.append(QueryRecord(image_path=q.image_path, label=label))
                conv_emb.append(c[conv_adapter.config.vector_field])
                dino_emb.append(d[dino_adapter.config.vector_field])
            if not aligned:
                continue
            conv_res = conv_adapter.search_by_embeddings(aligned, conv_emb, config.top_k, config.conv_search_params,
                                                         reranker, True, None, step)
            dino_res = dino_adapter.search_by_embeddings(aligned, dino_emb, config.top_k, config.dino_search_params,
                                                         reranker, True, None, step)
            for q, cr, dr in zip(aligned, conv_res, dino_res):
                cc = is_retrieval_correct(q.label, cr.retrieved, config.correctness)
                dc = is_retrieval_correct(q.label, dr.retrieved, config.correctness)
                group = assign_group(cc, dc)
                summary[group] += 1
                results.append(build_query_analysis_row(q, cr, dr, cc, dc, group))
        except Exception as exc:  # noqa: BLE001 - the reference records the failure per query and goes on
            for q in batch:
                errors.append({"query_image_path": q.image_path, "error": str(exc)})

    return {
        "coverage": {"present_in_conv_only": coverage["conv_only"], "present_in_dino_only": coverage["dino_only"],
                     "present_in_both": coverage["present_in_both"]},
        "missing_queries": [{"image_path": q.image_path, "label": q.label} for q in missing],
        "errors": errors,
        "summary": {g: summary[g] for g in (GROUP_BOTH_CORRECT, GROUP_BOTH_WRONG, GROUP_DINO_CORRECT_CONV_WRONG,
                                            GROUP_CONV_CORRECT_DINO_WRONG)} | {"evaluated_queries": len(results)},
        "results": results,
    }
