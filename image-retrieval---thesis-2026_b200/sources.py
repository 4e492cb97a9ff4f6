"""The INPUT side of the late-fusion path: where the two embedding sets come from and how they are joined before
``fusion.run_late_fusion_experiments`` searches them -- and the shaping of its results for the console / JSON / CSV.

Host-side data formats only; no device work happens here.  Drop-ins, same names and behaviour:

  * ``EmbeddingRecord``, ``EmbeddingSource``, ``FileEmbeddingSource`` (``.json`` with a ``records`` list, ``.npz`` with
    ``image_paths`` / ``labels`` / ``embeddings``), ``build_embedding_source``, ``align_embedding_sources``
    -- fusion_eval/align.py:16-229;
  * ``CollectionEmbeddingSource`` -- ``MilvusEmbeddingSource`` (align.py:44-93) over a local collection adapter
    (``collection.LocalCollectionAdapter``) instead of a Milvus connection;
  * ``load_query_set`` -- retrieval_analysis/comparison.py:41-83 (JSON / CSV / whitespace text);
  * ``experiment_rows``, ``format_results_table``, ``late_fusion_payload`` -- fusion_eval/run_late_fusion.py:55-121,
    188-203.

Pinned against the real reference functions by tests/test_sources.py (golden vectors: tests/golden/golden_sources.json).
"""
from __future__ import annotations

import csv
import json
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np

from .collection import QueryRecord
from .fusion import AlignedEmbeddings

RESULT_COLUMNS = ("experiment", "samples", "mP@1", "mP@5", "mP@10", "R@1", "R@5", "R@10", "mAP", "status")
_METRIC_COLUMNS = RESULT_COLUMNS[2:-1]


@dataclass(frozen=True)
class EmbeddingRecord:
    """One stored embedding (align.py:16-24)."""

    image_path: str
    label: Optional[str]
    embedding: np.ndarray
    source_name: str
    raw: Mapping[str, Any]


class EmbeddingSource:
    """Anything with ``fetch_all() -> List[EmbeddingRecord]`` (align.py:38-42)."""

    def fetch_all(self) -> List[EmbeddingRecord]:
        raise NotImplementedError


def _record(path, label, vector, source_name: str, raw: Mapping[str, Any]) -> EmbeddingRecord:
    return EmbeddingRecord(image_path=path, label=label, embedding=np.asarray(vector, dtype=np.float32),
                           source_name=source_name, raw=raw)


class FileEmbeddingSource(EmbeddingSource):
    """Embeddings from a local ``.json`` (``{"records": [{image_path, label?, embedding, ...}]}``) or ``.npz``
    (``image_paths``, ``embeddings``, optional ``labels``) file -- align.py:96-143."""

    def __init__(self, path, source_name: str):
        self.path = Path(path)
        self.source_name = source_name

    def fetch_all(self) -> List[EmbeddingRecord]:
        kind = self.path.suffix.lower()
        if kind == ".json":
            with self.path.open("r", encoding="utf-8") as fh:
                payload = json.load(fh)
            # the reference asks the payload for its "records" (a bare list has no .get: AttributeError, as there)
            rows = payload.get("records", payload)
            return [_record(row["image_path"], row.get("label"), row["embedding"], self.source_name, row) for row in rows]
        if kind == ".npz":
            z = np.load(self.path, allow_pickle=True)
            paths = z["image_paths"].tolist()
            labels = z["labels"].tolist() if "labels" in z else [None] * len(paths)
            return [_record(p, lab, vec, self.source_name, {}) for p, lab, vec in zip(paths, labels, z["embeddings"])]
        raise ValueError(f"Unsupported embedding file format: {self.path}")


class CollectionEmbeddingSource(EmbeddingSource):
    """Every row of one collection, paged through ``adapter.client.query`` with the ``image_path != ""`` filter the
    reference's ``MilvusEmbeddingSource`` uses (align.py:53-93).  ``adapter`` is a ``LocalCollectionAdapter`` (or anything
    with ``.config`` and ``.client.query(filter=, output_fields=, limit=, offset=)``)."""

    def __init__(self, adapter, batch_size: int = 1000):
        self.adapter = adapter
        self.config = adapter.config
        self.batch_size = int(batch_size)

    def fetch_all(self) -> List[EmbeddingRecord]:
        cfg = self.config
        fields = list(cfg.output_fields)
        for needed in (cfg.image_path_field, cfg.label_field, cfg.vector_field):
            if needed not in fields:
                fields.append(needed)
        flt = f'{cfg.image_path_field} != ""'
        rows: List[Mapping[str, Any]] = []
        while True:
            page = self.adapter.client.query(filter=flt, output_fields=fields, limit=self.batch_size, offset=len(rows))
            if not page:
                break
            rows.extend(page)
        return [_record(r[cfg.image_path_field], r.get(cfg.label_field), r[cfg.vector_field], cfg.name, r)
                for r in rows if r.get(cfg.image_path_field) is not None]


def build_embedding_source(config: Mapping[str, Any]) -> EmbeddingSource:
    """align.py:146-160.  ``{"type": "file", "path": ..., "name": ...}`` as in the reference; the collection-backed kind
    (the reference's default type, "milvus") takes the adapter object itself: ``{"type": "milvus", "adapter":
    LocalCollectionAdapter(...), "batch_size": 1000}`` -- a local collection has no connection settings to build from."""
    kind = config.get("type", "milvus")
    if kind == "file":
        return FileEmbeddingSource(path=config["path"], source_name=config["name"])
    if kind in ("milvus", "collection"):
        if config.get("adapter") is None:
            raise ValueError("a collection-backed embedding source needs config['adapter'] (a LocalCollectionAdapter)")
        return CollectionEmbeddingSource(config["adapter"], batch_size=config.get("batch_size", 1000))
    raise ValueError(f"Unsupported source type: {kind}")


def _by_path(records: Iterable[EmbeddingRecord], who: str) -> Dict[str, EmbeddingRecord]:
    table: Dict[str, EmbeddingRecord] = {}
    for rec in records:
        if rec.image_path in table:
            raise ValueError(f"Duplicate image_path found in {who}: {rec.image_path}")
        table[rec.image_path] = rec
    return table


def align_embedding_sources(conv_source: EmbeddingSource, dino_source: EmbeddingSource, query_set_path=None,
                            strict_label_check: bool = True) -> AlignedEmbeddings:
    """Join the two stores by ``image_path`` (align.py:163-217): the images both hold, in sorted order -- or, with a query
    set, its images in ITS order -- with the coverage lists; a label that differs between the stores is an error unless
    ``strict_label_check=False`` (the ConvNeXt label then wins, "unknown" when neither store has one)."""
    conv = _by_path(conv_source.fetch_all(), "ConvNeXt")
    dino = _by_path(dino_source.fetch_all(), "DINO")
    both = sorted(conv.keys() & dino.keys())
    coverage = {"present_in_conv_only": sorted(conv.keys() - dino.keys()),
                "present_in_dino_only": sorted(dino.keys() - conv.keys()),
                "present_in_both": both}
    if query_set_path:
        wanted = [q.image_path for q in load_query_set(query_set_path) if q.image_path in conv and q.image_path in dino]
    else:
        wanted = both
    labels: List[str] = []
    for path in wanted:
        a, b = conv[path].label, dino[path].label
        if strict_label_check and a != b:
            raise ValueError(f"Label mismatch for image_path={path}: conv={a!r}, dino={b!r}")
        labels.append(a or b or "unknown")
    if not wanted:
        raise ValueError("No aligned samples found across the requested sources")
    return AlignedEmbeddings(image_paths=list(wanted), labels=labels,
                             conv_embeddings=np.stack([conv[p].embedding for p in wanted]).astype(np.float32),
                             dino_embeddings=np.stack([dino[p].embedding for p in wanted]).astype(np.float32),
                             coverage=coverage)


def load_query_set(path) -> List[QueryRecord]:
    """An ORDERED query set (comparison.py:41-83).  ``.json``: a list, or an object with ``queries`` / ``results``; items
    carry ``image_path`` (or ``query_image_path``) and optionally ``label``.  ``.csv``: columns ``image_path`` /
    ``query_image_path`` and ``label`` / ``query_label``.  Anything else: one ``path [label]`` per line, ``#`` comments.
    Entries without a path are dropped."""
    file = Path(path)
    kind = file.suffix.lower()
    if kind == ".json":
        with file.open("r", encoding="utf-8") as fh:
            items = json.load(fh)
        if isinstance(items, dict):
            items = items.get("queries", items.get("results", []))
        out = []
        for item in items:
            where = item.get("image_path", item.get("query_image_path"))   # a present-but-empty image_path hides the other
            if where:
                out.append(QueryRecord(image_path=where, label=item.get("label")))
        return out
    if kind == ".csv":
        with file.open("r", encoding="utf-8", newline="") as fh:
            return [QueryRecord(image_path=row.get("image_path", row.get("query_image_path", "")),
                                label=row.get("label", row.get("query_label")))
                    for row in csv.DictReader(fh) if row.get("image_path") or row.get("query_image_path")]
    out = []
    with file.open("r", encoding="utf-8") as fh:
        for line in fh:
            words = line.split()
            if not words or words[0].startswith("#"):
                continue
            out.append(QueryRecord(image_path=words[0], label=words[1] if len(words) > 1 else None))
    return out


# ----------------------------------------------------------------------------------------------------------------
# results of run_late_fusion_experiments -> rows / table / payload (run_late_fusion.py:55-121, 188-203)
# ----------------------------------------------------------------------------------------------------------------
def experiment_rows(experiments) -> List[Dict[str, Any]]:
    """One display row per experiment: metrics as ``%.2f`` strings (0.00 for a metric the experiment did not report),
    empty cells and the reason in ``status`` for a skipped one."""
    rows = []
    for exp in experiments:
        row: Dict[str, Any] = {"experiment": exp.experiment_name, "samples": exp.num_samples}
        for col in _METRIC_COLUMNS:
            row[col] = "" if exp.skipped else f"{exp.metrics.get(col, 0.0):.2f}"
        row["status"] = (exp.skipped_reason or "skipped") if exp.skipped else "ok"
        rows.append(row)
    return rows


def format_results_table(rows: Sequence[Mapping[str, Any]]) -> str:
    """The console table of the runner: `` | ``-separated, left-justified columns, a ``-+-`` rule under the header."""
    width = {c: max([len(c)] + [len(str(r.get(c, ""))) for r in rows]) for c in RESULT_COLUMNS}
    line = lambda cells: " | ".join(str(cells.get(c, "")).ljust(width[c]) for c in RESULT_COLUMNS)  # noqa: E731
    rule = "-+-".join("-" * width[c] for c in RESULT_COLUMNS)
    return "\n".join([line({c: c for c in RESULT_COLUMNS}), rule] + [line(r) for r in rows])


def late_fusion_payload(aligned: AlignedEmbeddings, experiments) -> Dict[str, Any]:
    """``late_fusion_results.json`` of the runner (run_late_fusion.py:188-203)."""
    return {"coverage": aligned.coverage, "num_evaluated_samples": len(aligned.image_paths),
            "results": [{"experiment_name": e.experiment_name, "num_samples": e.num_samples, "metrics": e.metrics,
                         "skipped": e.skipped, "skipped_reason": e.skipped_reason} for e in experiments]}
