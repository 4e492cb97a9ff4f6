"""Local, device-resident drop-in for the Milvus / Zilliz collection classes of the reference.

The reference searches remote collections (``collection.search`` of pymilvus, ``MilvusClient.search``) one query at a
time over the network.  Here the gallery stays in HBM (a :class:`FlatIndex`, exact FLAT search) and the metadata
(``image_path``, ``label``, ``label_vector_json`` ...) in host-side columns keyed by the row id, behind the same call
shapes:

* :class:`LocalCollection`         -- ``insert`` / ``search`` / ``query`` with pymilvus-shaped results
  (``collection.search(data, anns_field, param, limit, output_fields)``, milvus/milvus_retrieval.py:80-86,
  nih_zilliz_utils.py:260-266; ``client.search`` / ``client.query``, retrieval_analysis/milvus_adapter.py:94-200)
* :class:`LocalCollectionAdapter`  -- ``MilvusCollectionAdapter`` (retrieval_analysis/milvus_adapter.py:63-306)
* :class:`LocalRetriever`          -- ``MilvusRetriever`` (milvus/milvus_retrieval.py:15-140)
* :func:`insert_rows`, :func:`search_collection` -- nih_zilliz_utils.py:244-280

Metric conventions follow the reference's observed behaviour: for COSINE and IP the hit ``distance`` IS
the similarity (larger = closer); for L2 it is the Euclidean distance and ``similarity = 1 - d^2/2``.
"""
from __future__ import annotations

import json
import os
import re
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np
import torch

from .search import FlatIndex

_METRIC_OF = {"COSINE": "cosine", "IP": "ip", "L2": "l2"}


class Hit(dict):
    """One pymilvus-like hit: ``hit.id``, ``hit.distance``, ``hit.entity.get(field)`` and the dict form
    ``{"id", "distance", "entity": {...}}`` that ``MilvusClient.search`` returns."""

    @property
    def id(self):
        return self["id"]

    @property
    def distance(self):
        return self["distance"]

    @property
    def entity(self):
        return self["entity"]


class LocalCollection:
    """An exact (FLAT) collection: embeddings in HBM, scalar fields on the host."""

    def __init__(self, name: str, dim: int, metric_type: str = "COSINE", precision: str = "fp32",
                 vector_field: str = "embedding", id_field: str = "id", device=None, l2_squared: bool = False):
        if metric_type not in _METRIC_OF:
            raise ValueError(f"metric_type must be one of {sorted(_METRIC_OF)}")
        self.name, self.dim, self.metric_type = name, int(dim), metric_type
        self.vector_field, self.id_field = vector_field, id_field
        self._precision, self._device = precision, device
        # NOTE on L2: ``hit.distance`` is the EUCLIDEAN distance -- what the reference's own mapping
        # ``similarity = 1 - d^2 / 2`` (milvus/milvus_retrieval.py:102-107) assumes and what its in-process search ranks
        # by (torch.cdist, test_ath.py:87).  A real Milvus L2 collection and faiss.IndexFlatL2 return the SQUARED
        # distance (same ranking): ``l2_squared=True`` reports that instead.
        self.l2_squared = bool(l2_squared)
        self._index: Optional[FlatIndex] = None    # created by the first search (metadata-only use needs no device)
        self.columns: Dict[str, List[Any]] = {}
        self._raw_rows: List[torch.Tensor] = []   # vectors as inserted (host)
        self._uploaded = 0                         # chunks of _raw_rows already appended to the device index
        self._count = 0

    @property
    def index(self) -> FlatIndex:
        if self._index is None:
            # COSINE collections normalise on insert and on query, as Milvus does for the COSINE metric
            self._index = FlatIndex(self.dim, _METRIC_OF[self.metric_type], self._precision,
                                    normalize=(self.metric_type == "COSINE"), device=self._device)
        return self._index

    # ------------------------------------------------------------------ ingest
    @property
    def num_entities(self) -> int:
        return self._count

    def insert(self, rows: Sequence[Mapping[str, Any]]) -> List[int]:
        """``client.insert(collection_name, rows)`` / ``collection.insert`` (ingest_embeddings.py:386-414): every row
        is a mapping holding the vector field plus scalar fields.  -> the assigned ids."""
        if not rows:
            return []
        start = self.num_entities
        vec = torch.as_tensor(np.asarray([np.asarray(r[self.vector_field], dtype=np.float32) for r in rows]))
        if vec.ndim != 2 or vec.shape[1] != self.dim:
            raise ValueError(f"{self.name}: expected vectors of dim {self.dim}, got {tuple(vec.shape)}")
        self._raw_rows.append(vec)   # uploaded to HBM by the next search (metadata-only use needs no device)
        self._count += len(rows)
        keys = set(self.columns) | {k for r in rows for k in r if k != self.vector_field} | {self.id_field}
        ids = list(range(start, start + len(rows)))
        for key in keys:
            col = self.columns.setdefault(key, [None] * start)
            col.extend(r.get(key) for r in rows)
        idcol = self.columns[self.id_field]
        for j, i in enumerate(ids):      # auto id = row number unless the row brought its own
            if idcol[start + j] is None:
                idcol[start + j] = i
        return ids

    def flush(self) -> None:
        """pymilvus API compatibility (everything is resident already)."""

    def load(self) -> None:
        """pymilvus API compatibility."""

    # ------------------------------------------------------------------ search
    def _entity(self, row: int, output_fields: Optional[Sequence[str]]) -> Dict[str, Any]:
        fields = list(output_fields) if output_fields else [k for k in self.columns]
        ent = {}
        for f in fields:
            if f == self.vector_field:
                ent[f] = self.vector(row).tolist()
            elif f in self.columns:
                ent[f] = self.columns[f][row]
        return ent

    def vector(self, row: int) -> np.ndarray:
        """The stored (un-normalised, as inserted) vector of a row."""
        for chunk in self._raw_rows:
            if row < chunk.shape[0]:
                return chunk[row].numpy()
            row -= chunk.shape[0]
        raise IndexError(row)

    def search(self, data, anns_field: Optional[str] = None, param: Optional[Mapping[str, Any]] = None,
               limit: int = 10, output_fields: Optional[Sequence[str]] = None,
               search_params: Optional[Mapping[str, Any]] = None, collection_name: Optional[str] = None,
               **_ignored) -> List[List[Hit]]:
        """``collection.search(data=[vec, ...], anns_field, param, limit, output_fields)`` and
        ``client.search(collection_name, data, anns_field, search_params, limit, output_fields)``: one list of hits per
        query, best first.  ``nprobe`` and other ANN parameters are accepted and ignored -- the search is exact."""
        if anns_field not in (None, self.vector_field):
            raise ValueError(f"{self.name}: unknown vector field {anns_field!r}")
        p = dict(param or search_params or {})
        mt = p.get("metric_type", self.metric_type)
        if mt != self.metric_type:
            raise ValueError(f"{self.name}: collection metric is {self.metric_type}, search asked for {mt}")
        if self.num_entities == 0:
            return [[] for _ in data]
        while self._uploaded < len(self._raw_rows):
            self.index.add(self._raw_rows[self._uploaded].to(self.index.device))
            self._uploaded += 1
        q = torch.as_tensor(np.asarray(data, dtype=np.float32))
        if q.ndim == 1:
            q = q[None]
        k = max(1, min(int(limit), self.num_entities))
        vals, idx = self.index.search(q.to(self.index.device), k)
        vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
        if self.metric_type == "L2" and self.l2_squared:
            vals = vals * vals
        out: List[List[Hit]] = []
        idcol = self.columns.get(self.id_field)
        for r in range(idx.shape[0]):
            hits = []
            for j in range(idx.shape[1]):
                row = int(idx[r, j])
                if row < 0:
                    continue
                hits.append(Hit(id=idcol[row] if idcol else row, distance=float(vals[r, j]),
                                score=float(vals[r, j]), entity=self._entity(row, output_fields)))
            out.append(hits)
        return out

    # ------------------------------------------------------------------ scalar queries
    _EQ = re.compile(r'^\s*(\w+)\s*(==|!=)\s*"((?:[^"\\]|\\.)*)"\s*$')
    _IN = re.compile(r'^\s*(\w+)\s+in\s+\[(.*)\]\s*$', re.S)

    @staticmethod
    def _unescape(s: str) -> str:
        return s.replace('\\"', '"').replace("\\\\", "\\")

    def query(self, filter: str = "", output_fields: Optional[Sequence[str]] = None, limit: Optional[int] = None,
              offset: int = 0, collection_name: Optional[str] = None, expr: Optional[str] = None,
              **_ignored) -> List[Dict[str, Any]]:
        """``client.query(collection_name, filter, output_fields, limit, offset)`` for the three expression forms the
        reference builds: ``field == "v"``, ``field != "v"`` and ``field in ["a", "b"]``
        (milvus_adapter.py:94-175, ``_eq_expr`` / ``_in_expr``)."""
        flt = filter or expr or ""
        n = self.num_entities
        rows: Iterable[int]
        m = self._EQ.match(flt)
        if m:
            col, op, val = m.group(1), m.group(2), self._unescape(m.group(3))
            data = self.columns.get(col, [None] * n)
            rows = [i for i in range(n) if (data[i] == val) == (op == "==") and (op == "==" or data[i] is not None)]
        else:
            m = self._IN.match(flt)
            if m:
                col = m.group(1)
                vals = {self._unescape(v) for v in re.findall(r'"((?:[^"\\]|\\.)*)"', m.group(2))}
                data = self.columns.get(col, [None] * n)
                rows = [i for i in range(n) if data[i] in vals]
            elif flt.strip() == "":
                rows = range(n)
            else:
                raise ValueError(f"{self.name}: unsupported filter expression {flt!r}")
        rows = list(rows)[int(offset):]
        if limit is not None:
            rows = rows[: int(limit)]
        return [self._entity(i, output_fields) for i in rows]


# ----------------------------------------------------------------------------------------------------------------
# retrieval_analysis/milvus_adapter.py
# ----------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class CollectionConfig:
    """milvus_adapter.py:12-30 without the connection settings."""

    name: str
    collection_name: str
    vector_field: str = "embedding"
    id_field: str = "id"
    image_path_field: str = "image_path"
    label_field: str = "label"
    output_fields: Sequence[str] = field(default_factory=lambda: ("id", "image_path", "label"))


@dataclass
class QueryRecord:
    """milvus_adapter.py:33-38."""

    image_path: str
    label: Optional[str] = None


@dataclass
class RetrievedItem:
    """milvus_adapter.py:41-50."""

    id: Optional[Any]
    image_path: Optional[str]
    label: Optional[str]
    score: Optional[float]
    distance: Optional[float]
    raw: Dict[str, Any] = field(default_factory=dict)


@dataclass
class SearchResult:
    """milvus_adapter.py:53-60."""

    query: QueryRecord
    query_source: str
    retrieved: List[RetrievedItem]
    query_embedding: Sequence[float]


class LocalCollectionAdapter:
    """``MilvusCollectionAdapter`` over a :class:`LocalCollection`: same methods, same return types."""

    def __init__(self, config: CollectionConfig, collection: LocalCollection):
        self.config = config
        self.client = collection

    def list_image_paths(self, batch_size: int = 1000) -> List[str]:
        """milvus_adapter.py:91-116 (paged ``image_path != ""`` query)."""
        out: List[str] = []
        offset = 0
        while True:
            rows = self.client.query(filter=f'{self.config.image_path_field} != ""',
                                     output_fields=[self.config.image_path_field], limit=batch_size, offset=offset)
            if not rows:
                break
            out.extend(r.get(self.config.image_path_field) for r in rows)
            offset += len(rows)
        return [p for p in out if p]

    def _fields(self, include_embedding: bool) -> List[str]:
        fields = list(self.config.output_fields)
        if include_embedding and self.config.vector_field not in fields:
            fields.append(self.config.vector_field)
        return fields

    def fetch_record_by_image_path(self, image_path: str, include_embedding: bool = True) -> Optional[Dict[str, Any]]:
        """milvus_adapter.py:118-137."""
        rows = self.client.query(filter=_eq_expr(self.config.image_path_field, image_path),
                                 output_fields=self._fields(include_embedding), limit=2)
        if not rows:
            return None
        if len(rows) > 1:
            raise ValueError(f"{self.config.name}: multiple rows found for image_path={image_path}")
        return rows[0]

    def fetch_records_by_image_paths(self, image_paths: Sequence[str], include_embedding: bool = True,
                                     batch_size: int = 100) -> Dict[str, Dict[str, Any]]:
        """milvus_adapter.py:139-175."""
        indexed: Dict[str, Dict[str, Any]] = {}
        for start in range(0, len(image_paths), batch_size):
            chunk = [p for p in image_paths[start:start + batch_size] if p]
            if not chunk:
                continue
            for row in self.client.query(filter=_in_expr(self.config.image_path_field, chunk),
                                         output_fields=self._fields(include_embedding), limit=len(chunk)):
                p = row.get(self.config.image_path_field)
                if p is None:
                    continue
                if p in indexed:
                    raise ValueError(f"{self.config.name}: multiple rows found for image_path={p}")
                indexed[p] = row
        return indexed

    def search_by_embedding(self, query: QueryRecord, query_embedding: Sequence[float], top_k: int,
                            search_params: Optional[Dict[str, Any]] = None, reranker: Optional[Any] = None,
                            exclude_self: bool = True, metadata_fields: Optional[Sequence[str]] = None) -> SearchResult:
        """milvus_adapter.py:177-216."""
        return self.search_by_embeddings([query], [query_embedding], top_k, search_params, reranker, exclude_self,
                                         metadata_fields)[0]

    def search_by_embeddings(self, queries: Sequence[QueryRecord], query_embeddings: Sequence[Sequence[float]],
                             top_k: int, search_params: Optional[Dict[str, Any]] = None, reranker: Optional[Any] = None,
                             exclude_self: bool = True, metadata_fields: Optional[Sequence[str]] = None,
                             batch_size: Optional[int] = None) -> List[SearchResult]:
        """milvus_adapter.py:218-275: ``limit = top_k + 1`` when the query itself must be dropped by image path, then
        the optional reranker, then the cut to ``top_k``.  All queries of a batch go through ONE fused search."""
        if not queries:
            return []
        if len(queries) != len(query_embeddings):
            raise ValueError("queries and query_embeddings must have the same length")
        fields = list(metadata_fields or self.config.output_fields)
        for f in (self.config.image_path_field, self.config.label_field):
            if f not in fields:
                fields.append(f)
        results: List[SearchResult] = []
        step = max(1, int(batch_size or len(queries)))
        for start in range(0, len(queries), step):
            bq, be = queries[start:start + step], query_embeddings[start:start + step]
            raw = self.client.search(data=[list(e) for e in be], anns_field=self.config.vector_field,
                                     search_params=search_params or {}, limit=top_k + 1 if exclude_self else top_k,
                                     output_fields=fields)
            for query, emb, hits in zip(bq, be, raw or []):
                retrieved = [self._normalize_hit(h) for h in hits]
                if exclude_self:
                    retrieved = [it for it in retrieved if it.image_path != query.image_path]
                if reranker is not None:
                    retrieved = list(reranker.rerank(query=query, results=retrieved))
                results.append(SearchResult(query, self.config.name, retrieved[:top_k], emb))
        return results

    def _normalize_hit(self, hit: Mapping[str, Any]) -> RetrievedItem:
        """milvus_adapter.py:277-289."""
        entity = hit.get("entity", {})
        score = hit.get("score")
        distance = hit.get("distance", score)
        return RetrievedItem(id=entity.get(self.config.id_field, hit.get("id")),
                             image_path=entity.get(self.config.image_path_field),
                             label=entity.get(self.config.label_field), score=score, distance=distance, raw=dict(hit))


def _escape(value: str) -> str:
    return value.replace("\\", "\\\\").replace('"', '\\"')


def _eq_expr(field_name: str, value: str) -> str:
    """milvus_adapter.py:291-294."""
    return f'{field_name} == "{_escape(value)}"'


def _in_expr(field_name: str, values: Sequence[str]) -> str:
    """milvus_adapter.py:296-302."""
    return f"{field_name} in [{', '.join(chr(34) + _escape(v) + chr(34) for v in values)}]"


# ----------------------------------------------------------------------------------------------------------------
# milvus/milvus_retrieval.py
# ----------------------------------------------------------------------------------------------------------------
def similarity_from_distance(distance: float, metric_type: str) -> Optional[float]:
    """milvus_retrieval.py:92-107 as implemented: COSINE / IP pass the value through, L2 -> 1 - d^2/2."""
    if metric_type in ("COSINE", "IP"):
        return distance
    if metric_type == "L2":
        return 1.0 - (distance * distance) / 2.0
    return None


class LocalRetriever:
    """``MilvusRetriever``: ``search(query, top_k, search_params, metric_type) -> (results, query_embedding)`` where a
    result is ``{"id", "image_path", "label", "distance", "similarity"}``.  The image -> embedding step of the
    reference (PIL + transform + model) is the caller's ``embed_fn``; a tensor / array query is taken as the embedding
    and L2-normalised like the reference does (milvus_retrieval.py:63)."""

    def __init__(self, collection: LocalCollection, embed_fn: Optional[Callable[[Any], torch.Tensor]] = None):
        self.collection = collection
        self.embed_fn = embed_fn

    def load_collection(self) -> LocalCollection:
        return self.collection

    def _embed(self, query) -> torch.Tensor:
        from .search import normalize

        if isinstance(query, (torch.Tensor, np.ndarray, list, tuple)):
            e = torch.as_tensor(np.asarray(query, dtype=np.float32) if not isinstance(query, torch.Tensor) else query)
        elif self.embed_fn is not None:
            e = self.embed_fn(query)
        else:
            raise ValueError("an embed_fn is needed to search by image path")
        e = e.float().reshape(1, -1)
        return normalize(e.cuda() if not e.is_cuda else e)

    def search(self, query_image_path, top_k: int = 10, search_params: Optional[Mapping[str, Any]] = None,
               metric_type: str = "COSINE"):
        query_embedding = self._embed(query_image_path)
        if search_params is None:
            search_params = {"metric_type": metric_type, "params": {"nprobe": 10}}
        hits = self.collection.search(data=[query_embedding[0].cpu().numpy()], anns_field=self.collection.vector_field,
                                      param=search_params, limit=top_k, output_fields=["image_path", "label"])
        results = [{"id": h.id, "image_path": h.entity.get("image_path"), "label": h.entity.get("label"),
                    "distance": h.distance, "similarity": similarity_from_distance(h.distance, metric_type)}
                   for h in hits[0]]
        return results, query_embedding

    def batch_search(self, query_image_paths, top_k: int = 10, search_params=None):
        """milvus_retrieval.py:122-140."""
        return [self.search(q, top_k, search_params)[0] for q in query_image_paths]


class PathMapper:
    """Stored (Kaggle) image paths -> paths on the local machine (milvus/path_mapper.py:10-107): a hit's ``image_path`` is
    re-rooted by FILE NAME under ``local_base_path``."""

    def __init__(self, kaggle_prefix: str = "/kaggle/input", local_base_path: Optional[str] = None):
        self.kaggle_prefix = kaggle_prefix
        self.local_base_path = local_base_path

    def extract_filename(self, kaggle_path: str) -> str:
        return os.path.basename(kaggle_path)

    def extract_relative_path(self, kaggle_path: str) -> str:
        """What follows ``.../input/<dataset>/``; the file name when the path has no ``input`` component."""
        parts = kaggle_path.split("/")
        if "input" in parts:
            return "/".join(parts[parts.index("input") + 2:])
        return self.extract_filename(kaggle_path)

    def remap_path(self, kaggle_path: str, local_base_path: Optional[str] = None) -> str:
        base = local_base_path or self.local_base_path
        if not base:
            raise ValueError("local_base_path must be provided")
        return os.path.join(base, self.extract_filename(kaggle_path))

    def verify_path(self, kaggle_path: str, local_base_path: Optional[str] = None):
        remapped = self.remap_path(kaggle_path, local_base_path)
        return os.path.exists(remapped), remapped

    def batch_remap(self, kaggle_paths, local_base_path: Optional[str] = None) -> List[str]:
        return [self.remap_path(p, local_base_path) for p in kaggle_paths]


class LocalRetrieverPatched(LocalRetriever):
    """``MilvusRetrieverPatched`` (milvus/milvus_retrieval_patched.py:9-133): the same search, with the ``image_path`` of
    every hit that starts with ``/kaggle/`` re-rooted under ``local_data_base_path`` (the reference also prints each
    remapped pair; this one does not)."""

    def __init__(self, collection: LocalCollection, embed_fn: Optional[Callable[[Any], torch.Tensor]] = None,
                 local_data_base_path: Optional[str] = None, enable_path_mapping: bool = True):
        super().__init__(collection, embed_fn)
        self.enable_path_mapping = enable_path_mapping
        self.path_mapper = PathMapper(local_base_path=local_data_base_path) \
            if (enable_path_mapping and local_data_base_path) else None

    def _remap_image_path(self, kaggle_path: str) -> str:
        if self.path_mapper is None or not kaggle_path.startswith("/kaggle/"):
            return kaggle_path
        return self.path_mapper.remap_path(kaggle_path)

    def _remap_results(self, results: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
        for r in results:
            r["image_path"] = self._remap_image_path(r["image_path"])
        return results

    def search(self, query_image_path, top_k: int = 10, search_params: Optional[Mapping[str, Any]] = None,
               metric_type: str = "COSINE"):
        results, query_embedding = super().search(query_image_path, top_k, search_params, metric_type)
        return self._remap_results(results), query_embedding


# ----------------------------------------------------------------------------------------------------------------
# nih_zilliz_utils.py
# ----------------------------------------------------------------------------------------------------------------
def insert_rows(collection: LocalCollection, rows: Sequence[Mapping[str, Any]]) -> None:
    """nih_zilliz_utils.py:244-251: rows hold ``image_path, image_name, label_names, multi_hot, embedding``; the
    collection stores ``label_text = "|".join(label_names)`` and ``label_vector_json = json.dumps(multi_hot)``."""
    collection.insert([{"image_path": r["image_path"], "image_name": r["image_name"],
                        "label_text": "|".join(r["label_names"]), "label_vector_json": json.dumps(r["multi_hot"]),
                        collection.vector_field: np.asarray(r["embedding"], dtype=np.float32)} for r in rows])
    collection.flush()


def search_collection(collection: LocalCollection, query_vector, top_k: int, nprobe: int = 10) -> List[Dict[str, Any]]:
    """nih_zilliz_utils.py:254-280 (one query) -- and ``search_collection_batch`` for many at once."""
    return search_collection_batch(collection, [query_vector], top_k, nprobe)[0]


def search_collection_batch(collection: LocalCollection, query_vectors, top_k: int, nprobe: int = 10):
    res = collection.search(data=query_vectors, anns_field=collection.vector_field,
                            param={"metric_type": "COSINE", "params": {"nprobe": nprobe}}, limit=top_k,
                            output_fields=["image_path", "image_name", "label_text", "label_vector_json"])
    return [[{"id": h.id, "score": float(h.distance), "image_path": h.entity.get("image_path"),
              "image_name": h.entity.get("image_name"), "label_text": h.entity.get("label_text"),
              "label_vector": json.loads(h.entity.get("label_vector_json"))} for h in hits] for hits in res]
