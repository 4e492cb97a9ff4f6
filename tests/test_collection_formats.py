"""Collection drop-in and wire formats (8(f)-2).

CPU part: metadata store, filter expressions, result containers and every file format, against the schemas the
reference writes / reads (no device needed: vectors are uploaded lazily by the first search).
GPU part: search results of the local classes against the golden outputs of the REAL reference classes
(oracle/make_golden_collection.py)."""
import csv
import json
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_collection as mgc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gc():
    with open(os.path.join(GOLDEN, "golden_collection.json")) as fh:
        return json.load(fh)


def _collection(precision="fp32"):
    from b200knn.collection import LocalCollection

    x, paths, labels = mgc.inputs()
    coll = LocalCollection("conv", x.shape[1], "COSINE", precision)
    coll.insert([{"image_path": p, "label": l, "embedding": v} for p, l, v in zip(paths, labels, x)])
    return coll, x, paths, labels


# ------------------------------------------------------------------------------------------------ CPU
def test_metadata_queries_follow_the_reference_filter_forms():
    from b200knn.collection import CollectionConfig, LocalCollectionAdapter, _eq_expr, _in_expr

    coll, x, paths, labels = _collection()
    assert coll.num_entities == len(paths)
    ad = LocalCollectionAdapter(CollectionConfig("conv", "c"), coll)
    assert ad.list_image_paths(batch_size=64) == paths
    rec = ad.fetch_record_by_image_path(paths[17])
    assert rec["image_path"] == paths[17] and rec["label"] == labels[17] and rec["id"] == 17
    assert np.allclose(rec["embedding"], x[17])
    assert ad.fetch_record_by_image_path("nope.png") is None
    many = ad.fetch_records_by_image_paths([paths[3], paths[250], "", "missing"], include_embedding=False, batch_size=2)
    assert sorted(many) == sorted([paths[3], paths[250]]) and "embedding" not in many[paths[3]]
    weird = 'a"b\\c.png'
    coll.insert([{"image_path": weird, "label": "x", "embedding": x[0]}])
    assert coll.query(filter=_eq_expr("image_path", weird), output_fields=["label"]) == [{"label": "x"}]
    assert len(coll.query(filter=_in_expr("image_path", [weird, paths[0]]), output_fields=["image_path"])) == 2
    with pytest.raises(ValueError):
        coll.query(filter="label like 'x%'")


def test_similarity_mapping_and_reranker_hook():
    from b200knn.collection import similarity_from_distance

    assert similarity_from_distance(0.9, "COSINE") == 0.9 and similarity_from_distance(0.9, "IP") == 0.9
    assert abs(similarity_from_distance(0.5, "L2") - 0.875) < 1e-12 and similarity_from_distance(1.0, "HAMMING") is None


def test_formats_roundtrip(tmp_path):
    from b200knn import formats as F
    from b200knn.collection import QueryRecord, RetrievedItem, SearchResult

    rs = np.random.RandomState(0)
    n = 12
    emb = rs.standard_normal((n, 4)).astype(np.float32)
    lab = rs.randint(0, 3, n)
    s = emb @ emb.T
    np.fill_diagonal(s, -np.inf)
    cls = {1: {"precision_macro": 1.0, "recall_macro": 2.0, "f1_macro": 3.0, "precision_weighted": 4.0,
               "recall_weighted": 5.0, "f1_weighted": 6.0, "accuracy": 7.0}}
    p = F.save_evaluation_npz(str(tmp_path / "run"), emb, lab, [1, 5, 10], np.array([50.0, 60.0, 70.0], np.float32),
                              0.42, np.array([0.5, 0.4, 0.3]), cls, dists=torch.from_numpy(s))
    z = np.load(p, allow_pickle=True)
    assert sorted(z.files) == sorted(["embeds", "labels", "dists", "kappas", "acc", "mAP", "pr",
                                      "classification_k_values", "classification_k1"])   # test.py:1122-1126
    assert np.isposinf(np.diag(z["dists"])).all() and np.array_equal(z["classification_k1"], np.arange(1.0, 8.0))
    want = np.argsort(np.where(np.isinf(-s), np.nan, -s), axis=1, kind="stable")[:, :3]
    order = np.argsort(-s, axis=1, kind="stable")[:, :5]
    p2 = F.save_evaluation_npz(str(tmp_path / "sparse"), emb, lab, [1], np.array([1.0]), 0.1, np.array([0.1]), cls,
                               topk_dists=-np.take_along_axis(s, order, 1), topk_idx=order)
    pred2, idx2 = F.rank_retrieval(np.load(p2, allow_pickle=True), topk=3)
    assert np.array_equal(idx2, want) and np.array_equal(pred2, lab[want])   # sparse bundles: host slicing only

    q = QueryRecord("q.png", "a")
    res = SearchResult(q, "conv", [RetrievedItem(1, "x.png", "a", 0.9, 0.9), RetrievedItem(2, "y.png", "b", 0.8, 0.8)], [0.0])
    row = F.build_query_analysis_row(q, res, res, True, False, "conv_correct_dino_wrong")
    assert list(row) == ["query_image_path", "query_label", "conv", "dino", "conv_correct", "dino_correct",
                         "assigned_group"]                                                # comparison.py:255-264
    assert list(row["conv"]) == ["image_paths", "labels", "scores", "distances", "hits"]
    flat = F.flatten_query_result(row)
    assert json.loads(flat["conv_topk_image_paths"]) == ["x.png", "y.png"] and flat["group"] == "conv_correct_dino_wrong"
    out = F.write_csv(str(tmp_path / "sub" / "rows.csv"), [flat, {**flat, "extra": 1}])
    with open(out) as fh:
        rows = list(csv.DictReader(fh))
    assert len(rows) == 2 and rows[1]["extra"] == "1" and list(rows[0])[:3] == ["query_image_path", "query_label", "group"]
    items = F.nih_query_results([{"image_path": "a", "image_name": "a", "label_names": ["x"], "multi_hot": [1, 0]}],
                                [[{"id": 0, "score": 0.5, "label_vector": [1, 0]}]])
    assert list(items[0]) == ["query_image_path", "query_image_name", "query_label_names", "query_label_vector", "results"]
    F.write_json(str(tmp_path / "j" / "r.json"), items)
    assert json.load(open(tmp_path / "j" / "r.json"))[0]["results"][0]["score"] == 0.5
    F.save_fused_embeddings(str(tmp_path / "fused.npz"), ["a"], ["x"], emb[:1])
    assert sorted(np.load(tmp_path / "fused.npz").files) == ["embeddings", "image_paths", "labels"]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_rank_retrieval_of_a_dense_bundle_on_the_device(tmp_path):
    """compute_saliency.py:19-29: dense bundles are ranked by the library (knn_rank_rows), the query's own column last."""
    from b200knn import formats as F

    rs = np.random.RandomState(0)
    n = 12
    emb = rs.standard_normal((n, 4)).astype(np.float32)
    lab = rs.randint(0, 3, n)
    s = emb @ emb.T
    np.fill_diagonal(s, -np.inf)
    p = F.save_evaluation_npz(str(tmp_path / "run"), emb, lab, [1], np.array([1.0]), 0.1, np.array([0.1]), {},
                              dists=torch.from_numpy(s))
    pred, idx = F.rank_retrieval(np.load(p, allow_pickle=True), topk=3)
    want = np.argsort(np.where(np.isinf(-s), np.nan, -s), axis=1, kind="stable")[:, :3]
    assert np.array_equal(idx, want) and np.array_equal(pred, lab[want])


@pytest.mark.gpu
@pytest.mark.parametrize("exclude_self", [True, False])
def test_adapter_matches_the_reference_adapter(gc, exclude_self):
    from b200knn.collection import CollectionConfig, LocalCollectionAdapter, QueryRecord

    coll, x, paths, labels = _collection()
    ad = LocalCollectionAdapter(CollectionConfig("conv", "c"), coll)
    c = mgc.CASE
    qi = list(range(0, c["n"], c["n"] // c["n_queries"]))[: c["n_queries"]]
    res = ad.search_by_embeddings([QueryRecord(paths[i], labels[i]) for i in qi], [x[i].tolist() for i in qi],
                                  c["top_k"], exclude_self=exclude_self, batch_size=16)
    want = gc[f"adapter_exclude_self_{exclude_self}"]
    assert len(res) == len(want)
    for r, w in zip(res, want):
        assert r.query.image_path == w["query"] and r.query_source == w["source"]
        assert [it.image_path for it in r.retrieved] == w["image_paths"]
        assert [it.label for it in r.retrieved] == w["labels"] and [it.id for it in r.retrieved] == w["ids"]
        assert np.allclose([it.score for it in r.retrieved], w["scores"], rtol=0, atol=2e-6)
        assert all(it.distance == it.score for it in r.retrieved)     # COSINE "distance" is the similarity (Q12)
    one = ad.search_by_embedding(QueryRecord(paths[qi[0]], labels[qi[0]]), x[qi[0]].tolist(), c["top_k"],
                                 exclude_self=exclude_self)
    assert [it.image_path for it in one.retrieved] == want[0]["image_paths"]

    class Reverse:
        def rerank(self, query, results):
            return list(results)[::-1]

    rr = ad.search_by_embedding(QueryRecord(paths[qi[0]], labels[qi[0]]), x[qi[0]].tolist(), c["top_k"],
                                reranker=Reverse(), exclude_self=False)
    assert [it.image_path for it in rr.retrieved] == gc["adapter_exclude_self_False"][0]["image_paths"][::-1]


@pytest.mark.gpu
def test_nih_collection_flow_matches_the_reference(gc):
    """insert_rows -> search_collection -> hits JSON -> evaluate_results metrics, as query_nih_zilliz.py +
    evaluate_nih_zilliz.py run it against Zilliz."""
    from b200knn import formats as F
    from b200knn import metrics as M
    from b200knn.collection import LocalCollection, insert_rows, search_collection, search_collection_batch

    c = mgc.NIH
    emb, lab = mgc.nih_inputs()
    g, q = emb[: c["n"]], emb[c["n"]:]
    coll = LocalCollection("nih", c["d"], "COSINE")
    insert_rows(coll, [{"image_path": f"nih/{i}.npy", "image_name": f"{i}.npy",
                        "label_names": [str(j) for j in np.flatnonzero(lab[i])], "multi_hot": lab[i].astype(int).tolist(),
                        "embedding": g[i]} for i in range(c["n"])])
    hits = search_collection_batch(coll, [v.tolist() for v in q], c["top_k"])
    assert [[h["id"] for h in hs] for hs in hits] == gc["nih_hit_ids"]
    first = search_collection(coll, q[0].tolist(), c["top_k"])[:5]
    for got, want in zip(first, gc["nih_hits_first_query"]):
        assert {k: got[k] for k in ("id", "image_path", "image_name", "label_text", "label_vector")} == \
               {k: want[k] for k in ("id", "image_path", "image_name", "label_text", "label_vector")}
        assert abs(got["score"] - want["score"]) < 2e-6
    rows = [{"image_path": f"q/{r}.npy", "image_name": f"{r}.npy", "label_names": [], "multi_hot":
             lab[c["n"] + r].astype(int).tolist()} for r in range(c["nq"])]
    items = F.nih_query_results(rows, hits)
    vals = torch.tensor([[h["score"] for h in it["results"]] for it in items], dtype=torch.float32).cuda()
    idx = torch.tensor([[h["id"] for h in it["results"]] for it in items], dtype=torch.int64).cuda()
    m = M.evaluate_results_from_topk(vals, idx, torch.from_numpy(lab[c["n"]:]).cuda(), torch.from_numpy(lab[: c["n"]]).cuda(),
                                     0.4, [1, 5, 10, 20])
    for k, v in gc["nih_metrics"].items():
        assert abs(m[k] - v) < 1e-6, (k, m[k], v)


@pytest.mark.gpu
def test_retriever_result_dicts(gc):
    from b200knn.collection import LocalRetriever

    coll, x, paths, labels = _collection()
    r = LocalRetriever(coll)
    results, qe = r.search(x[5], top_k=4)
    assert list(results[0]) == ["id", "image_path", "label", "distance", "similarity"]   # milvus_retrieval.py:109-115
    assert results[0]["image_path"] == paths[5] and abs(results[0]["similarity"] - 1.0) < 1e-6
    assert qe.shape == (1, x.shape[1]) and abs(float(qe.norm()) - 1.0) < 1e-6
    assert len(r.batch_search([x[1], x[2]], top_k=3)) == 2
    with pytest.raises(ValueError):
        r.search("some/image.png")


@pytest.mark.gpu
def test_l2_collection_distance_conventions():
    """hit.distance of an L2 collection is the Euclidean distance (what milvus_retrieval.py:102-107 assumes);
    l2_squared=True reports the squared distance a real Milvus / faiss.IndexFlatL2 returns -- same ranking."""
    from b200knn.collection import LocalCollection

    rs = np.random.RandomState(3)
    x = rs.standard_normal((50, 16)).astype(np.float32)
    rows = [{"image_path": f"p{i}", "label": "a", "embedding": v} for i, v in enumerate(x)]
    plain, squared = LocalCollection("a", 16, "L2"), LocalCollection("b", 16, "L2", l2_squared=True)
    plain.insert(rows)
    squared.insert(rows)
    h1, h2 = plain.search([x[3]], limit=5)[0], squared.search([x[3]], limit=5)[0]
    assert [h.id for h in h1] == [h.id for h in h2]
    want = np.sqrt(((x[[h.id for h in h1]] - x[3]) ** 2).sum(1))
    assert np.allclose([h.distance for h in h1], want, atol=1e-5)
    assert np.allclose([h.distance for h in h2], want ** 2, atol=1e-5)
