"""Golden outputs for the INPUT side of the late-fusion path, produced by the REAL reference functions.

    python -m oracle.make_golden_sources          (build container only: needs /root/reference)

Run UNMODIFIED: ``FileEmbeddingSource`` / ``build_embedding_source`` / ``align_embedding_sources``
(fusion_eval/align.py:96-229), ``load_query_set`` (retrieval_analysis/comparison.py:41-83) and the result shaping of the
runner, ``experiment_rows`` / ``format_results_table`` (fusion_eval/run_late_fusion.py:55-121).  The golden file holds the
CONTENTS of every input file (so the test can recreate them in a temporary directory) next to what the reference made
of them.  Output: tests/golden/golden_sources.json.
"""
from __future__ import annotations

import json
import os
import tempfile

import numpy as np

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _inputs():
    """Two stores over overlapping image sets, with the irregularities the loaders must cope with: a store without
    labels, a label that differs between the stores, rows only one store has, a row without a label in both."""
    rs = np.random.RandomState(61)
    paths = [f"nih/img_{i:03d}.png" for i in range(14)]
    labels = [f"class{i % 3}" for i in range(14)]
    conv = rs.standard_normal((14, 6)).astype(np.float32)
    dino = rs.standard_normal((14, 4)).astype(np.float32)
    conv_rows = list(range(0, 12))                  # conv has images 0..11
    dino_rows = [13, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]   # dino has 2..13, stored in another order
    files = {}
    files["conv.npz"] = {"kind": "npz", "image_paths": [paths[i] for i in conv_rows],
                         "labels": [labels[i] for i in conv_rows], "embeddings": conv[conv_rows].tolist()}
    files["conv_nolabels.npz"] = {"kind": "npz", "image_paths": [paths[i] for i in conv_rows], "labels": None,
                                  "embeddings": conv[conv_rows].tolist()}
    dino_records = [{"image_path": paths[i], "label": labels[i], "embedding": dino[i].tolist(), "extra": i}
                    for i in dino_rows]
    files["dino.json"] = {"kind": "json", "payload": {"records": dino_records}}
    files["dino_list.json"] = {"kind": "json", "payload": dino_records}            # a bare list is NOT a store
    bad = [dict(r) for r in dino_records]
    bad[3]["label"] = "another"                                                    # label mismatch for one image
    files["dino_mismatch.json"] = {"kind": "json", "payload": {"records": bad}}
    nolab = [dict(r) for r in dino_records]
    for r in nolab:
        r.pop("label")
    files["dino_nolabels.json"] = {"kind": "json", "payload": {"records": nolab}}
    files["dino_dup.json"] = {"kind": "json", "payload": {"records": dino_records + [dino_records[1]]}}
    files["dino_disjoint.json"] = {"kind": "json", "payload": {"records": [
        {"image_path": "other/a.png", "label": "x", "embedding": [0.0, 1.0, 2.0, 3.0]}]}}
    files["store.txt"] = {"kind": "text", "text": "not an embedding store\n"}
    # ordered query sets in the four accepted spellings
    files["queries.json"] = {"kind": "json", "payload": {"queries": [
        {"image_path": paths[9], "label": labels[9]}, {"image_path": paths[3]}, {"image_path": paths[0], "label": "q"},
        {"label": "no path"}, {"image_path": paths[13], "label": labels[13]}, {"image_path": paths[5], "label": labels[5]}]}}
    files["results.json"] = {"kind": "json", "payload": {"results": [
        {"query_image_path": paths[4], "label": labels[4]}, {"image_path": "", "query_image_path": paths[2]}]}}
    files["list.json"] = {"kind": "json", "payload": [{"image_path": paths[7], "label": None}, {"image_path": paths[6]}]}
    files["queries.csv"] = {"kind": "text", "text": "image_path,label\n" + f"{paths[8]},{labels[8]}\n" + f"{paths[2]},\n"
                            + ",orphan\n" + f"{paths[11]},{labels[11]}\n"}
    files["queries2.csv"] = {"kind": "text", "text": "query_image_path,query_label,rank\n" + f"{paths[10]},{labels[10]},1\n"
                             + f"{paths[3]},{labels[3]},2\n"}
    files["queries.txt"] = {"kind": "text", "text": f"# ordered query list\n{paths[6]} {labels[6]}\n\n{paths[2]}\n"
                            + f"  {paths[12]}   {labels[12]}  trailing words\n# {paths[1]}\n{paths[4]}\t{labels[4]}\n"}
    return files


def write_files(files, root):
    """Recreate the input files under ``root`` (shared with tests/test_sources.py)."""
    for name, spec in files.items():
        path = os.path.join(root, name)
        if spec["kind"] == "npz":
            arrays = {"image_paths": np.asarray(spec["image_paths"]),
                      "embeddings": np.asarray(spec["embeddings"], dtype=np.float32)}
            if spec["labels"] is not None:
                arrays["labels"] = np.asarray(spec["labels"])
            np.savez(path, **arrays)
        elif spec["kind"] == "json":
            with open(path, "w", encoding="utf-8") as fh:
                json.dump(spec["payload"], fh)
        else:
            with open(path, "w", encoding="utf-8", newline="") as fh:
                fh.write(spec["text"])


def _records(recs):
    return [{"image_path": r.image_path, "label": r.label, "embedding": np.asarray(r.embedding).tolist(),
             "source_name": r.source_name, "embedding_dtype": str(np.asarray(r.embedding).dtype),
             "raw_keys": sorted(r.raw)} for r in recs]


def _aligned(a):
    return {"image_paths": list(a.image_paths), "labels": list(a.labels),
            "conv_embeddings": a.conv_embeddings.tolist(), "dino_embeddings": a.dino_embeddings.tolist(),
            "conv_dtype": str(a.conv_embeddings.dtype), "coverage": a.coverage}


def _error(fn):
    try:
        fn()
    except Exception as exc:  # noqa: BLE001 - the type and the message are the golden
        return {"type": type(exc).__name__, "message": str(exc)}
    return None


def main():
    A = ref_shim.module("fusion_eval.align")
    C = ref_shim.module("retrieval_analysis.comparison")
    R = ref_shim.module("fusion_eval.run_late_fusion")
    E = ref_shim.module("fusion_eval.evaluate")
    files = _inputs()
    gold = {"files": files, "records": {}, "aligned": {}, "errors": {}, "query_sets": {}}
    with tempfile.TemporaryDirectory() as tmp:
        write_files(files, tmp)
        p = lambda n: os.path.join(tmp, n)  # noqa: E731
        for name in ("conv.npz", "conv_nolabels.npz", "dino.json", "dino_nolabels.json"):
            gold["records"][name] = _records(A.FileEmbeddingSource(p(name), f"src:{name}").fetch_all())
        built = A.build_embedding_source({"type": "file", "path": p("dino.json"), "name": "built"})
        gold["records"]["built:dino.json"] = _records(built.fetch_all())
        src = lambda n: A.FileEmbeddingSource(p(n), n)  # noqa: E731
        gold["aligned"]["conv.npz+dino.json"] = _aligned(A.align_embedding_sources(src("conv.npz"), src("dino.json")))
        gold["aligned"]["conv.npz+dino.json@queries.json"] = _aligned(
            A.align_embedding_sources(src("conv.npz"), src("dino.json"), query_set_path=p("queries.json")))
        gold["aligned"]["conv.npz+dino.json@queries.txt"] = _aligned(
            A.align_embedding_sources(src("conv.npz"), src("dino.json"), query_set_path=p("queries.txt")))
        gold["aligned"]["conv.npz+dino_mismatch.json:lenient"] = _aligned(
            A.align_embedding_sources(src("conv.npz"), src("dino_mismatch.json"), strict_label_check=False))
        gold["aligned"]["conv_nolabels.npz+dino.json:lenient"] = _aligned(
            A.align_embedding_sources(src("conv_nolabels.npz"), src("dino.json"), strict_label_check=False))
        gold["aligned"]["conv_nolabels.npz+dino_nolabels.json"] = _aligned(
            A.align_embedding_sources(src("conv_nolabels.npz"), src("dino_nolabels.json")))
        gold["errors"]["label_mismatch"] = _error(
            lambda: A.align_embedding_sources(src("conv.npz"), src("dino_mismatch.json")))
        gold["errors"]["labels_missing_on_one_side"] = _error(
            lambda: A.align_embedding_sources(src("conv_nolabels.npz"), src("dino.json")))
        gold["errors"]["duplicate_dino"] = _error(lambda: A.align_embedding_sources(src("conv.npz"), src("dino_dup.json")))
        gold["errors"]["duplicate_conv"] = _error(lambda: A.align_embedding_sources(src("dino_dup.json"), src("conv.npz")))
        gold["errors"]["nothing_aligned"] = _error(
            lambda: A.align_embedding_sources(src("conv.npz"), src("dino_disjoint.json")))
        gold["errors"]["unsupported_format"] = _error(lambda: src("store.txt").fetch_all())
        gold["errors"]["json_without_records_object"] = _error(lambda: src("dino_list.json").fetch_all())
        gold["errors"]["unsupported_source_type"] = _error(lambda: A.build_embedding_source({"type": "s3", "name": "x"}))
        for key in ("label_mismatch", "labels_missing_on_one_side", "duplicate_dino", "duplicate_conv", "nothing_aligned",
                    "unsupported_format", "unsupported_source_type", "json_without_records_object"):
            assert gold["errors"][key] is not None, key
            gold["errors"][key]["message"] = gold["errors"][key]["message"].replace(tmp + os.sep, "<tmp>/")
        for name in ("queries.json", "results.json", "list.json", "queries.csv", "queries2.csv", "queries.txt"):
            gold["query_sets"][name] = [[q.image_path, q.label] for q in C.load_query_set(p(name))]
    exps = [E.ExperimentResult("convnext_baseline", 12, {"mP@1": 91.66666, "mP@5": 80.0, "mP@10": 71.25, "R@1": 2.5,
                                                         "R@5": 11.0049, "R@10": 19.995, "mAP": 66.6651, "num_samples": 12.0}),
            E.ExperimentResult("weighted_sum_alpha_0.5", 12, {}, True, "embedding dimensions differ: 6 vs 4"),
            E.ExperimentResult("score_fusion_alpha_0.2", 12, {"mP@1": 100.0, "R@1": 3.0, "mAP": 70.005}),
            E.ExperimentResult("skipped_without_reason", 0, {}, True, None)]
    gold["experiments"] = [{"experiment_name": e.experiment_name, "num_samples": e.num_samples, "metrics": e.metrics,
                            "skipped": e.skipped, "skipped_reason": e.skipped_reason} for e in exps]
    rows = R.experiment_rows(exps)
    gold["experiment_rows"] = rows
    gold["results_table"] = R.format_results_table(rows)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "golden_sources.json"), "w", encoding="utf-8") as fh:
        json.dump(gold, fh, indent=1, sort_keys=True)
    print("wrote", os.path.join(OUT, "golden_sources.json"))


if __name__ == "__main__":
    main()
