"""CPU: the C-ABI library loads and exports exactly what include/b200knn.h declares; argument validation
that needs no device returns the documented error codes."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200knn.h")).read()
    return sorted(set(re.findall(r"KNN_API\s+[\w\s\*]+?\b(knn_\w+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for must in ("knn_normalize", "knn_search", "knn_search_workspace", "knn_merge_topk", "knn_rank_rows",
                 "knn_relevance_single", "knn_relevance_multilabel", "knn_ranked_stats", "knn_map_full"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from b200knn import _lib

    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/b200knn.h but not exported"
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    assert lib.knn_version() == 1


def test_argument_validation_without_a_device():
    from b200knn import _lib

    lib = _lib.load()
    # bad dtype / metric / k are rejected before any CUDA call
    rc = lib.knn_search(None, None, None, None, 4, 4, 8, 7, 1, 0, 0, 0, 0, None, None, None, 0, None)
    assert rc == -1 and b"dtype" in lib.knn_last_error()
    rc = lib.knn_search(None, None, None, None, 4, 4, 8, 0, 1, 9, 0, 0, 0, None, None, None, 0, None)
    assert rc == -1 and b"metric" in lib.knn_last_error()
    rc = lib.knn_search(None, None, None, None, 4, 4, 8, 0, 0, 0, 0, 0, 0, None, None, None, 0, None)
    assert rc == -1  # null q/g
    rc = lib.knn_normalize(None, None, None, 4, 0, 0, 0, 1e-12, 0, None)
    assert rc == -1
    rc = lib.knn_merge_topk(None, None, 2, 4, 8, 0, None, None, None)
    assert rc == -1
    # KNN_BF16X3 rows are three parts of a multiple of 8 columns
    rc = lib.knn_search(None, None, None, None, 4, 4, 40, 2, 1, 0, 0, 0, 0, None, None, None, 0, None)
    assert rc == -1 and b"KNN_BF16X3" in lib.knn_last_error()
    assert lib.knn_search_workspace(300, 100000, 192, 2, 64) == lib.knn_search_workspace(300, 100000, 192, 1, 64)
    rc = lib.knn_rescore_exact(None, None, None, None, 4, 4, 8, 0, 0, 0, 0, None, None, 8, 16, None, None, None, None, None)
    assert rc == -1 and b"kc" in lib.knn_last_error()            # k must not exceed the candidate count
    rc = lib.knn_lesion_rerank(None, None, 4, 10, 11, None, None, None, None, 4, 1, 8, 0.5, None, None, None, None)
    assert rc == -1
    rc = lib.knn_merge_topk_parts(None, None, 17, 4, 8, 0, None, None, None)
    assert rc == -1
    assert lib.knn_launch_count() >= 0
    assert lib.knn_search_workspace(0, 100, 8, 0, 10) == 0
    assert lib.knn_search_workspace(300, 100000, 64, 0, 100) > 0
    assert lib.knn_rank_rows_workspace(4, 100) == 0 and lib.knn_rank_rows_workspace(4, 5000) == 4 * 8192 * 8


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "image-retrieval---thesis-2026_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "oracle" not in src.replace("# oracle", ""), name


def test_cpu_tensors_are_rejected_loudly():
    import torch

    import b200knn

    with pytest.raises(b200knn.KnnError):
        b200knn.search(torch.zeros(2, 8), torch.zeros(4, 8), 1)
    with pytest.raises(b200knn.KnnError):
        b200knn.normalize(torch.zeros(2, 8))


def test_exact_engine_host_logic():
    """Host-side decisions of the tensor-core exact engine (no device needed)."""
    import importlib

    S = importlib.import_module("b200knn.search")
    assert [S._filter_k(k) for k in (1, 10, 24, 25, 50, 100, 113, 128, 228, 229)] == [32, 32, 32, 64, 64, 128, 128, 256,
                                                                                      256, 512]
    assert S.exact_engine(25_000, 112_000, 1024, 50) == "tensor"      # BASELINE config 3
    assert S.exact_engine(400, 400, 1024, 10) == "ffma"                # config 1: sub-GFLOP, launch bound
    assert S.exact_engine(600, 2000, 256, 10) == "ffma"                # config 2
    assert S.exact_engine(25_000, 112_000, 1024, 250) == "ffma"        # no room for slack candidates below 256
    assert S.rerun_ranges([], 1000) == []
    assert S.rerun_ranges([5, 130, 131, 600], 2000) == [(0, 256), (512, 640)]
    assert S.rerun_ranges([1999], 2000) == [(1920, 2000)]
    assert S.rerun_ranges(list(range(0, 2000, 128)), 2000) == [(0, 2000)]


def test_search_geometry_of_the_seeding_pre_pass():
    """knn_search_geometry is host arithmetic: the threshold-seeding pre-pass of the large-batch bf16 path collects chunk
    maxima (seed_stride >= 1), its lists fit their capacity, the sample stays below 2 % of the gallery; small batches
    and fp32 keep the selecting pre-pass."""
    from b200knn import _lib

    lib = _lib.load()

    def geo(nq, ng, d, dtype, k):
        out = (C.c_int64 * 8)()
        assert lib.knn_search_geometry(nq, ng, d, dtype, k, out) == 0, lib.knn_last_error()
        return dict(zip(("qblocks", "splits", "groups", "split_len", "L", "seed_units", "seed_len", "seed_stride"), out))

    for nq, ng, d, k in [(8192, 50_000_000, 512, 100), (8192, 6_250_000, 512, 100), (8192, 2_000_000, 512, 100),
                         (1280, 820_000, 64, 10), (1280, 820_000, 64, 256), (4096, 30_000_000, 768, 32)]:
        g = geo(nq, ng, d, 1, k)
        assert g["seed_stride"] >= 1 and g["seed_units"] >= 1, (nq, ng, k, g)
        assert g["seed_len"] % 256 == 0 and g["seed_units"] * g["seed_len"] <= ng // 50
        per_thread = g["seed_len"] // 256 * 8 // g["groups"]            # chunks of one selection thread in a unit
        keys = -(-per_thread // g["seed_stride"])
        assert keys <= g["L"], g                                          # a list never overflows
        kp = 32
        while kp < k:
            kp *= 2
        assert g["seed_units"] * g["groups"] * keys >= 2 * kp, g          # enough maxima for a useful bound
    g = geo(64, 10_000_000, 768, 1, 100)                                  # one query block: TMEM-resident kernel,
    assert g["seed_stride"] == 1 << 30 and g["seed_units"] * g["groups"] >= 200   # one maximum per (unit, thread)
    assert geo(64, 8_000_000, 1024, 1, 100)["seed_stride"] == 0           # d > 768: one-CTA shared-memory-A kernel
    assert geo(1024, 5_000_000, 512, 1, 100)["seed_stride"] == 0          # <= 1024 rows: seeds from list maxima
    assert geo(8192, 2_000_000, 512, 0, 100)["seed_stride"] == 0          # fp32 FFMA kernel
    out = (C.c_int64 * 8)()
    assert lib.knn_search_geometry(0, 10, 8, 0, 1, out) == -1 and lib.knn_search_geometry(4, 10, 8, 9, 1, out) == -1


def test_search_geometry_invariants_over_random_problems():
    """Property sweep of the host-side work decomposition: the splits cover the gallery, the seeding sample is a prefix of
    at most 2 % of it, a maxima-mode list never holds more keys than its capacity, and the workspace query grows with
    the decomposition it describes."""
    import random

    from b200knn import _lib

    lib = _lib.load()
    rnd = random.Random(20261019)
    names = ("qblocks", "splits", "groups", "split_len", "L", "seed_units", "seed_len", "seed_stride")
    for _ in range(3000):
        nq = rnd.choice([1, 7, 64, 128, 129, 300, 1024, 1025, 2048, 5000, 8192, 25000, 40000])
        ng = int(10 ** rnd.uniform(2, 8.2))
        d = rnd.choice([8, 64, 256, 512, 768, 1024, 2048])
        k = rnd.choice([1, 10, 32, 33, 50, 100, 128, 200, 256])
        dtype = rnd.choice([0, 1])
        out = (C.c_int64 * 8)()
        assert lib.knn_search_geometry(nq, ng, d, dtype, k, out) == 0, lib.knn_last_error()
        g = dict(zip(names, out))
        tag = (nq, ng, d, k, dtype, g)
        assert g["qblocks"] * 128 >= nq and g["groups"] in (1, 2), tag
        assert g["splits"] >= 1 and g["splits"] * g["split_len"] >= ng and (g["splits"] - 1) * g["split_len"] < ng, tag
        kp = 32
        while kp < k:
            kp *= 2
        assert g["L"] == 2 * kp, tag
        if g["seed_units"] == 0:
            assert g["seed_stride"] == 0 and g["seed_len"] == 0, tag
            continue
        assert g["seed_units"] * g["seed_len"] <= ng // 50, tag
        if g["seed_stride"] == 0:
            continue
        assert dtype == 1, tag
        if g["seed_stride"] == 1 << 30:                               # one maximum per (unit, selection thread)
            assert g["qblocks"] == 1 and 2 * kp <= g["seed_units"] * g["groups"] <= 4096, tag
            assert g["seed_len"] % 32 == 0 or g["seed_len"] == 224, tag
        else:
            assert g["qblocks"] * 128 > 1024 and g["seed_len"] % 256 == 0, tag
            per_thread = g["seed_len"] // 256 * 8 // g["groups"]
            keys = -(-per_thread // g["seed_stride"])
            assert 1 <= keys <= g["L"], tag
            assert g["seed_units"] * g["groups"] * keys >= 2 * kp, tag
        assert lib.knn_search_workspace(nq, ng, d, dtype, k) > 0


def test_two_product_filter_rule(monkeypatch):
    """Host logic of the exact engine: the two-product filter is tried first only for keep-mode batches of >= 1024
    queries over a gallery (shard) large enough that a third of the filter outweighs the wider re-scoring."""
    import importlib

    S = importlib.import_module("b200knn.search")        # (`b200knn.search` the attribute is the function)
    monkeypatch.delenv("KNN_EXACT_PRODUCTS", raising=False)
    assert S._two_product_filter(25_000, 112_000, "keep")
    assert not S._two_product_filter(25_000, 14_000, "keep")         # the 8-GPU shard of config 3
    assert not S._two_product_filter(512, 112_000, "keep")
    assert not S._two_product_filter(25_000, 112_000, "exclude")
    monkeypatch.setenv("KNN_EXACT_PRODUCTS", "2")
    assert S._two_product_filter(25_000, 14_000, "keep")
    monkeypatch.setenv("KNN_EXACT_PRODUCTS", "3")
    assert not S._two_product_filter(25_000, 112_000, "keep")
