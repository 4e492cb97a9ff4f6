"""CPU: the reference-signature ``metrics.evaluate_results(items, jaccard_threshold, ks)`` (evaluate_nih_zilliz.py:34-64).
The device work is ``evaluate_results_from_topk`` (GPU-tested against the same reference function in
tests/test_gpu_metrics.py and tests/test_collection_formats.py); checked here, without a device: the arrays built from
the hits JSON mean what the reference's loop means (the numpy restatement of the metric, fed with them, reproduces the
REAL function's result: tests/golden/golden_nih_eval.json, oracle/make_golden_nih_eval.py), the edge cases answered on
the host, and the hand-over to the device function."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import reference_metrics as rm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_nih_eval.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def M():
    return importlib.import_module("b200knn.metrics")


def test_arrays_built_from_the_items_carry_the_reference_semantics(M, gold):
    vals, idx, qlab, glab = M.hits_to_arrays(gold["items"])
    nq, k = len(gold["items"]), len(gold["items"][0]["results"])
    assert vals.shape == idx.shape == (nq, k) and vals.dtype == np.float32 and idx.dtype == np.int64
    assert qlab.shape == (nq, 14) and glab.shape == (nq * k, 14) and qlab.dtype == glab.dtype == np.float32
    assert np.array_equal(idx, np.arange(nq * k).reshape(nq, k))
    assert vals[3, 2] == np.float32(gold["items"][3]["results"][2]["score"])
    assert glab[3 * k + 2].tolist() == gold["items"][3]["results"][2]["label_vector"]
    for thr, ks, key in ((gold["threshold"], gold["ks"], "metrics"), (0.0, [1, 3], "metrics_thr_0")):
        got = rm.evaluate_results(vals, idx, qlab, glab, thr, ks)
        assert list(got) == list(gold[key])
        for name, want in gold[key].items():
            assert got[name] == pytest.approx(want, abs=1e-9), (key, name)


def test_edge_cases_are_answered_like_the_reference(M, gold):
    assert M.evaluate_results([], 0.4, [1, 5]) == gold["empty"]
    no_hits = [{"query_label_vector": [1, 0], "results": []}, {"query_label_vector": [0, 1], "results": []}]
    assert M.evaluate_results(no_hits, 0.4, [1, 5]) == gold["no_hits"]
    assert list(M.evaluate_results([], 0.4, [1, 5])) == list(gold["empty"])
    ragged = [dict(gold["items"][0]), dict(gold["items"][1], results=gold["items"][1]["results"][:3])]
    with pytest.raises(ValueError):
        M.hits_to_arrays(ragged)
    with pytest.raises(ValueError):
        M.evaluate_results([no_hits[0], gold["items"][0]], 0.4, [1])


def test_hand_over_to_the_device_function(M, gold, monkeypatch):
    torch = importlib.import_module("torch")
    seen = {}

    def fake(vals, idx, qlab, glab, thr, ks):
        seen.update(vals=vals, idx=idx, qlab=qlab, glab=glab, thr=thr, ks=ks)
        return {"mAP": 1.0}

    monkeypatch.setattr(M, "evaluate_results_from_topk", fake)
    out = M.evaluate_results(gold["items"], 0.4, (1, 5), device=torch.device("cpu"))
    assert out == {"mAP": 1.0} and seen["thr"] == 0.4 and seen["ks"] == [1, 5]
    want = M.hits_to_arrays(gold["items"])
    for got, w in zip((seen["vals"], seen["idx"], seen["qlab"], seen["glab"]), want):
        assert isinstance(got, torch.Tensor) and np.array_equal(got.numpy(), w)


@pytest.mark.gpu
def test_device_result_equals_the_real_reference_function(M, gold):
    """The whole entry on a B200: hits JSON in, the REAL ``evaluate_results``' metric dictionary out (keys in its order)."""
    for thr, ks, key in ((gold["threshold"], gold["ks"], "metrics"), (0.0, [1, 3], "metrics_thr_0")):
        got = M.evaluate_results(gold["items"], thr, ks)
        assert list(got) == list(gold[key])
        for name, want in gold[key].items():
            assert abs(got[name] - want) < 1e-9, (key, name, got[name], want)
