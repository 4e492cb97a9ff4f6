"""Generate tests/golden/* by running the REAL, unmodified reference (from /root/reference) on seeded inputs.

    python -m oracle.make_golden          (build container only: needs /root/reference)

The fixtures hold seeds + the reference's outputs, never its sources.  Where the reference ranks with an
unstable sort, the ranking handed to its metric functions is the stable order (score best-first, ties by
ascending index) of the reference's OWN score matrix, also recorded here.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import re
import tempfile
import types
import warnings

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "c1": dict(n=400, d=1024, classes=3, seed=0, noise=5.0),
    "c1_exact": dict(n=400, d=1024, seed=1, n_dup=8),
    "c2": dict(nq=600, ng=2000, d=256, classes=3, seed=2, noise=5.0, priors=[0.67, 0.17, 0.16]),
    "c3s": dict(nq=64, ng=512, d=128, seed=3, noise=1.2, k=50),
    "ml_self": dict(n=200, d=64, seed=7, noise=1.0),
    "sl_self": dict(n=300, d=128, classes=4, seed=11, noise=2.5),
}


def stable_desc_rank_cols(S: torch.Tensor) -> np.ndarray:
    """column-wise stable descending ranking of the reference's own matrix -> [N, nq] like argsort(dim=0)."""
    s = S.cpu().numpy()
    return np.argsort(-s, axis=0, kind="stable")


def stable_topk_rows(S: np.ndarray, k: int, largest=True):
    order = np.argsort(-S if largest else S, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(S, order, axis=1), order


class _Identity(torch.nn.Module):
    def __init__(self, as_dict=False):
        super().__init__()
        self.as_dict = as_dict

    def forward(self, x):
        return {"embedding": x} if self.as_dict else x


def _f(x):
    return float(x)


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # fixed MKL blocking -> reproducible reference matrices
    ref_test = ref_shim.module("test")
    ref_train = ref_shim.module("train")
    ref_ath = ref_shim.module("test_ath")
    ref_nihz = ref_shim.module("evaluate_nih_zilliz")
    ref_nih = ref_shim.module("nih_multilabel_training")
    ref_fus = ref_shim.module("fusion_eval.metrics")
    G = {"cases": CASES, "torch": torch.__version__, "numpy": np.__version__}
    A = {}

    # ---------------- C1: 400 x 400 x 1024 self-retrieval, cosine (test.py:1005-1007 style) ----------------
    c = CASES["c1"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    E = F.normalize(torch.from_numpy(x), p=2, dim=1)
    labels = torch.from_numpy(lab)
    S = torch.mm(E, E.t())
    S.fill_diagonal_(-float("inf"))
    acc = ref_test.retrieval_accuracy(S, labels, topk=(1, 5, 10))
    ranks = stable_desc_rank_cols(S)
    mAP, aps, pr, prs = ref_test.compute_map(ranks, lab, [1, 5, 10])
    cls = ref_test.compute_classification_metrics(labels, S, [1, 5, 10, 15, 20])
    G["c1_cosine"] = {"acc": [_f(a) for a in acc], "mAP": _f(mAP), "pr": [_f(v) for v in pr],
                      "classification": {str(k): {m: _f(v) for m, v in d.items()} for k, d in cls.items()}}
    A["c1_cosine_aps"] = aps
    A["c1_cosine_prs"] = prs
    tv, ti = stable_topk_rows(S.numpy(), 10)
    A["c1_cosine_top10_val"], A["c1_cosine_top10_idx"] = tv, ti
    A["c1_normalized_rows_0_8"] = E[:8].numpy()
    A["c1_cosine_ranks_rowmajor"] = np.ascontiguousarray(ranks.T).astype(np.int16)  # the reference's own full ranking

    # ---------------- C1 through the reference's own evaluate() (-cdist, test.py:1066-1126) ----------------
    with tempfile.TemporaryDirectory() as tmp:
        args = types.SimpleNamespace(save_dir=tmp, resume="golden/c1.pth")
        with contextlib.redirect_stdout(io.StringIO()):
            ref_test.evaluate(_Identity(), [(E, labels)], torch.device("cpu"), args)
        z = np.load(os.path.join(tmp, "c1.npz"))
        G["c1_cdist"] = {"acc": [_f(v) for v in z["acc"]], "mAP": _f(z["mAP"]), "pr": [_f(v) for v in z["pr"]],
                         "classification": {str(k): [_f(v) for v in z[f"classification_k{k}"]]
                                            for k in z["classification_k_values"]}}
        dists = z["dists"]  # positive distances, +inf diagonal
        tv, ti = stable_topk_rows(dists, 10, largest=False)
        A["c1_cdist_top10_val"], A["c1_cdist_top10_idx"] = tv, ti
        A["c1_cdist_ranks_rowmajor"] = np.ascontiguousarray(np.argsort(dists, axis=0, kind="stable").T).astype(np.int16)

    # ---------------- C1 strict-exact grid: ip and l2, real ties (duplicated rows) ----------------
    c = CASES["c1_exact"]
    xe = torch.from_numpy(synth.exact_grid(c["n"], c["d"], c["seed"], c["n_dup"]))
    Se = xe @ xe.t()
    Se.fill_diagonal_(-float("inf"))
    A["c1_exact_ip_top10_val"], A["c1_exact_ip_top10_idx"] = stable_topk_rows(Se.numpy(), 10)
    De = torch.cdist(xe, xe)
    De.fill_diagonal_(float("inf"))
    A["c1_exact_l2_top10_val"], A["c1_exact_l2_top10_idx"] = stable_topk_rows(De.numpy(), 10, largest=False)

    # ---------------- C2: 600 q x 2000 g x 256, L2, test_ath.compute_metrics (test_ath.py:90-172) ----------------
    c = CASES["c2"]
    x, lab = synth.clustered(c["nq"] + c["ng"], c["d"], c["classes"], c["seed"], c["noise"], c["priors"])
    En = F.normalize(torch.from_numpy(x), p=2, dim=1)
    q, g = En[: c["nq"]], En[c["nq"]:]
    ql, gl = torch.from_numpy(lab[: c["nq"]]), torch.from_numpy(lab[c["nq"]:])
    logits = torch.zeros((c["nq"], c["classes"]))
    out = ref_ath.compute_metrics(q, ql, g, gl, logits, (1, 5, 10), False)
    G["c2_l2"] = {"retrieval": {str(k): {m: _f(v) for m, v in d.items()} for k, d in out["retrieval"].items()}}
    D2 = torch.cdist(q.float(), g.float(), p=2).numpy()
    A["c2_top10_val"], A["c2_top10_idx"] = stable_topk_rows(D2, 10, largest=False)

    # ---------------- C3 (small): multilabel query x gallery, top-50, evaluate_nih_zilliz ----------------
    c = CASES["c3s"]
    lab_all = synth.multihot(c["nq"] + c["ng"], c["seed"])
    emb = F.normalize(torch.from_numpy(synth.labelset_clustered(lab_all, c["d"], c["seed"] + 100, c["noise"])), dim=1)
    q, g = emb[: c["nq"]], emb[c["nq"]:]
    qlab, glab = lab_all[: c["nq"]], lab_all[c["nq"]:]
    S3 = (q @ g.t()).numpy()
    tv, ti = stable_topk_rows(S3, c["k"])
    items = [{"query_label_vector": qlab[i].tolist(),
              "results": [{"score": float(tv[i, j]), "label_vector": glab[ti[i, j]].tolist()} for j in range(c["k"])]}
             for i in range(c["nq"])]
    G["c3s_nih"] = {k: _f(v) for k, v in ref_nihz.evaluate_results(items, 0.4, [1, 5, 10, 20, 50]).items()}
    A["c3s_top50_val"], A["c3s_top50_idx"] = tv, ti

    # ---------------- multilabel self-retrieval: train.py D7, test.py D4 + D5, nih evaluate_map D8 ----------------
    c = CASES["ml_self"]
    mlab = synth.multihot(c["n"], c["seed"])
    memb = torch.from_numpy(synth.labelset_clustered(mlab, c["d"], c["seed"] + 100, c["noise"]))
    tl = torch.from_numpy(mlab)
    G["ml_self_train"] = {k: _f(v) for k, v in ref_train._compute_multilabel_retrieval_metrics(memb, tl).items()}
    mn = F.normalize(memb, p=2, dim=1)
    Sm = torch.mm(mn, mn.t())
    Sm.fill_diagonal_(-float("inf"))
    G["ml_self_map_multilabel"] = {str(t): _f(ref_test.compute_map_multilabel(Sm, tl, threshold=t)) for t in (0.25, 0.5)}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref_test.evaluate_multilabels(_Identity(), [(memb, tl)], torch.device("cpu"), types.SimpleNamespace(save_dir=None))
    rows = re.findall(r"^(\d+)\s+\|\s+([\d.]+)\s+%\s+\|\s+([\d.]+)\s+%", buf.getvalue(), flags=re.M)
    G["ml_self_prk_printed"] = {r[0]: [float(r[1]), float(r[2])] for r in rows}
    G["ml_self_evaluate_map"] = _f(ref_nih.evaluate_map(_Identity(as_dict=True), [(memb, tl)], torch.device("cpu"), 0.4))

    # ---------------- single-label self-retrieval: train.py D6, fusion_eval D11 ----------------
    c = CASES["sl_self"]
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    G["sl_self_train"] = {k: _f(v) for k, v in
                          ref_train._compute_single_label_retrieval_metrics(torch.from_numpy(x), torch.from_numpy(lab)).items()}
    G["sl_self_fusion"] = {k: _f(v) for k, v in ref_fus.evaluate_retrieval_metrics(
        x, [f"class{v}" for v in lab], [f"img{i}.png" for i in range(len(lab))], (1, 5, 10)).items()}

    with open(os.path.join(OUT, "golden.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "golden_arrays.npz"), **A)
    print("wrote", OUT, {k: (v.shape, str(v.dtype)) for k, v in A.items()})
    print(json.dumps({k: v for k, v in G.items() if k not in ("cases",)}, indent=1)[:3000])


if __name__ == "__main__":
    main()
