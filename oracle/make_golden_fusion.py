"""Golden vectors for the late-fusion / re-ranking row, produced by the REAL reference.

    python -m oracle.make_golden_fusion          (build container only: needs /root/reference)

Runs ``fusion_eval.evaluate.run_late_fusion_experiments`` / ``normalize_similarity_matrix`` /
``confidence_based_fusion`` of the unmodified reference on seeded inputs and records their outputs.  The text
re-ranking of test.py:608-623 is inline code of a long script function, so its golden matrix comes from the torch
statements of those lines re-executed here verbatim on seeded tensors (`_rerank_like_test_py`).
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np
import torch

from . import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "fusion": dict(n=240, d_conv=96, d_dino=64, classes=4, seed=21, noise_conv=3.0, noise_dino=2.0),
    "fusion_same_dim": dict(n=120, d=48, classes=3, seed=23, noise_conv=2.5, noise_dino=1.5),
    "rerank": dict(n=150, d=64, classes=5, seed=31, noise=2.5, rerank_k=20, text_weight=0.7),
}


def fusion_inputs(c):
    conv, lab = synth.clustered(c["n"], c.get("d_conv", c.get("d")), c["classes"], c["seed"], c["noise_conv"])
    rs = np.random.RandomState(c["seed"] + 1)
    mu = rs.standard_normal((c["classes"], c.get("d_dino", c.get("d"))))
    dino = (mu[lab] + c["noise_dino"] * rs.standard_normal((c["n"], mu.shape[1]))).astype(np.float32)
    return conv, dino, lab


def rerank_inputs(c):
    x, lab = synth.clustered(c["n"], c["d"], c["classes"], c["seed"], c["noise"])
    rs = np.random.RandomState(c["seed"] + 1)
    text = rs.standard_normal((c["classes"], c["d"])).astype(np.float32)
    return x, lab, text


def _rerank_like_test_py(embeds, text_embeds, labels, rerank_k, text_weight):
    # test.py:604-623, statement for statement
    img_sim = embeds @ embeds.t()
    img_text_sim = embeds @ text_embeds.t()
    dists = img_sim.clone()
    alpha = text_weight
    beta = 1.0 - alpha
    for i in range(len(labels)):
        top_k_scores, top_k_indices = torch.topk(img_sim[i], k=min(rerank_k, len(labels)), largest=True)
        for j in top_k_indices:
            if i != j:
                text_score = img_text_sim[j, labels[i]]
                dists[i, j] = alpha * img_sim[i, j] + beta * text_score
    dists.fill_diagonal_(float("-inf"))
    return img_sim, img_text_sim, dists


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    ev = ref_shim.module("fusion_eval.evaluate")
    al = ref_shim.module("fusion_eval.align")
    G, A = {"cases": CASES, "numpy": np.__version__, "torch": torch.__version__}, {}

    for name in ("fusion", "fusion_same_dim"):
        c = CASES[name]
        conv, dino, lab = fusion_inputs(c)
        labels = [f"class{v}" for v in lab]
        paths = [f"img{i}.png" for i in range(len(lab))]
        aligned = al.AlignedEmbeddings(image_paths=paths, labels=labels, conv_embeddings=conv, dino_embeddings=dino,
                                       coverage={})
        for mode in ("none", "zscore", "minmax"):
            res = ev.run_late_fusion_experiments(aligned, alpha_values=(0.2, 0.5, 0.8), k_values=(1, 5, 10),
                                                 include_score_fusion=True, score_normalization=mode,
                                                 include_confidence_fusion=True)
            G[f"{name}_{mode}"] = [{"experiment_name": r.experiment_name, "num_samples": r.num_samples,
                                    "metrics": {k: float(v) for k, v in r.metrics.items()}, "skipped": bool(r.skipped),
                                    "skipped_reason": r.skipped_reason} for r in res]
        if name == "fusion":
            from fusion_eval.metrics import compute_similarity_matrix
            for mode in ("zscore", "minmax"):
                cs = ev.normalize_similarity_matrix(compute_similarity_matrix(ev.l2_normalize(conv)), mode)
                ds = ev.normalize_similarity_matrix(compute_similarity_matrix(ev.l2_normalize(dino)), mode)
                fused = 0.5 * cs + 0.5 * ds
                f = fused.copy()
                np.fill_diagonal(f, -np.inf)
                order = np.argsort(-f, axis=1, kind="stable")[:, :10]
                A[f"fusion_{mode}_a0.5_top10_val"] = np.take_along_axis(f, order, axis=1)
                A[f"fusion_{mode}_a0.5_top10_idx"] = order
                A[f"fusion_{mode}_conv_rows_0_4"] = cs[:4]
            conf = ev.confidence_based_fusion(
                ev.normalize_similarity_matrix(compute_similarity_matrix(ev.l2_normalize(conv)), "none"),
                ev.normalize_similarity_matrix(compute_similarity_matrix(ev.l2_normalize(dino)), "none"))
            f = conf["similarity"].copy()
            np.fill_diagonal(f, -np.inf)
            order = np.argsort(-f, axis=1, kind="stable")[:, :10]
            A["fusion_conf_top10_val"] = np.take_along_axis(f, order, axis=1)
            A["fusion_conf_top10_idx"] = order
            G["fusion_conf"] = {k: (float(v) if not isinstance(v, int) else v) for k, v in conf.items()
                                if k != "similarity"}

    c = CASES["rerank"]
    x, lab, text = rerank_inputs(c)
    e = torch.from_numpy(x)
    e = e / e.norm(dim=-1, keepdim=True)          # test.py:597
    t = torch.from_numpy(text)
    t = t / t.norm(dim=-1, keepdim=True)
    img_sim, img_text_sim, dists = _rerank_like_test_py(e, t, torch.from_numpy(lab), c["rerank_k"], c["text_weight"])
    d = dists.numpy()
    order = np.argsort(-d, axis=1, kind="stable")[:, :10]
    A["rerank_top10_val"] = np.take_along_axis(d, order, axis=1)
    A["rerank_top10_idx"] = order
    A["rerank_dists_rows_0_4"] = d[:4]
    A["rerank_table"] = img_text_sim.numpy()

    with open(os.path.join(OUT, "golden_fusion.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "golden_fusion_arrays.npz"), **A)
    print("wrote", {k: v.shape for k, v in A.items()})
    print(json.dumps(G["fusion_none"][:4], indent=1)[:1500])


if __name__ == "__main__":
    main()
