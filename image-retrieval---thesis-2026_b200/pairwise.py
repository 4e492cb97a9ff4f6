"""Training-time pairwise operations over a batch, forward values only (SURVEY 8(f)-4 tail).

* :func:`batch_hard_triplet_loss`, :func:`batch_all_triplet_loss` -- loss.py:60-112 (``TripletMarginLoss`` mining)
* :func:`compute_jaccard_sim`                                    -- loss.py:237-242 / 158-173
* :func:`nearest_centroid_scores`                               -- anomaly/test_anomaly.py:31-48

The pairwise distance matrix comes from the library's fused distance kernel (``knn_scores_dense``, GEMM form of
``torch.cdist``), the reductions over it from csrc/pairwise.cu.  These are the VALUES the reference's training / anomaly
scripts compute; gradients stay with the training framework (out of scope).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .metrics import pack_multihot
from .search import _ptr, _require_cuda, _stream, scores_dense


def _mine(embeddings: torch.Tensor, labels: torch.Tensor, margin: float, want_all: bool):
    _require_cuda(embeddings)
    e = embeddings.float().contiguous()
    n = e.shape[0]
    dev = e.device
    lab = labels.to(device=dev, dtype=torch.int64).contiguous().view(-1)
    dist = scores_dense(e, e, "l2")                                 # torch.cdist(embeddings, embeddings, p=2)
    hard = torch.empty((n,), dtype=torch.float32, device=dev)
    s = torch.empty((n,), dtype=torch.float64, device=dev) if want_all else None
    pos = torch.empty((n,), dtype=torch.int64, device=dev) if want_all else None
    val = torch.empty((n,), dtype=torch.int64, device=dev) if want_all else None
    with torch.cuda.device(dev):
        rc = L.load().knn_triplet_mine(_ptr(dist), _ptr(lab), n, float(margin), _ptr(hard), _ptr(s), _ptr(pos), _ptr(val),
                                       _stream(e))
    L.check(rc, "knn_triplet_mine")
    return hard, s, pos, val


def batch_hard_triplet_loss(labels: torch.Tensor, embeddings: torch.Tensor, margin: float, p: float = 2.0):
    """loss.py:60-83: mean over the anchors of max(hardest positive - hardest negative + margin, 0) -> (loss, -1)."""
    if p != 2.0:
        raise ValueError("only the Euclidean distance (p=2) of the reference's configurations is built")
    hard, _, _, _ = _mine(embeddings, labels, margin, False)
    return float(hard.cpu().numpy().astype(np.float64).mean()), -1   # per-anchor fp32 terms, float64 mean on the host


def batch_all_triplet_loss(labels: torch.Tensor, embeddings: torch.Tensor, margin: float, p: float = 2.0) -> Tuple[float, float]:
    """loss.py:86-112 -> (mean of the positive triplet terms, fraction of the valid triplets that are positive)."""
    if p != 2.0:
        raise ValueError("only the Euclidean distance (p=2) of the reference's configurations is built")
    _, s, pos, val = _mine(embeddings, labels, margin, True)
    packed = torch.stack([s, pos.double(), val.double()]).sum(dim=1).cpu().numpy()   # one read-back
    total, npos, nvalid = float(packed[0]), float(packed[1]), float(packed[2])
    return total / (npos + 1e-16), npos / (nvalid + 1e-16)


def compute_jaccard_sim(labels_multihot: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """loss.py:237-242: ``intersection / (union + eps)`` for every pair of multi-hot rows -> fp32 [B, B]."""
    _require_cuda(labels_multihot)
    m = pack_multihot(labels_multihot).contiguous()
    n = m.shape[0]
    out = torch.empty((n, n), dtype=torch.float32, device=m.device)
    with torch.cuda.device(m.device):
        rc = L.load().knn_jaccard_matrix(_ptr(m), _ptr(m), n, n, float(eps), _ptr(out), _stream(m))
    L.check(rc, "knn_jaccard_matrix")
    return out


def class_means(embeds: torch.Tensor, labels: torch.Tensor, classes: Sequence[int]) -> torch.Tensor:
    """``embeds[labels == c].mean(axis=0)`` for every c (anomaly/test_anomaly.py:31-32) -> fp32 [len(classes), D]."""
    _require_cuda(embeds)
    x = embeds.float().contiguous()
    n, d = x.shape
    lab = labels.to(device=x.device, dtype=torch.int64).contiguous().view(-1)
    cls = torch.as_tensor([int(c) for c in classes], dtype=torch.int64, device=x.device)
    means = torch.empty((len(classes), d), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.load().knn_class_means(_ptr(x), _ptr(lab), n, d, _ptr(cls), len(classes), _ptr(means), None, _stream(x))
    L.check(rc, "knn_class_means")
    return means


def nearest_centroid_scores(train_embeds: torch.Tensor, train_labels: torch.Tensor, test_embeds: torch.Tensor,
                            classes: Sequence[int] = (0, 1)) -> np.ndarray:
    """anomaly/test_anomaly.py:31-48: distance of every test embedding to the nearest class centre of the training
    embeddings, divided by the largest such distance -> float64 [N] (the anomaly score fed to the AUROC)."""
    _require_cuda(test_embeds)
    cent = class_means(train_embeds, train_labels, classes)
    x = test_embeds.float().contiguous()
    n, d = x.shape
    out = torch.empty((n,), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.load().knn_centroid_min_dist(_ptr(x), _ptr(cent), n, d, len(classes), _ptr(out), _stream(x))
    L.check(rc, "knn_centroid_min_dist")
    dists = out.cpu().numpy()
    dists /= dists.max()
    return dists
