"""Retrieval metrics with the reference's names, signatures and return types, computed on the GPU.

Per-query work (label gather, relevance, hit counts, AP, majority vote, trapezoidal AP over a full ranking)
runs in the kernels of csrc/metrics.cu with the reference's own operation order in IEEE double; the Python
layer only shapes results into the reference's dicts and takes the final mean over the per-query array the
way the reference does (``np.mean`` / a sequential Python ``+=``), which keeps those numbers bit-identical.

Two entry styles per metric family:
  * the reference signature (dense ``dists`` / ``ranks`` matrices) for small problems, and
  * an embeddings-in / top-k-in variant that never materialises N x N.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import fullrank as FR
from .search import search_hamming, _ptr, _require_cuda, _stream, rank_rows, search

_NAN = float("nan")


# --------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# --------------------------------------------------------------------------------------------------
_CUTOFFS: Dict[Tuple[Tuple[int, ...], str], torch.Tensor] = {}


def _cutoffs_dev(cutoffs: Iterable[int], device) -> torch.Tensor:
    """int32 device tensor of cut-offs, cached per (values, device): the same few tuples ((1, 5, 10), 1..20) come back on
    every metric call and a pageable host->device copy costs more than the small-problem kernels themselves."""
    key = (tuple(int(c) for c in cutoffs), str(device))
    t = _CUTOFFS.get(key)
    if t is None:
        if len(_CUTOFFS) > 256:
            _CUTOFFS.clear()
        t = _CUTOFFS[key] = torch.as_tensor(list(key[0]), dtype=torch.int32, device=device)
    return t


def _dev_i64(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int64).contiguous()
    return torch.as_tensor(np.asarray(x), dtype=torch.int64, device=device).contiguous()


def relevance_single(indices: torch.Tensor, qlabels, glabels) -> Tuple[torch.Tensor, torch.Tensor]:
    """rel[i,j] = glabels[indices[i,j]] == qlabels[i]  and the retrieved labels.  -> (uint8 [Q,k], int64 [Q,k])"""
    _require_cuda(indices)
    idx = indices.contiguous().long()
    nq, k = idx.shape
    ql, gl = _dev_i64(qlabels, idx.device).view(-1), _dev_i64(glabels, idx.device).view(-1)
    rel = torch.empty((nq, k), dtype=torch.uint8, device=idx.device)
    lab = torch.empty((nq, k), dtype=torch.int64, device=idx.device)
    with torch.cuda.device(idx.device):
        rc = L.load().knn_relevance_single(_ptr(idx), nq, k, _ptr(ql), _ptr(gl), gl.numel(), _ptr(rel), _ptr(lab),
                                           _stream(idx))
    L.check(rc, "knn_relevance_single")
    return rel, lab


def pack_multihot(labels: torch.Tensor) -> torch.Tensor:
    """[N,C] multi-hot (C <= 64) -> int64 bitmask per row (bit c set iff labels[:,c] != 0)."""
    if labels.dim() != 2 or labels.shape[1] > 64:
        raise ValueError("multi-hot labels must be [N, C] with C <= 64")
    bits = (labels != 0).to(torch.int64)
    weights = (torch.ones((), dtype=torch.int64, device=labels.device) << torch.arange(labels.shape[1],
                                                                                   device=labels.device))
    return (bits * weights).sum(dim=1)


def relevance_multilabel(indices: torch.Tensor, qmask: torch.Tensor, gmask: torch.Tensor, jaccard_threshold: float,
                         arith: str = "fp32") -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (rel_jaccard uint8 [Q,k], rel_any uint8 [Q,k]); arith 'fp32' = torch/numpy fp32 tensors
    (train.py:462-466), 'fp64' = Python floats (evaluate_nih_zilliz.py:12-17)."""
    _require_cuda(indices)
    idx = indices.contiguous().long()
    nq, k = idx.shape
    qm, gm = _dev_i64(qmask, idx.device), _dev_i64(gmask, idx.device)
    rj = torch.empty((nq, k), dtype=torch.uint8, device=idx.device)
    ra = torch.empty((nq, k), dtype=torch.uint8, device=idx.device)
    with torch.cuda.device(idx.device):
        rc = L.load().knn_relevance_multilabel(_ptr(idx), nq, k, _ptr(qm), _ptr(gm), gm.numel(),
                                               float(jaccard_threshold), 0 if arith == "fp32" else 1, _ptr(rj),
                                               _ptr(ra), _stream(idx))
    L.check(rc, "knn_relevance_multilabel")
    return rj, ra


def ranked_stats(rel: torch.Tensor, kk: Optional[int] = None):
    """-> (hits int32 [Q], first int32 [Q] (1-based, 0 = none), ap_topk f64 [Q], prec_sum f64 [Q]) at cut-off kk."""
    _require_cuda(rel)
    rel = rel.contiguous()
    nq, k = rel.shape
    kk = k if kk is None else min(int(kk), k)
    dev = rel.device
    hits = torch.empty((nq,), dtype=torch.int32, device=dev)
    first = torch.empty((nq,), dtype=torch.int32, device=dev)
    ap = torch.empty((nq,), dtype=torch.float64, device=dev)
    ps = torch.empty((nq,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().knn_ranked_stats(_ptr(rel), nq, k, kk, _ptr(hits), _ptr(first), _ptr(ap), _ptr(ps), _stream(rel))
    L.check(rc, "knn_ranked_stats")
    return hits, first, ap, ps


def ranked_stats_multi(rel: torch.Tensor, cutoffs: Sequence[int], want: Sequence[str] = ("hits", "first", "ap", "ps")):
    """The statistics of :func:`ranked_stats` at every cut-off in ONE launch (``knn_ranked_stats_multi``):
    -> numpy (hits int32 [Q, nk], first int32 [Q] over the first max(cutoffs) items, ap_topk f64 [Q, nk],
    prec_sum f64 [Q, nk]); entries not named in ``want`` are neither computed nor read back (None).  Cut-offs beyond
    the list length are clamped to it."""
    _require_cuda(rel)
    rel = rel.contiguous()
    nq, k = rel.shape
    dev = rel.device
    cuts = [int(c) for c in cutoffs]
    res = {"hits": [], "first": [], "ap": [], "ps": []}
    for c0 in range(0, len(cuts), 16):
        part = cuts[c0:c0 + 16]
        nk = len(part)
        kks = _cutoffs_dev(part, dev)
        bufs = {"hits": torch.empty((nq, nk), dtype=torch.int32, device=dev) if "hits" in want else None,
                "first": torch.empty((nq,), dtype=torch.int32, device=dev) if "first" in want else None,
                "ap": torch.empty((nq, nk), dtype=torch.float64, device=dev) if "ap" in want else None,
                "ps": torch.empty((nq, nk), dtype=torch.float64, device=dev) if "ps" in want else None}
        with torch.cuda.device(dev):
            rc = L.load().knn_ranked_stats_multi(_ptr(rel), nq, k, _ptr(kks), nk, _ptr(bufs["hits"]), _ptr(bufs["first"]),
                                                 _ptr(bufs["ap"]), _ptr(bufs["ps"]), _stream(rel))
        L.check(rc, "knn_ranked_stats_multi")
        for key, buf in bufs.items():
            if buf is not None:
                res[key].append(buf)
    out = {}
    for key in ("hits", "ap", "ps"):
        out[key] = torch.cat(res[key], 1).cpu().numpy() if res[key] else None
    if res["first"]:
        f = torch.stack(res["first"], 1).cpu().numpy()
        out["first"] = np.where((f > 0).any(1), np.where(f > 0, f, np.iinfo(np.int32).max).min(1), 0).astype(np.int32)
    else:
        out["first"] = None
    return out["hits"], out["first"], out["ap"], out["ps"]


def majority_vote_multi(retrieved_labels: torch.Tensor, cutoffs: Sequence[int], tie: str = "first") -> np.ndarray:
    """Majority label at every cut-off in one launch + one read-back (``knn_majority_vote_multi``) -> int64 [Q, nk]."""
    _require_cuda(retrieved_labels)
    lab = retrieved_labels.contiguous().long()
    nq, k = lab.shape
    kks = _cutoffs_dev(cutoffs, lab.device)
    vote = torch.empty((nq, len(cutoffs)), dtype=torch.int64, device=lab.device)
    with torch.cuda.device(lab.device):
        rc = L.load().knn_majority_vote_multi(_ptr(lab), nq, k, _ptr(kks), len(cutoffs), 0 if tie == "first" else 1,
                                              _ptr(vote), _stream(lab))
    L.check(rc, "knn_majority_vote_multi")
    return vote.cpu().numpy()


def majority_vote_labels(retrieved_labels: torch.Tensor, kk: int, tie: str = "first") -> torch.Tensor:
    """Majority label of the first kk retrieved labels.  tie='first': collections.Counter.most_common
    (test.py:149-161); tie='smallest': torch.mode / np.unique+argmax (train_ath.py:208)."""
    _require_cuda(retrieved_labels)
    lab = retrieved_labels.contiguous().long()
    nq, k = lab.shape
    vote = torch.empty((nq,), dtype=torch.int64, device=lab.device)
    with torch.cuda.device(lab.device):
        rc = L.load().knn_majority_vote(_ptr(lab), nq, k, min(int(kk), k), 0 if tie == "first" else 1, _ptr(vote),
                                        _stream(lab))
    L.check(rc, "knn_majority_vote")
    return vote


def ap_sklearn(vals: torch.Tensor, rel: torch.Tensor) -> torch.Tensor:
    """sklearn.metrics.average_precision_score per ranked list (tied scores grouped).  NaN = no relevant item."""
    _require_cuda(vals, rel)
    vals, rel = vals.contiguous().float(), rel.contiguous()
    nq, k = vals.shape
    ap = torch.empty((nq,), dtype=torch.float64, device=vals.device)
    lib = L.load()
    with torch.cuda.device(vals.device):
        nbytes = lib.knn_ap_sklearn_workspace(nq, k)
        ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=vals.device)
        rc = lib.knn_ap_sklearn(_ptr(vals), _ptr(rel), nq, k, _ptr(ap), _ptr(ws), ws.numel(), _stream(vals))
    L.check(rc, "knn_ap_sklearn")
    return ap


def _seq_sum(x: np.ndarray, axis=None):
    """Left-to-right float64 accumulation: what a Python ``acc = acc + v`` loop does (test.py:134,141)."""
    if x.size == 0:
        return 0.0 if axis is None else np.zeros(x.shape[1:])
    return np.add.accumulate(x, axis=0 if axis is None else axis)[-1]


# --------------------------------------------------------------------------------------------------
# D1  retrieval_accuracy  (test.py:38-54; train.py:560; test_nonclip.py:31; eval_medsiglip.py:16)
# --------------------------------------------------------------------------------------------------
def recall_at_k_from_topk(indices: torch.Tensor, qlabels, glabels, topk: Sequence[int] = (1,)) -> List[torch.Tensor]:
    """R@K from retrieved indices [Q, >=max(topk)]: list of 0-d fp32 tensors, 100 * mean[any label match in top-k]."""
    _require_cuda(indices)
    idx = indices.contiguous().long()
    nq, k = idx.shape
    dev = idx.device
    ql, gl = _dev_i64(qlabels, dev).view(-1), _dev_i64(glabels, dev).view(-1)
    kks = _cutoffs_dev(topk, dev)
    counts = torch.empty((len(topk),), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().knn_recall_counts(_ptr(idx), nq, k, _ptr(ql), _ptr(gl), gl.numel(), _ptr(kks), len(topk),
                                        _ptr(counts), _stream(idx))
    L.check(rc, "knn_recall_counts")
    scale = torch.tensor(100.0 / nq, dtype=torch.float32)
    # fp32, as `correct_k.sum(dtype=float32) * (100.0 / batch)`; one read-back for every k
    return [c * scale for c in counts.cpu().to(torch.float32)]


def retrieval_accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)) -> List[torch.Tensor]:
    """Reference signature: ``output`` is the dense [N,N] score matrix (larger = more similar, diagonal already
    masked), ``target`` the labels.  Row-wise top-maxk with ties broken by ascending column."""
    _require_cuda(output)
    out = output.float()
    if out.stride(1) != 1:
        out = out.contiguous()
    nq, ng = out.shape
    lab = _dev_i64(target, out.device).view(-1)
    first = torch.empty((nq,), dtype=torch.int32, device=out.device)
    with torch.cuda.device(out.device):
        # R@K only asks whether the best-scoring row with the query's label sits in the first K: one launch gives its
        # rank (ties by ascending column, the order of rank_rows), one read-back settles every K
        rc = L.load().knn_first_relevant_rank(_ptr(out), out.stride(0), nq, ng, 1, _ptr(lab), _ptr(lab), _ptr(first),
                                              _stream(out))
    L.check(rc, "knn_first_relevant_rank")
    first_np = first.cpu().numpy()
    res = []
    for k in topk:
        correct_k = torch.tensor(float(((first_np > 0) & (first_np <= int(k))).sum()), dtype=torch.float32)
        res.append(correct_k * (100.0 / nq))            # fp32, as `correct_k.sum(dtype=float32) * (100.0 / batch)`
    return res


# --------------------------------------------------------------------------------------------------
# D2  compute_ap / compute_map  (test.py:58-146)
# --------------------------------------------------------------------------------------------------
def map_full(ranks_rowmajor: torch.Tensor, qlabels, glabels, kappas: Sequence[int] = ()):
    """Per-query trapezoidal AP + precision@kappas from a full ranking [Q, N] (best first).
    -> (ap f64 [Q], prs f64 [Q, len(kappas)], npos int32 [Q]) on the device."""
    _require_cuda(ranks_rowmajor)
    rk = ranks_rowmajor.contiguous().long()
    nq, ng = rk.shape
    dev = rk.device
    ql, gl = _dev_i64(qlabels, dev).view(-1), _dev_i64(glabels, dev).view(-1)
    kap = _cutoffs_dev(kappas, dev)
    ap = torch.empty((nq,), dtype=torch.float64, device=dev)
    prs = torch.empty((nq, max(len(kappas), 1)), dtype=torch.float64, device=dev)
    npos = torch.empty((nq,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().knn_map_full(_ptr(rk), nq, ng, _ptr(ql), _ptr(gl), _ptr(kap), len(kappas), _ptr(ap), _ptr(prs),
                                   _ptr(npos), _stream(rk))
    L.check(rc, "knn_map_full")
    return ap, prs[:, : len(kappas)], npos


def compute_map(ranks, gnd, kappas=[]):
    """Reference signature (test.py:95): ``ranks`` is [db_size, nq] (column i = ranking of query i, best first),
    ``gnd`` the labels.  -> (mAP, aps [nq], pr [len(kappas)], prs [nq, len(kappas)]) as float64 numpy."""
    dev = ranks.device if isinstance(ranks, torch.Tensor) and ranks.is_cuda else torch.device("cuda")
    rk = torch.as_tensor(ranks, device=dev).t().contiguous()
    gl = _dev_i64(gnd, dev)
    ap, prs, npos = map_full(rk, gl, gl, kappas)
    aps = ap.cpu().numpy()
    prs_np = prs.cpu().numpy().reshape(len(aps), len(kappas))
    valid = npos.cpu().numpy() > 0
    nq, nempty = len(aps), int((~valid).sum())
    mAP = _seq_sum(aps[valid]) / (nq - nempty)
    pr = _seq_sum(prs_np[valid], axis=0) / (nq - nempty) if len(kappas) else np.zeros(0)
    return float(mAP), aps, np.asarray(pr, dtype=np.float64), prs_np


def compute_map_from_embeddings(embeds: torch.Tensor, labels, kappas=(), metric: str = "l2", normalize: bool = False,
                                distributed: bool = False, precision: str = "fp32"):
    """``compute_map(argsort(dists, dim=0), labels, kappas)`` (test.py:1090-1091) straight from the embeddings, without
    the N x N ``dists`` / ``ranks`` matrices: ``dists = -cdist(e, e)`` (``metric="l2"``) or ``e @ e.T`` with the
    diagonal at -inf, the trapezoidal AP and mP@k of test.py:58-146 over the full ranking.  The query counts among its
    own positives and ranks last.  Same return value as :func:`compute_map`."""
    _require_cuda(embeds)
    lab = _dev_i64(labels, embeds.device).view(-1)
    st = FR.full_ranking_stats(embeds, embeds, FR.REL_SINGLE, lab, lab, metric=metric, normalize=normalize,
                               self_mode="exclude", drop_self=True, kappas=kappas, self_last_positive=True,
                               outputs=("ap_trapz", "prs", "nres"), distributed=distributed, precision=precision)
    aps = st["ap_trapz"].cpu().numpy()
    prs_np = st["prs"].cpu().numpy().reshape(len(aps), len(kappas))
    valid = st["nres"].cpu().numpy() > 0
    nq, nempty = len(aps), int((~valid).sum())
    mAP = _seq_sum(aps[valid]) / (nq - nempty)
    pr = _seq_sum(prs_np[valid], axis=0) / (nq - nempty) if len(kappas) else np.zeros(0)
    return float(mAP), aps, np.asarray(pr, dtype=np.float64), prs_np


# --------------------------------------------------------------------------------------------------
# D3  compute_classification_metrics  (test.py:149-223)
# --------------------------------------------------------------------------------------------------
def _prf_from_predictions(true: np.ndarray, pred: np.ndarray) -> Dict[str, float]:
    """macro / weighted precision, recall, F1 (zero_division=0) and accuracy from the confusion counts, with
    sklearn's formulas (precision_recall_fscore_support): host math on a C x C matrix built by ONE bincount."""
    labels, codes = np.unique(np.concatenate([true, pred]), return_inverse=True)
    c = len(labels)
    tc, pc = codes[: len(true)], codes[len(true):]
    conf = np.bincount(tc * c + pc, minlength=c * c).reshape(c, c).astype(np.float64)   # [true, predicted]
    tp, pred_sum, true_sum = np.diag(conf).copy(), conf.sum(axis=0), conf.sum(axis=1)

    def div(a, b):
        out = np.zeros_like(a)
        np.divide(a, b, out=out, where=b != 0)
        return out

    precision, recall = div(tp, pred_sum), div(tp, true_sum)
    f1 = div(2.0 * tp, true_sum + pred_sum)
    w = true_sum

    def avg(x, weights=None):
        return float(np.average(x, weights=weights))

    return {
        "precision_macro": avg(precision) * 100.0, "recall_macro": avg(recall) * 100.0, "f1_macro": avg(f1) * 100.0,
        "precision_weighted": avg(precision, w) * 100.0, "recall_weighted": avg(recall, w) * 100.0,
        "f1_weighted": avg(f1, w) * 100.0,
        "accuracy": float(tp.sum() / len(true)) * 100.0,
    }


def classification_metrics_from_topk(indices: torch.Tensor, qlabels, glabels, k_values=(1, 5, 10, 15, 20),
                                     tie: str = "first") -> Dict[int, Dict[str, float]]:
    _, lab = relevance_single(indices, qlabels, glabels)
    true = _dev_i64(qlabels, indices.device).view(-1).cpu().numpy()
    votes = majority_vote_multi(lab, list(k_values), tie)          # every k in one launch, one read-back
    return {k: _prf_from_predictions(true, votes[:, j]) for j, k in enumerate(k_values)}


def compute_classification_metrics(labels: torch.Tensor, dists: torch.Tensor, k_values=[1, 5, 10, 15, 20]):
    """Reference signature (test.py:164): dense ``dists`` (higher = more similar); ranks COLUMNS
    (``argsort(dists, dim=0)``), ties by ascending row."""
    _require_cuda(dists)
    ranks = rank_rows(dists.t().contiguous(), largest_first=True)[:, : max(k_values)].contiguous()
    return classification_metrics_from_topk(ranks, labels, labels, k_values)


# --------------------------------------------------------------------------------------------------
# D10 compute_metrics  (test_ath.py:90-172): true query x gallery retrieval
# --------------------------------------------------------------------------------------------------
def compute_metrics(query_codes, query_labels, gallery_codes, gallery_labels, query_logits=None,
                    topk_values=(1, 5, 10), binary_codes: bool = False, precision: str = "fp32"):
    """mHR, mAP@k, mRR, mP@k, R@k and majority-vote accuracy per cut-off; ranking by ascending L2 distance, or by
    Hamming distance over 0/1 codes with ``binary_codes`` (``pairwise_distance``, test_ath.py:80-87)."""
    _require_cuda(query_codes, gallery_codes)
    kmax = max(topk_values)
    if binary_codes:
        _, idx = search_hamming(query_codes, gallery_codes, kmax)
    else:
        _, idx = search(query_codes.float(), gallery_codes.float(), kmax, "l2", precision=precision)
    dev = idx.device
    ql, gl = _dev_i64(query_labels, dev).view(-1), _dev_i64(gallery_labels, dev).view(-1)
    rel, lab = relevance_single(idx, ql, gl)
    # total relevant items per query (integer counting, exact)
    uniq, counts = torch.unique(gl, return_counts=True)
    pos = torch.searchsorted(uniq, ql).clamp(max=uniq.numel() - 1)
    total_rel = torch.where(uniq[pos] == ql, counts[pos], torch.zeros_like(counts[pos])).cpu().numpy()
    retrieval = {}
    cuts = [int(t) for t in topk_values]
    hits_all, first_np, ap_all, _ = ranked_stats_multi(rel, cuts, want=("hits", "first", "ap"))    # every cut-off: one launch, one read-back
    votes = majority_vote_multi(lab, cuts, "first")
    ql_np = ql.cpu().numpy()
    for j, topk in enumerate(topk_values):
        hits_np = hits_all[:, j]
        first_k = np.where((first_np > 0) & (first_np <= cuts[j]), first_np, 0)   # first hit inside this cut-off
        with np.errstate(divide="ignore", invalid="ignore"):
            rr = np.where(first_k > 0, 1.0 / np.maximum(first_k, 1), 0.0)
            rec = np.where(total_rel > 0, hits_np / np.maximum(total_rel, 1), 0.0)
        retrieval[topk] = {
            "mhr": float(np.mean((hits_np > 0).astype(np.float64))),
            "map": float(np.mean(ap_all[:, j])),
            "mrr": float(np.mean(rr)),
            "mp@k": float(np.mean(hits_np / topk)),
            "r@k": float(np.mean(rec)),
            "majority_acc": float(np.mean((votes[:, j] == ql_np).astype(np.float64))),
        }
    classification_acc = None
    if query_logits is not None:
        classification_acc = query_logits.argmax(dim=1).eq(ql.to(query_logits.device)).float().mean().item()
    return {"classification_acc": classification_acc, "retrieval": retrieval}


# --------------------------------------------------------------------------------------------------
# D9  evaluate_results  (evaluate_nih_zilliz.py:34-64) over top-k hit lists, and D5 hit-rate (test.py:1016-1056)
# --------------------------------------------------------------------------------------------------
def evaluate_results_from_topk(vals: torch.Tensor, indices: torch.Tensor, qlabels_multihot: torch.Tensor,
                               glabels_multihot: torch.Tensor, jaccard_threshold: float = 0.4,
                               ks: Iterable[int] = (1, 5, 10, 20, 50)) -> Dict[str, float]:
    """``evaluate_results`` on device-resident top-k lists instead of the JSON hit lists."""
    qm, gm = pack_multihot(qlabels_multihot), pack_multihot(glabels_multihot)
    rel, _ = relevance_multilabel(indices, qm, gm, jaccard_threshold, arith="fp64")
    nhits = rel.shape[1]
    total_pos, _, _, _ = ranked_stats(rel)
    total_np = total_pos.cpu().numpy().astype(np.int64)
    ap = ap_sklearn(vals, rel).cpu().numpy()
    aps = ap[total_np > 0]
    metrics = {
        "mAP": float(np.mean(aps) * 100.0) if len(aps) else 0.0,
        "num_queries": float(rel.shape[0]),
        "num_valid_ap_queries": float(len(aps)),
    }
    ks = list(ks)
    hits_all = ranked_stats_multi(rel, [min(int(k), nhits) for k in ks], want=("hits",))[0] if ks else None
    for j, k in enumerate(ks):
        kk = min(int(k), nhits)
        hk = hits_all[:, j].astype(np.float64)
        # precision_at_k = float(np.mean(relevances[:k])) over float64 0/1 entries: sum is exact, one division
        p = hk / kk
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.where(total_np > 0, hk / np.maximum(total_np, 1), 0.0)
        metrics[f"P@{k}"] = float(np.mean(p) * 100.0) if len(p) else 0.0
        metrics[f"R@{k}"] = float(np.mean(r) * 100.0) if len(r) else 0.0
    return metrics


def hits_to_arrays(items: Sequence[Mapping]):
    """The hits JSON of query_nih_zilliz.py:58-72 (``formats.nih_query_results``) as arrays: scores fp32 [Q, K], hit
    positions int64 [Q, K] (= rows of the label table below), query multi-hots fp32 [Q, C] and the multi-hots of all
    hits fp32 [Q * K, C].  Every query must carry the same number of hits (what ``search_collection(top_k)`` returns)."""
    nq = len(items)
    k = len(items[0]["results"]) if nq else 0
    for it in items:
        if len(it["results"]) != k:
            raise ValueError(f"evaluate_results on the device needs the same number of hits for every query "
                             f"(got {len(it['results'])} and {k})")
    qlab = np.asarray([it["query_label_vector"] for it in items], dtype=np.float32).reshape(nq, -1)
    vals = np.asarray([[h["score"] for h in it["results"]] for it in items], dtype=np.float32).reshape(nq, k)
    glab = np.asarray([h["label_vector"] for it in items for h in it["results"]], dtype=np.float32).reshape(nq * k, -1)
    idx = np.arange(nq * k, dtype=np.int64).reshape(nq, k)
    return vals, idx, qlab, glab


def evaluate_results(items: Sequence[Mapping], jaccard_threshold: float, ks: Sequence[int], device=None) -> Dict[str, float]:
    """Reference signature (evaluate_nih_zilliz.py:34-64): ``items`` is the hits JSON -- one entry per query with its
    ``query_label_vector`` and ``results`` (each hit: ``score``, ``label_vector``).  Same metric dictionary; the work runs
    in :func:`evaluate_results_from_topk` on the device."""
    ks = [int(k) for k in ks]
    if not items or not len(items[0]["results"]):
        for it in items:
            if len(it["results"]):
                raise ValueError("evaluate_results on the device needs the same number of hits for every query")
        # no query, or no hit for any query: every list the reference averages is empty or all zero
        out = {"mAP": 0.0, "num_queries": float(len(items)), "num_valid_ap_queries": 0.0}
        for k in ks:
            out[f"P@{k}"] = 0.0
            out[f"R@{k}"] = 0.0
        return out
    vals, idx, qlab, glab = hits_to_arrays(items)
    dev = device or torch.device("cuda", torch.cuda.current_device())
    to = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    return evaluate_results_from_topk(to(vals), to(idx), to(qlab), to(glab), jaccard_threshold, ks)


def multilabel_hit_rate_from_topk(indices: torch.Tensor, qlabels_multihot: torch.Tensor,
                                  glabels_multihot: torch.Tensor, k_values=(1, 5, 10, 15, 20)):
    """Precision@K (share of the top-K sharing >= 1 label with the query) and Recall@K = hit-rate
    (test.py:1031-1056).  -> {k: (precision_percent, recall_percent)}"""
    qm, gm = pack_multihot(qlabels_multihot), pack_multihot(glabels_multihot)
    _, rel_any = relevance_multilabel(indices, qm, gm, 0.0)
    nq = rel_any.shape[0]
    out = {}
    hits_all = ranked_stats_multi(rel_any, list(k_values), want=("hits",))[0]
    for j, k in enumerate(k_values):
        hk = hits_all[:, j]
        total_precision = _seq_sum(hk.astype(np.float64) / k)   # total_precision += num_matches / k
        total_recall = int((hk > 0).sum())
        out[k] = (float(total_precision / nq * 100), float(total_recall / nq * 100))
    return out


# --------------------------------------------------------------------------------------------------
# D6 / D11 / D13: embeddings-in single-label evaluation (train.py:399-441; fusion_eval/metrics.py:41-94)
# --------------------------------------------------------------------------------------------------
def _compute_single_label_retrieval_metrics(embeds: torch.Tensor, labels: torch.Tensor, topk=(1, 5, 10),
                                            distributed: bool = False, precision: str = "fp32"):
    """train.py:399-441: cosine self-retrieval, standard AP over the full ranking / (#same-label - 1), R@K."""
    if len(labels) <= 1:
        return {"mAP": 0.0, **{f"R@{k}": 0.0 for k in topk}}
    _require_cuda(embeds)
    labels = _dev_i64(labels, embeds.device).view(-1)
    st = FR.full_ranking_stats(embeds, embeds, FR.REL_SINGLE, labels, labels, metric="cosine", normalize=True,
                               self_mode="exclude", drop_self=True, outputs=("prec_sum", "first"),
                               distributed=distributed, precision=precision)
    hits_np, ps = st["npos"].cpu().numpy(), st["prec_sum"].cpu().numpy()
    aps = np.where(hits_np > 0, ps / np.maximum(hits_np, 1), 0.0)   # relevant_counts == hits over the full ranking
    metrics = {"mAP": float(np.mean(aps) * 100.0)}
    first_np = st["first"].cpu().numpy()
    for k in topk:
        actual_k = min(k, len(labels) - 1)
        metrics[f"R@{k}"] = float(np.mean(((first_np > 0) & (first_np <= actual_k)).astype(np.float32))) * 100.0
    return metrics


def retrieval_metrics_from_ranking(idx: torch.Tensor, labels: Sequence,
                                   k_values: Iterable[int] = (1, 5, 10)) -> Dict[str, float]:
    """fusion_eval/metrics.py:41-94 from a FULL self-retrieval ranking ``idx`` [N, N-1] (self removed): standard AP
    over the ranking / (#same-label - 1), ``mP@k = hits/k``, ``R@k`` any-hit, 0.0 for queries without relevant items.
    Image paths are assumed unique (the reference's aligned embedding sets are), so dropping the self entry is the
    reference's by-path filter."""
    _, inv = np.unique(np.asarray(labels), return_inverse=True)
    lab = torch.as_tensor(inv, dtype=torch.int64, device=idx.device)
    k_values = sorted(set(int(k) for k in k_values))
    rel, _ = relevance_single(idx, lab, lab)
    hits, _, _, prec_sum = ranked_stats(rel)
    hits_np, ps = hits.cpu().numpy(), prec_sum.cpu().numpy()
    aps = np.where(hits_np > 0, ps / np.maximum(hits_np, 1), 0.0)
    metrics: Dict[str, float] = {"num_samples": float(len(inv)), "mAP": float(np.mean(aps) * 100.0)}
    hits_all = ranked_stats_multi(rel, k_values, want=("hits",))[0] if k_values else None
    for j, k in enumerate(k_values):
        hk = np.where(hits_np > 0, hits_all[:, j], 0)
        metrics[f"mP@{k}"] = float(np.mean(hk / k) * 100.0)
        metrics[f"R@{k}"] = float(np.mean((hk > 0).astype(np.float64)) * 100.0)
    return metrics


def evaluate_retrieval_metrics(embeddings, labels: Sequence, image_paths: Optional[Sequence] = None,
                               k_values: Iterable[int] = (1, 5, 10)) -> Dict[str, float]:
    """fusion_eval/metrics.py:26-94 (labels may be strings): cosine self-retrieval, standard AP over the full ranking
    divided by (#same-label - 1), ``mP@k = hits/k``, ``R@k`` any-hit, 0.0 for queries without relevant items.  Every
    gallery row that shares the query's image path is left out of the ranking (``ranked_indices[image_paths[...] !=
    image_paths[q]]``, metrics.py:67) -- with unique paths that is the query itself.  No N x N matrix is built."""
    emb = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)) if not isinstance(embeddings, torch.Tensor) \
        else embeddings
    emb = emb.cuda() if not emb.is_cuda else emb
    n = emb.shape[0]
    if len(labels) != n or (image_paths is not None and len(image_paths) != n):
        raise ValueError("Labels, image_paths, and similarity matrix must have matching sizes")
    _, inv = np.unique(np.asarray(labels), return_inverse=True)
    inv = inv.reshape(-1)
    lab = torch.as_tensor(inv, dtype=torch.int64, device=emb.device)
    grp = None
    if image_paths is not None:
        upaths, pinv = np.unique(np.asarray(image_paths), return_inverse=True)
        if len(upaths) != n:   # shared paths: drop by group
            grp = torch.as_tensor(pinv.reshape(-1), dtype=torch.int64, device=emb.device)
    k_values = sorted(set(int(k) for k in k_values))
    relevant_count = np.bincount(inv)[inv] - 1            # np.sum(labels == labels[q]) - 1
    metrics: Dict[str, float] = {"num_samples": float(n)}
    aps = np.zeros(n)
    hits_k = {}
    for c in range(0, max(len(k_values), 1), 8):          # the kernel takes up to 8 cut-offs per pass
        ks = k_values[c:c + 8]
        st = FR.full_ranking_stats(emb, emb, FR.REL_SINGLE, lab, lab, metric="cosine", normalize=True,
                                   self_mode="exclude", drop_self=True, q_group=grp, g_group=grp, kappas=ks,
                                   outputs=("prec_sum", "hits_at") if c == 0 else ("hits_at",))
        if c == 0:
            hits_np, ps = st["npos"].cpu().numpy(), st["prec_sum"].cpu().numpy()
            ok = (relevant_count > 0) & (hits_np > 0)
            aps = np.where(ok, ps / np.maximum(relevant_count, 1), 0.0)
        ha = st["hits_at"].cpu().numpy()
        for j, k in enumerate(ks):
            hits_k[k] = np.where(relevant_count > 0, ha[:, j], 0)
    metrics["mAP"] = float(np.mean(aps) * 100.0)
    for k in k_values:
        metrics[f"mP@{k}"] = float(np.mean(hits_k[k] / k) * 100.0)
        metrics[f"R@{k}"] = float(np.mean((hits_k[k] > 0).astype(np.float64)) * 100.0)
    return metrics


def is_retrieval_correct(query_label, results, config=None) -> bool:
    """retrieval_analysis/evaluator.py:18-26 (``is_retrieval_correct(query_label, results, config)``): lives in
    :mod:`analysis` next to ``CorrectnessConfig`` and the four-way grouping; re-exported here with the metric rows."""
    from .analysis import CorrectnessConfig, is_retrieval_correct as _impl

    return _impl(query_label, results, config if config is not None else CorrectnessConfig())


# --------------------------------------------------------------------------------------------------
# D4 / D7 / D8: multilabel AP over the full ranking
# --------------------------------------------------------------------------------------------------
def compute_map_multilabel_from_embeddings(embeds: torch.Tensor, labels_multihot: torch.Tensor,
                                           threshold: float = 0.5, distributed: bool = False, precision: str = "fp32") -> float:
    """test.py:941-985 on cosine self-retrieval: rank-by-rank AP, relevance = Jaccard > threshold, self removed,
    queries without relevant items skipped."""
    _require_cuda(embeds)
    m = pack_multihot(labels_multihot.to(embeds.device))
    st = FR.full_ranking_stats(embeds, embeds, FR.REL_JACCARD_F32, m, m, metric="cosine", normalize=True,
                               self_mode="exclude", drop_self=True, jaccard_threshold=float(threshold),
                               outputs=("prec_sum",), distributed=distributed, precision=precision)
    hits_np, ps = st["npos"].cpu().numpy(), st["prec_sum"].cpu().numpy()
    aps = (ps / np.maximum(hits_np, 1))[hits_np > 0]
    return float(np.mean(aps)) if len(aps) else 0


def compute_map_multilabel(dists: torch.Tensor, labels: torch.Tensor, threshold: float = 0.5) -> float:
    """Reference signature (test.py:941): dense ``dists`` (larger = closer, diagonal already -inf), multi-hot
    ``labels``; ranks COLUMNS (``np.argsort(-dists, axis=0)``, ties by ascending row), relevance = Jaccard > threshold
    with the query itself removed, rank-by-rank AP normalised by the number of relevant items, queries without any
    skipped.  For small N (the dense matrix exists); the embeddings-in form above never builds it."""
    _require_cuda(dists)
    n = dists.shape[0]
    m = pack_multihot(labels.to(dists.device))
    ranks = rank_rows(dists.t().contiguous(), largest_first=True)
    rel, _ = relevance_multilabel(ranks, m, m, threshold, arith="fp32")
    rel[ranks == torch.arange(n, device=ranks.device)[:, None]] = 0       # binary_relevance[i] = 0
    hits, _, ap, _ = ranked_stats(rel)
    aps = ap.cpu().numpy()[hits.cpu().numpy() > 0]
    return float(np.mean(aps)) if len(aps) else 0


def _compute_multilabel_retrieval_metrics(embeds: torch.Tensor, labels: torch.Tensor, topk=(1, 5, 10),
                                          relevance_threshold: float = 0.4, distributed: bool = False,
                                          precision: str = "fp32"):
    """train.py:444-487: sklearn AP (tied scores grouped) with self masked out, R@K = any relevant in top-k."""
    if len(labels) <= 1:
        return {"mAP": 0.0, **{f"R@{k}": 0.0 for k in topk}}
    _require_cuda(embeds)
    m = pack_multihot(labels.to(embeds.device))
    st = FR.full_ranking_stats(embeds, embeds, FR.REL_JACCARD_F32, m, m, metric="cosine", normalize=True,
                               self_mode="exclude", drop_self=True, jaccard_threshold=float(relevance_threshold),
                               sklearn_ap=True, outputs=("first",), distributed=distributed, precision=precision)
    hits_np, first_np = st["npos"].cpu().numpy(), st["first"].cpu().numpy()
    aps = st["ap_sklearn"].cpu().numpy()[hits_np > 0]
    metrics = {"mAP": float(np.mean(aps) * 100.0) if len(aps) > 0 else 0.0}
    for k in topk:
        actual_k = min(k, len(labels) - 1)
        metrics[f"R@{k}"] = float(np.mean(((first_np > 0) & (first_np <= actual_k)).astype(np.float64)) * 100.0)
    return metrics


def evaluate_map_embeddings(embeddings: torch.Tensor, labels: torch.Tensor, jaccard_threshold: float = 0.4,
                            distributed: bool = False, precision: str = "fp32") -> float:
    """nih_multilabel_training.py:66-99 on given embeddings: self is KEPT with similarity -1 and counts as a
    relevant item (J(self, self) = 1 > threshold) ranked wherever -1 falls.  ``distributed=True`` inside a
    ``torch.distributed`` job shards the QUERIES over the ranks (every rank passes the same embeddings, as after the
    reference's ``dist.all_gather`` of train.py:604-609) and returns the same value on every rank."""
    _require_cuda(embeddings)
    m = pack_multihot(labels.to(embeddings.device))
    st = FR.full_ranking_stats(embeddings, embeddings, FR.REL_JACCARD_F32, m, m, metric="cosine", normalize=True,
                               self_mode="minus1", drop_self=False, jaccard_threshold=float(jaccard_threshold),
                               sklearn_ap=True, outputs=(), distributed=distributed, precision=precision)
    hits = st["npos"].cpu().numpy()
    aps = st["ap_sklearn"].cpu().numpy()[hits > 0]
    if len(aps) == 0:
        return 0.0
    return float(np.mean(aps) * 100.0)


# --------------------------------------------------------------------------------------------------
# D12 evaluate_retrieval  (evaluate_medsiglip.py:142-163)
# --------------------------------------------------------------------------------------------------
def evaluate_retrieval(image_features: torch.Tensor, labels, topk_values) -> Dict[str, float]:
    """Self-retrieval over (already normalised) features: ``sim = f @ f.T``, ``fill_diagonal_(-1.0)``, row-wise
    ranking; R@k any-hit, and majority vote with ``np.unique`` + ``argmax`` (ties -> the SMALLEST label) scored by
    accuracy and macro F1, all in percent.  Same keys as the reference: ``r_at_{k}``, ``majority_accuracy_at_{k}``,
    ``majority_macro_f1_at_{k}``."""
    _require_cuda(image_features)
    kmax = max(int(k) for k in topk_values)
    _, idx = search(image_features, image_features, kmax, "ip", self_mode="minus1")
    lab_dev = _dev_i64(labels, idx.device).view(-1)
    rel, lab = relevance_single(idx, lab_dev, lab_dev)
    true = lab_dev.cpu().numpy()
    results: Dict[str, float] = {}
    cuts = [int(k) for k in topk_values]
    hits_all = ranked_stats_multi(rel, cuts, want=("hits",))[0]
    votes = majority_vote_multi(lab, cuts, "smallest")
    for j, k in enumerate(topk_values):
        hits = hits_all[:, j]
        prf = _prf_from_predictions(true, votes[:, j])
        results[f"r_at_{k}"] = float(np.mean(hits > 0) * 100.0)
        results[f"majority_accuracy_at_{k}"] = prf["accuracy"]
        results[f"majority_macro_f1_at_{k}"] = prf["f1_macro"]
    return results


# --------------------------------------------------------------------------------------------------
# D10' compute_retrieval_metrics  (train_ath.py:171-218): the training-time variant of compute_metrics
# --------------------------------------------------------------------------------------------------
def compute_retrieval_metrics(query_codes, query_labels, gallery_codes, gallery_labels, topk_values,
                              binary_codes: bool, precision: str = "fp32") -> Dict[int, Dict[str, float]]:
    """mHR, mAP@k (precision sum / positives in the top-k), mRR and majority accuracy with ``torch.mode``
    (ties -> the smallest label), ranking by ascending L2 or Hamming distance."""
    _require_cuda(query_codes, gallery_codes)
    kmax = max(int(k) for k in topk_values)
    if binary_codes:
        _, idx = search_hamming(query_codes, gallery_codes, kmax)
    else:
        _, idx = search(query_codes.float(), gallery_codes.float(), kmax, "l2", precision=precision)
    ql, gl = _dev_i64(query_labels, idx.device).view(-1), _dev_i64(gallery_labels, idx.device).view(-1)
    rel, lab = relevance_single(idx, ql, gl)
    results = {}
    cuts = [int(t) for t in topk_values]
    hits_all, first_np, ap_all, _ = ranked_stats_multi(rel, cuts, want=("hits", "first", "ap"))
    votes = majority_vote_multi(lab, cuts, "smallest")
    ql_np = ql.cpu().numpy()
    for j, topk in enumerate(topk_values):
        hits_np = hits_all[:, j]
        first_k = np.where((first_np > 0) & (first_np <= cuts[j]), first_np, 0)
        rr = np.where(first_k > 0, 1.0 / np.maximum(first_k, 1), 0.0)
        results[topk] = {
            "mhr": float(np.mean((hits_np > 0).astype(np.float64))),
            "map": float(np.mean(ap_all[:, j])),
            "mrr": float(np.mean(rr)),
            "majority_acc": float(np.mean((votes[:, j] == ql_np).astype(np.float64))),
        }
    return results


# --------------------------------------------------------------------------------------------------
# D14 retrieval_accuracy_from_ranks / compute_classification_metrics_from_ranks  (ChestMIR/chestmir_eval.py:191-272)
# --------------------------------------------------------------------------------------------------
def _labels_to_codes(labels) -> Tuple[np.ndarray, np.ndarray]:
    """Arbitrary (string / object) labels -> (int64 codes, the distinct labels); equality is all the metrics need."""
    arr = np.asarray(labels)
    uniq, inv = np.unique(arr.astype(str) if arr.dtype == object else arr, return_inverse=True)
    return inv.astype(np.int64).reshape(-1), uniq


def _ranks_topk_device(ranks: np.ndarray, kmax: int, device) -> torch.Tensor:
    """The reference's column layout ``ranks[:k, i]`` = top-k of query i  ->  int64 [N, kmax] on the device."""
    r = np.ascontiguousarray(np.asarray(ranks)[:kmax, :].T).astype(np.int64)
    return torch.from_numpy(r).to(device)


def retrieval_accuracy_from_ranks(ranks: np.ndarray, labels, topk, device=None) -> np.ndarray:
    """R@K in percent (float64 array) from a full column-wise ranking matrix; any label match in ``ranks[:k, i]``."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    codes, _ = _labels_to_codes(labels)
    n = len(codes)
    kmax = min(max(int(k) for k in topk), np.asarray(ranks).shape[0])
    idx = _ranks_topk_device(ranks, kmax, device)
    lab = torch.from_numpy(codes).to(device)
    rel, _ = relevance_single(idx, lab, lab)
    hits_all = ranked_stats_multi(rel, [min(int(k), kmax) for k in topk], want=("hits",))[0]
    out = [(int((hits_all[:, j] > 0).sum()) * 100.0) / max(1, n) for j in range(len(topk))]
    return np.array(out, dtype=np.float64)


def compute_classification_metrics_from_ranks(labels, ranks: np.ndarray, k_values, device=None):
    """Majority vote (``Counter.most_common``: first label met in rank order wins a tie) + the reference's hand-rolled
    per-class precision / recall / F1 (``f1 = 2pr / (p + r)``), macro and support-weighted, in percent."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    codes, _ = _labels_to_codes(labels)
    kmax = min(max(int(k) for k in k_values), np.asarray(ranks).shape[0])
    idx = _ranks_topk_device(ranks, kmax, device)
    lab = torch.from_numpy(codes).to(device)
    _, retrieved = relevance_single(idx, lab, lab)
    results = {}
    votes = majority_vote_multi(retrieved, [min(int(k), kmax) for k in k_values], "first")
    for j, k in enumerate(k_values):
        y_pred = votes[:, j]
        y_true = codes
        classes = np.unique(np.concatenate([y_true, y_pred], axis=0))
        per_p, per_r, per_f, supports = [], [], [], []
        for c in classes:
            tp = int(np.sum((y_true == c) & (y_pred == c)))
            fp = int(np.sum((y_true != c) & (y_pred == c)))
            fn = int(np.sum((y_true == c) & (y_pred != c)))
            p = tp / (tp + fp) if (tp + fp) > 0 else 0.0
            r = tp / (tp + fn) if (tp + fn) > 0 else 0.0
            per_p.append(p)
            per_r.append(r)
            per_f.append((2.0 * p * r / (p + r)) if (p + r) > 0 else 0.0)
            supports.append(int(np.sum(y_true == c)))
        sup = np.asarray(supports, dtype=np.float64)
        wsum = float(np.sum(sup))
        weights = sup / (wsum if wsum > 0 else 1.0)
        results[k] = {
            "accuracy": float(np.mean(y_true == y_pred)) * 100.0,
            "precision_macro": (float(np.mean(per_p)) if per_p else 0.0) * 100.0,
            "recall_macro": (float(np.mean(per_r)) if per_r else 0.0) * 100.0,
            "f1_macro": (float(np.mean(per_f)) if per_f else 0.0) * 100.0,
            "precision_weighted": (float(np.sum(np.asarray(per_p) * weights)) if per_p else 0.0) * 100.0,
            "recall_weighted": (float(np.sum(np.asarray(per_r) * weights)) if per_r else 0.0) * 100.0,
            "f1_weighted": (float(np.sum(np.asarray(per_f) * weights)) if per_f else 0.0) * 100.0,
        }
    return results


# --------------------------------------------------------------------------------------------------
# D9 helpers  (evaluate_nih_zilliz.py:12-31): the per-item functions behind evaluate_results
# --------------------------------------------------------------------------------------------------
def jaccard_score(query_label, gallery_label) -> float:
    """``intersection / (union + 1e-8)`` of two multi-hot vectors: float32 products / clipped sums, Python-float ratio
    (host scalar math on two label vectors; the batched form is relevance_multilabel(arith="fp64"))."""
    q = np.asarray(query_label, dtype=np.float32)
    g = np.asarray(gallery_label, dtype=np.float32)
    return float((q * g).sum()) / (float(np.clip(q + g, 0.0, 1.0).sum()) + 1e-8)


def precision_at_k(binary_relevance, k: int) -> float:
    if not len(binary_relevance):
        return 0.0
    k = min(k, len(binary_relevance))
    return float(np.mean(binary_relevance[:k]))


def recall_at_k(binary_relevance, total_positives: int, k: int) -> float:
    if total_positives <= 0:
        return 0.0
    k = min(k, len(binary_relevance))
    return float(np.sum(binary_relevance[:k]) / total_positives)


# --------------------------------------------------------------------------------------------------
# The evaluation body of test.py:1077-1126 / test_nonclip.py:150-172 / eval_medsiglip.py:238-260 after the model forward
# --------------------------------------------------------------------------------------------------
def evaluate_embeddings(embeds: torch.Tensor, labels, metric: str = "l2", kappas=(1, 5, 10),
                        k_values=(1, 5, 10, 15, 20), save_path: Optional[str] = None):
    """What ``evaluate(model, loader, device, args)`` computes once the embeddings exist: ``dists = -cdist(e, e)``
    (``metric="l2"``, test.py:1080) or ``e @ e.T`` (``"cosine"`` / ``"ip"``, test.py:1006, eval_medsiglip.py:238) with the
    diagonal at -inf; R@K; the trapezoidal mAP and mP@K over the column-wise full ranking; majority-vote classification
    metrics; optionally the ``np.savez`` bundle of test.py:1122-1126.
    -> {"acc": float32 [len(kappas)], "mAP": float, "pr": float64 [len(kappas)], "classification": {k: {...}}}.
    Nothing of size N x N is built (unless ``save_path`` asks for the bundle, which stores ``dists``): R@K and the vote
    come from one fused top-k search, the full-ranking mAP from the ranks of the positives (fullrank.py)."""
    _require_cuda(embeds)
    kappas, k_values = list(kappas), list(k_values)
    lab = _dev_i64(labels, embeds.device).view(-1)
    n = embeds.shape[0]
    # R@K and the majority vote read the first max(k) neighbours: one fused top-k search (rows = columns: the engine's
    # scores are symmetric, so the column ranking of test.py:1090 is the row ranking)
    kmax = max(1, min(max(kappas + k_values), n - 1))
    _, idx = search(embeds, embeds, kmax, metric, exclude_self=True)
    acc = torch.stack(recall_at_k_from_topk(idx, lab, lab, kappas)).numpy()
    # trapezoidal mAP / mP@K over the FULL ranking from the ranks of the positives -- no N x N matrix
    mAP, _, pr, _ = compute_map_from_embeddings(embeds, lab, kappas, metric=metric)
    classification = classification_metrics_from_topk(idx, lab, lab, k_values)
    out = {"acc": acc, "mAP": mAP, "pr": pr, "classification": classification}
    if save_path is not None:   # the bundle of test.py:1122-1126 stores the dense `dists`: only built when asked for
        from .formats import save_evaluation_npz
        from .search import scores_dense

        if metric == "l2":
            dists = -scores_dense(embeds, embeds, "l2", self_mode="exclude")      # -cdist, diagonal -inf
        else:
            dists = scores_dense(embeds, embeds, metric, self_mode="exclude")
        save_evaluation_npz(save_path, embeds, lab, kappas, acc, mAP, pr, classification, dists=dists)
    return out


def evaluate_multilabel_embeddings(embeds: torch.Tensor, labels_multihot: torch.Tensor, thresholds=(0.25, 0.5),
                                   k_values=(1, 5, 10, 15, 20), save_path: Optional[str] = None):
    """What ``evaluate_multilabels`` (test.py:987-1062) computes once the embeddings exist: cosine self-retrieval,
    mAP with Jaccard > t relevance for every t, and the Precision@K / Recall@K (hit-rate) table.  Never builds the
    N x N matrix: the full-ranking AP runs on a k = N-1 search, the table on a top-max(k) search.
    -> {"mAP": {t: float}, "precision_recall_at_k": {k: (precision %, recall %)}}; ``save_path`` writes the
    ``embeds`` / ``labels`` bundle of test.py:1059-1061."""
    _require_cuda(embeds)
    lab = labels_multihot.to(embeds.device)
    out = {"mAP": {float(t): compute_map_multilabel_from_embeddings(embeds, lab, float(t)) for t in thresholds}}
    kmax = min(max(int(k) for k in k_values), embeds.shape[0] - 1)
    _, idx = search(embeds, embeds, kmax, "cosine", normalize=True, exclude_self=True)
    out["precision_recall_at_k"] = multilabel_hit_rate_from_topk(idx, lab, lab, k_values)
    if save_path is not None:
        np.savez(save_path, embeds=embeds.detach().cpu().numpy(), labels=lab.detach().cpu().numpy())
    return out
