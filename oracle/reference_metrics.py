"""CPU restatement of the reference's retrieval METRICS -- TEST INFRASTRUCTURE ONLY.

Each function restates one reference function (file:line in its docstring) over explicit rankings, so the only
thing that differs from running the reference itself is where the ranking comes from: here it is always the
deterministic order (score best-first, ties by ascending gallery index) instead of torch's / numpy's unstable
sort.  Pinned against the live reference by tests/golden/*.npz (oracle/make_golden.py).
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, Iterable, List, Sequence

import numpy as np


# ---- D1 ---------------------------------------------------------------------------------------
def retrieval_accuracy(topk_idx: np.ndarray, qlabels: np.ndarray, glabels: np.ndarray, topk=(1,)) -> List[np.float32]:
    """test.py:38-54: 100 * (#queries with a label match in the first k) / #queries, in float32."""
    match = glabels[topk_idx] == qlabels[:, None]
    n = np.float32(100.0 / len(qlabels))
    return [np.float32(np.float32(match[:, :k].any(axis=1).sum()) * n) for k in topk]


# ---- D2 ---------------------------------------------------------------------------------------
def compute_ap(pos_ranks: Sequence[int], nres: int) -> float:
    """test.py:58-92: trapezoidal area under the PR curve; ranks are 0-based positions of the positives."""
    ap = 0
    recall_step = 1.0 / nres
    for j, rank in enumerate(pos_ranks):
        p0 = 1.0 if rank == 0 else float(j) / rank
        p1 = float(j + 1) / (rank + 1)
        ap += (p0 + p1) * recall_step / 2.0
    return ap


def compute_map(ranks_rowmajor: np.ndarray, qlabels: np.ndarray, glabels: np.ndarray, kappas=()):
    """test.py:95-146 with ranks given row-major [nq, ng] (the reference takes the transpose).  Positives of query
    i = every gallery row with its label (for self-retrieval this includes the query itself, which sits last)."""
    nq = len(qlabels)
    aps = np.zeros(nq)
    prs = np.zeros((nq, len(kappas)))
    pr = np.zeros(len(kappas))
    mAP, nempty = 0.0, 0
    for i in range(nq):
        positives = np.where(glabels == qlabels[i])[0]
        if positives.shape[0] == 0:
            aps[i] = np.nan
            prs[i, :] = np.nan
            nempty += 1
            continue
        pos = np.arange(ranks_rowmajor.shape[1])[np.isin(ranks_rowmajor[i], positives)]
        ap = compute_ap(pos, len(positives))
        mAP = mAP + ap
        aps[i] = ap
        pos = pos + 1
        for j, kappa in enumerate(kappas):
            kq = min(max(pos), kappa)
            prs[i, j] = (pos <= kq).sum() / kq
        pr = pr + prs[i, :]
    return mAP / (nq - nempty), aps, pr / (nq - nempty), prs


# ---- D3 ---------------------------------------------------------------------------------------
def majority_vote(labels_in_rank_order: Sequence, tie: str = "first"):
    """test.py:149-161 (Counter.most_common -> first label reaching the max count); tie='smallest' is the
    torch.mode / np.unique+argmax convention (train_ath.py:208, evaluate_medsiglip.py:156-157)."""
    cnt = Counter(list(labels_in_rank_order))
    if tie == "first":
        return cnt.most_common(1)[0][0]
    best = max(cnt.values())
    return min(l for l, c in cnt.items() if c == best)


def compute_classification_metrics(topk_idx: np.ndarray, qlabels: np.ndarray, glabels: np.ndarray,
                                   k_values=(1, 5, 10, 15, 20)) -> Dict[int, Dict[str, float]]:
    """test.py:164-223 from a ranking [nq, >=max(k)]."""
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score

    out = {}
    for k in k_values:
        pred = [majority_vote(glabels[topk_idx[i, :k]]) for i in range(len(qlabels))]
        true = list(qlabels)
        out[k] = {
            "precision_macro": precision_score(true, pred, average="macro", zero_division=0) * 100.0,
            "recall_macro": recall_score(true, pred, average="macro", zero_division=0) * 100.0,
            "f1_macro": f1_score(true, pred, average="macro", zero_division=0) * 100.0,
            "precision_weighted": precision_score(true, pred, average="weighted", zero_division=0) * 100.0,
            "recall_weighted": recall_score(true, pred, average="weighted", zero_division=0) * 100.0,
            "f1_weighted": f1_score(true, pred, average="weighted", zero_division=0) * 100.0,
            "accuracy": accuracy_score(true, pred) * 100.0,
        }
    return out


# ---- Jaccard relevance ---------------------------------------------------------------------------
def jaccard_fp32(a: np.ndarray, B: np.ndarray) -> np.ndarray:
    """train.py:462-464 / nih_multilabel_training.py:90-92 / test.py:956-959: float32 tensors."""
    a, B = a.astype(np.float32), B.astype(np.float32)
    inter = (a[None, :] * B).sum(axis=1, dtype=np.float32)
    union = np.minimum(a[None, :] + B, np.float32(1.0)).sum(axis=1, dtype=np.float32)
    return inter / (union + np.float32(1e-8))


def jaccard_fp64(a: Sequence[float], b: Sequence[float]) -> float:
    """evaluate_nih_zilliz.py:12-17: float32 sums converted to Python floats, double division."""
    a32, b32 = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    inter = float((a32 * b32).sum())
    union = float(np.clip(a32 + b32, 0.0, 1.0).sum())
    return inter / (union + 1e-8)


# ---- D4 ---------------------------------------------------------------------------------------
def compute_map_multilabel(ranks_rowmajor: np.ndarray, labels: np.ndarray, threshold: float = 0.5):
    """test.py:941-985: rank-by-rank AP over the full ranking, relevance = Jaccard > threshold (fp32), self
    removed, queries without relevant items skipped.  ranks_rowmajor [n, n] may contain the query itself."""
    n = labels.shape[0]
    aps = []
    for i in range(n):
        rel = (jaccard_fp32(labels[i], labels) > np.float32(threshold)).astype(float)
        rel[i] = 0
        if rel.sum() > 0:
            count_pos, ap = 0, 0
            for rank, j in enumerate(ranks_rowmajor[i]):
                if j >= 0 and rel[j] > 0:
                    count_pos += 1
                    ap += count_pos / (rank + 1)
            aps.append(ap / rel.sum())
    return np.mean(aps) if aps else 0


# ---- D5 ---------------------------------------------------------------------------------------
def multilabel_precision_recall_at_k(topk_idx: np.ndarray, qlabels: np.ndarray, glabels: np.ndarray,
                                     k_values=(1, 5, 10, 15, 20)):
    """test.py:1031-1056: match = shares >= 1 label; Precision@K = mean(#match/K); Recall@K = hit-rate."""
    out = {}
    nq = qlabels.shape[0]
    for k in k_values:
        total_p, total_r = 0, 0
        for i in range(nq):
            matches = (glabels[topk_idx[i, :k]] * qlabels[i]).sum(axis=1) > 0
            nm = np.sum(matches)
            total_p += nm / k
            if nm > 0:
                total_r += 1
        out[k] = ((total_p / nq) * 100, (total_r / nq) * 100)
    return out


# ---- sklearn AP (D7, D8, D9 use it) ----------------------------------------------------------------
def average_precision_ranked(scores_desc: np.ndarray, rel: np.ndarray) -> float:
    """sklearn.metrics.average_precision_score restated for a list already sorted by descending score:
    thresholds at the last index of each run of equal scores; AP = -sum(diff(recall) * precision[:-1]) on the
    reversed curves (sklearn/metrics/_ranking.py)."""
    rel = np.asarray(rel, dtype=np.float64)
    s = np.asarray(scores_desc)
    distinct = np.where(np.diff(s))[0]
    thr = np.r_[distinct, len(s) - 1]
    tps = np.cumsum(rel)[thr]
    fps = 1 + thr - tps
    ps = tps + fps
    precision = np.zeros_like(tps)
    np.divide(tps, ps, out=precision, where=ps != 0)
    recall = tps / tps[-1]
    precision = np.hstack((precision[::-1], [1.0]))
    recall = np.hstack((recall[::-1], [0.0]))
    return float(-np.sum(np.diff(recall) * precision[:-1]))


# ---- D6 ---------------------------------------------------------------------------------------
def single_label_retrieval_metrics(full_idx: np.ndarray, labels: np.ndarray, topk=(1, 5, 10)) -> Dict[str, float]:
    """train.py:399-441 from the full self-excluded ranking [n, n-1].  The reference does the AP arithmetic in
    torch float32 (cumsum / positions, .sum()); here it is float64 -> compared with a 1e-5 tolerance."""
    n = len(labels)
    if n <= 1:
        return {"mAP": 0.0, **{f"R@{k}": 0.0 for k in topk}}
    rel = labels[full_idx] == labels[:, None]
    counts = (labels[:, None] == labels[None, :]).sum(axis=1) - 1
    aps = []
    for i in range(n):
        hits = np.flatnonzero(rel[i])
        if counts[i] <= 0 or hits.size == 0:
            aps.append(0.0)
            continue
        prec = np.cumsum(rel[i].astype(np.float64))[hits] / (hits + 1.0)
        aps.append(float(prec.sum() / counts[i]))
    m = {"mAP": float(np.mean(aps) * 100.0)}
    for k in topk:
        kk = min(k, rel.shape[1])
        m[f"R@{k}"] = float(rel[:, :kk].any(axis=1).astype(np.float32).mean()) * 100.0 if kk > 0 else 0.0
    return m


# ---- D7 / D8 ------------------------------------------------------------------------------------
def multilabel_retrieval_metrics(full_val: np.ndarray, full_idx: np.ndarray, labels: np.ndarray, topk=(1, 5, 10),
                                 relevance_threshold: float = 0.4) -> Dict[str, float]:
    """train.py:444-487 from the full self-excluded ranking (values + indices, [n, n-1])."""
    n = labels.shape[0]
    aps, recalls = [], {k: [] for k in topk}
    for i in range(n):
        rel = (jaccard_fp32(labels[i], labels) > np.float32(relevance_threshold)).astype(np.float64)
        rel[i] = 0.0
        ranked_rel = rel[full_idx[i]]
        if rel.sum() > 0:
            aps.append(average_precision_ranked(full_val[i], ranked_rel))
        for k in topk:
            kk = min(k, ranked_rel.size)
            recalls[k].append(float(ranked_rel[:kk].any()) if kk > 0 else 0.0)
    m = {"mAP": float(np.mean(aps) * 100.0) if aps else 0.0}
    for k in topk:
        m[f"R@{k}"] = float(np.mean(recalls[k]) * 100.0) if recalls[k] else 0.0
    return m


def evaluate_map(full_val: np.ndarray, full_idx: np.ndarray, labels: np.ndarray, jaccard_threshold: float = 0.4):
    """nih_multilabel_training.py:83-99 from the full ranking INCLUDING self at similarity -1 ([n, n])."""
    aps = []
    for i in range(labels.shape[0]):
        rel = (jaccard_fp32(labels[i], labels) > np.float32(jaccard_threshold)).astype(np.float64)
        if rel.sum() > 0:
            aps.append(average_precision_ranked(full_val[i], rel[full_idx[i]]))
    return float(np.mean(aps) * 100.0) if aps else 0.0


# ---- D9 ---------------------------------------------------------------------------------------
def evaluate_results(topk_val: np.ndarray, topk_idx: np.ndarray, qlabels: np.ndarray, glabels: np.ndarray,
                     jaccard_threshold: float = 0.4, ks: Iterable[int] = (1, 5, 10, 20, 50)) -> Dict[str, float]:
    """evaluate_nih_zilliz.py:34-64 over hit lists given as arrays (scores + gallery rows, best first)."""
    aps = []
    ks = list(ks)
    P = {k: [] for k in ks}
    R = {k: [] for k in ks}
    for i in range(topk_idx.shape[0]):
        rels = [1.0 if jaccard_fp64(qlabels[i], glabels[j]) > jaccard_threshold else 0.0 for j in topk_idx[i]]
        total = int(sum(rels))
        if total > 0:
            aps.append(average_precision_ranked(topk_val[i], np.asarray(rels)))
        for k in ks:
            kk = min(k, len(rels))
            P[k].append(float(np.mean(rels[:kk])) if rels else 0.0)
            R[k].append(float(np.sum(rels[:kk]) / total) if total > 0 else 0.0)
    m = {"mAP": float(np.mean(aps) * 100.0) if aps else 0.0, "num_queries": float(topk_idx.shape[0]),
         "num_valid_ap_queries": float(len(aps))}
    for k in ks:
        m[f"P@{k}"] = float(np.mean(P[k]) * 100.0) if P[k] else 0.0
        m[f"R@{k}"] = float(np.mean(R[k]) * 100.0) if R[k] else 0.0
    return m


# ---- D10 --------------------------------------------------------------------------------------
def ath_compute_metrics(sorted_idx: np.ndarray, qlabels: np.ndarray, glabels: np.ndarray, topk_values=(1, 5, 10)):
    """test_ath.py:90-172 from the ascending-distance ranking [nq, >= max(topk)]."""
    retrieval = {}
    total_rel = [(glabels == int(l)).sum() for l in qlabels]
    for topk in topk_values:
        hit, ap, rr, vote, pk, rk = [], [], [], [], [], []
        for row, label in enumerate(qlabels):
            label = int(label)
            ranked = glabels[sorted_idx[row, :topk]]
            matches = (ranked == label).astype(np.int32)
            hit.append(float(matches.any()))
            nrel = matches.sum()
            pk.append(nrel / topk)
            rk.append(nrel / total_rel[row] if total_rel[row] > 0 else 0.0)
            if nrel == 0:
                ap.append(0.0)
                rr.append(0.0)
            else:
                first, psum, positives = None, 0.0, 0
                for rank, m in enumerate(matches, start=1):
                    if m:
                        positives += 1
                        psum += positives / rank
                        if first is None:
                            first = rank
                ap.append(psum / positives)
                rr.append(1.0 / first)
            vote.append(float(majority_vote(ranked.tolist()) == label))
        retrieval[topk] = {"mhr": float(np.mean(hit)), "map": float(np.mean(ap)), "mrr": float(np.mean(rr)),
                           "mp@k": float(np.mean(pk)), "r@k": float(np.mean(rk)),
                           "majority_acc": float(np.mean(vote))}
    return retrieval


# ---- D11 --------------------------------------------------------------------------------------
def fusion_evaluate_retrieval_metrics(full_idx: np.ndarray, labels: Sequence, k_values=(1, 5, 10)):
    """fusion_eval/metrics.py:41-94 from the full self-excluded ranking [n, n-1] (unique image paths)."""
    labels = np.asarray(labels)
    k_values = sorted(set(int(k) for k in k_values))
    aps, P, R = [], {k: [] for k in k_values}, {k: [] for k in k_values}
    for q in range(len(labels)):
        relevant = labels[full_idx[q]] == labels[q]
        count = int(np.sum(labels == labels[q]) - 1)
        if count <= 0:
            aps.append(0.0)
            for k in k_values:
                P[k].append(0.0)
                R[k].append(0.0)
            continue
        hits = np.flatnonzero(relevant)
        if len(hits) == 0:
            aps.append(0.0)
        else:
            prec = np.cumsum(relevant.astype(np.int32))[hits] / (hits + 1)
            aps.append(float(np.sum(prec) / count))
        for k in k_values:
            h = int(np.sum(relevant[:k]))
            P[k].append(h / k)
            R[k].append(1.0 if h > 0 else 0.0)
    m = {"num_samples": float(len(labels)), "mAP": float(np.mean(aps) * 100.0)}
    for k in k_values:
        m[f"mP@{k}"] = float(np.mean(P[k]) * 100.0)
        m[f"R@{k}"] = float(np.mean(R[k]) * 100.0)
    return m


# ---- D12 --------------------------------------------------------------------------------------
def medsiglip_evaluate_retrieval(topk_idx: np.ndarray, labels: np.ndarray, topk_values) -> Dict[str, float]:
    """evaluate_medsiglip.py:142-163 from the row-wise ranking of ``f @ f.T`` with the diagonal at -1 [n, >= max k]:
    R@k any-hit, majority by np.unique + argmax (smallest label on ties), accuracy and macro F1 in percent."""
    from sklearn.metrics import accuracy_score, f1_score

    out = {}
    for k in topk_values:
        retrieved = labels[topk_idx[:, :k]]
        hit = (retrieved == labels[:, None]).any(axis=1).mean() * 100.0
        majority = np.array([majority_vote(row.tolist(), tie="smallest") for row in retrieved])
        out[f"r_at_{k}"] = float(hit)
        out[f"majority_accuracy_at_{k}"] = accuracy_score(labels, majority) * 100.0
        out[f"majority_macro_f1_at_{k}"] = f1_score(labels, majority, average="macro") * 100.0
    return out


# ---- D10' -------------------------------------------------------------------------------------
def ath_train_retrieval_metrics(sorted_idx: np.ndarray, qlabels: np.ndarray, glabels: np.ndarray, topk_values):
    """train_ath.py:171-218 from the ascending-distance ranking: mhr, map, mrr and majority accuracy with torch.mode
    (smallest label on ties)."""
    full = ath_compute_metrics(sorted_idx, qlabels, glabels, topk_values)
    out = {}
    for topk in topk_values:
        vote = [float(majority_vote(glabels[sorted_idx[r, :topk]].tolist(), tie="smallest") == int(l))
                for r, l in enumerate(qlabels)]
        out[topk] = {"mhr": full[topk]["mhr"], "map": full[topk]["map"], "mrr": full[topk]["mrr"],
                     "majority_acc": float(np.mean(vote))}
    return out


# ---- D14 --------------------------------------------------------------------------------------
def chestmir_accuracy_from_ranks(ranks_rowmajor: np.ndarray, labels: np.ndarray, topk) -> np.ndarray:
    """chestmir_eval.py:262-272 with the ranking given row-major [n, >= max k] (the reference indexes columns)."""
    n = len(labels)
    return np.array([(sum(bool(np.any(labels[ranks_rowmajor[i, :k]] == labels[i])) for i in range(n)) * 100.0) / max(1, n)
                     for k in topk], dtype=np.float64)


def chestmir_classification_from_ranks(labels: np.ndarray, ranks_rowmajor: np.ndarray, k_values):
    """chestmir_eval.py:191-259: Counter majority vote + hand-rolled per-class P / R / F1 (f1 = 2pr / (p + r))."""
    out = {}
    n = len(labels)
    for k in k_values:
        y_pred = np.asarray([majority_vote(labels[ranks_rowmajor[i, :k]].tolist()) for i in range(n)], dtype=object)
        y_true = np.asarray(list(labels), dtype=object)
        classes = np.unique(np.concatenate([y_true, y_pred], axis=0))
        P, R, F, S = [], [], [], []
        for c in classes:
            tp = int(np.sum((y_true == c) & (y_pred == c)))
            fp = int(np.sum((y_true != c) & (y_pred == c)))
            fn = int(np.sum((y_true == c) & (y_pred != c)))
            p = tp / (tp + fp) if (tp + fp) > 0 else 0.0
            r = tp / (tp + fn) if (tp + fn) > 0 else 0.0
            P.append(p)
            R.append(r)
            F.append((2.0 * p * r / (p + r)) if (p + r) > 0 else 0.0)
            S.append(int(np.sum(y_true == c)))
        sup = np.asarray(S, dtype=np.float64)
        w = sup / (float(np.sum(sup)) or 1.0)
        out[k] = {"accuracy": float(np.mean(y_true == y_pred)) * 100.0,
                  "precision_macro": float(np.mean(P)) * 100.0, "recall_macro": float(np.mean(R)) * 100.0,
                  "f1_macro": float(np.mean(F)) * 100.0,
                  "precision_weighted": float(np.sum(np.asarray(P) * w)) * 100.0,
                  "recall_weighted": float(np.sum(np.asarray(R) * w)) * 100.0,
                  "f1_weighted": float(np.sum(np.asarray(F) * w)) * 100.0}
    return out


# ---- lesion re-ranking ----------------------------------------------------------
def lesion_rerank(base_val: np.ndarray, base_idx: np.ndarray, lesion_maps, query_choice, rerank_topk: int,
                  global_weight: float) -> np.ndarray:
    """rerank_with_specific_lesion / rerank_with_adaptive_lesion (chestmir_eval.py:507-650) on top-K lists:
    ``base_val / base_idx`` [n, K] = the global self-excluded ranking, ``query_choice[i]`` = (lesion key, query vector) or
    None.  The first ``rerank_topk`` entries are re-ordered by (gw * global + (1 - gw) * best region cosine, global)
    descending (stable); untouched when the query has no vector or no candidate has a region of that lesion."""
    out = base_idx.copy()
    m = min(rerank_topk, base_idx.shape[1])
    for i in range(base_idx.shape[0]):
        if query_choice[i] is None:
            continue
        key, qv = query_choice[i]
        scored, matched = [], 0
        for pos in range(m):
            j = int(base_idx[i, pos])
            cands = lesion_maps[j].get(key, [])
            region = max(float(np.dot(qv, c)) for c in cands) if len(cands) else -1.0
            matched += region >= 0.0
            scored.append((j, global_weight * float(base_val[i, pos]) + (1.0 - global_weight) * region,
                           float(base_val[i, pos])))
        if matched == 0:
            continue
        scored.sort(key=lambda x: (x[1], x[2]), reverse=True)
        out[i, :m] = [s[0] for s in scored]
    return out


# ---- 8(f)-4 tail: training-time pairwise operations (forward values) -------------------------------
def triplet_batch_hard(dist: np.ndarray, labels: np.ndarray, margin: float) -> float:
    """loss.py:60-83 on a given fp32 distance matrix: mean_i max(hardest positive - hardest negative + margin, 0)."""
    d = dist.astype(np.float32)
    n = len(labels)
    same = labels[None, :] == labels[:, None]
    ap = (same & ~np.eye(n, dtype=bool)).astype(np.float32)
    hp = (ap * d).max(axis=1)
    rmax = d.max(axis=1, keepdims=True)
    hn = (d + rmax * (np.float32(1.0) - (~same).astype(np.float32))).min(axis=1)
    t = (hp - hn) + np.float32(margin)
    return float(np.maximum(t, np.float32(0.0)).astype(np.float64).mean())


def triplet_batch_all(dist: np.ndarray, labels: np.ndarray, margin: float):
    """loss.py:86-112 on a given fp32 distance matrix -> (loss, fraction of positive triplets)."""
    d = dist.astype(np.float32)
    n = len(labels)
    total, npos, nvalid = 0.0, 0, 0
    for i in range(n):
        pos = np.flatnonzero((labels == labels[i]) & (np.arange(n) != i))
        neg = np.flatnonzero(labels != labels[i])
        if len(pos) == 0 or len(neg) == 0:
            continue
        t = (d[i, pos][:, None] - d[i, neg][None, :]) + np.float32(margin)
        nvalid += t.size
        t = np.maximum(t, np.float32(0.0))
        total += float(t.astype(np.float64).sum())
        npos += int((t > 1e-16).sum())
    return total / (npos + 1e-16), npos / (nvalid + 1e-16)


def jaccard_sim_matrix(labels_multihot: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """loss.py:237-242 in float32."""
    lab = labels_multihot.astype(np.float32)
    inter = lab @ lab.T
    sums = lab.sum(axis=1, keepdims=True)
    return inter / ((sums + sums.T - inter) + np.float32(eps))


def nearest_centroid_scores(train: np.ndarray, train_labels: np.ndarray, test: np.ndarray, classes=(0, 1)) -> np.ndarray:
    """anomaly/test_anomaly.py:31-48: float32 class means (rows added in order), float64 direct-form distances, min,
    division by the maximum."""
    cents = []
    for c in classes:
        rows = train[train_labels == c].astype(np.float32)
        acc = rows[0].copy()
        for r in rows[1:]:
            acc = acc + r
        cents.append(acc / np.float32(len(rows)))
    cents = np.stack(cents).astype(np.float64)
    x = test.astype(np.float64)
    out = np.empty(len(x))
    for i in range(len(x)):
        best = np.inf
        for c in cents:
            s = 0.0
            for k in range(x.shape[1]):
                df = x[i, k] - c[k]
                s += df * df
            best = min(best, float(np.sqrt(s)))
        out[i] = best
    return out / out.max()
