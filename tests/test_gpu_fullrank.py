"""Full-ranking metrics without the N x N ranking (csrc/rank_positives.cu, fullrank.py).

* knn_rank_of_positives / knn_ap_from_ranks / knn_ap_sklearn_from_ranks against the CPU oracle (oracle/rank_oracle.py +
  the restated reference AP functions of oracle/reference_metrics.py): bit-exact, ties / massive ties / all-equal rows /
  dropped rows / both orders / every relevance mode;
* the embeddings-in metric functions against the DENSE path they replace (full ranking by knn_rank_rows + the ranked-list
  kernels) at N <= 8 k: bit-equal;
* NIH scale (112 k x 1024 self-retrieval): evaluate_map_embeddings runs without any N x N allocation; the per-query APs
  of a 2 k-query slice equal the oracle's (exact-fp32 scores restated in C + the restated sklearn AP);
* fusion metrics with non-unique image paths against the golden of the real reference function."""
import json
import os
import time

import numpy as np
import pytest
import torch

import oracle
from oracle import make_golden_fullrank as mgf
from oracle import rank_oracle as RO
from oracle import reference_metrics as RM
from oracle import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def knn():
    import b200knn

    b200knn.load_library()
    return b200knn


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _popcount(x):
    return np.array([bin(int(v)).count("1") for v in x.ravel()]).reshape(x.shape)


def _relevance(mode, ql, gl, thr):
    if mode == 0:
        return gl[None, :] == ql[:, None]
    inter, uni = _popcount(ql[:, None] & gl[None, :]), _popcount(ql[:, None] | gl[None, :])
    if mode == 1:
        return (inter.astype(np.float32) / (uni.astype(np.float32) + np.float32(1e-8))) > np.float32(thr)
    if mode == 2:
        return (inter.astype(np.float64) / (uni.astype(np.float64) + 1e-8)) > thr
    return inter > 0


CASES = [
    # nq, ng, mode, ties, drop_self, group, quant, largest
    (4, 1, 0, True, True, False, None, True),
    (4, 2, 0, False, True, False, None, True),
    (5, 37, 0, True, False, False, None, True),
    (5, 400, 0, False, True, True, None, True),
    (5, 2000, 1, True, True, False, None, True),
    (5, 5000, 2, False, False, False, None, False),
    (4, 9000, 3, True, True, True, None, True),
    (4, 40000, 0, True, True, False, None, True),
    (4, 30000, 0, True, True, False, 4, True),        # ~30 distinct scores: every bin is refined on the full key
    (4, 30000, 1, False, False, False, 4, False),
    (3, 70000, 0, True, False, False, 1000, True),
]


@pytest.mark.parametrize("nq,ng,mode,ties,drop,group,quant,largest", CASES)
def test_rank_of_positives_and_ap_kernels_match_the_oracle(knn, nq, ng, mode, ties, drop, group, quant, largest):
    FR = knn.fullrank
    rs = np.random.RandomState(ng + mode)
    s = rs.standard_normal((nq, ng)).astype(np.float32)
    if quant:
        s = (np.round(s * quant) / quant).astype(np.float32)
    s[0, : min(ng, 3)] = -0.0                                       # -0.0 and +0.0 tie
    if mode == 0:
        ql, gl = rs.randint(0, 3, nq).astype(np.int64), rs.randint(0, 3, ng).astype(np.int64)
    else:
        ql, gl = rs.randint(1, 1 << 14, nq).astype(np.int64), rs.randint(1, 1 << 14, ng).astype(np.int64)
    thr, off = 0.4, 2
    rel = _relevance(mode, ql, gl, thr)
    qg = gg = None
    if group:
        gg = rs.randint(0, max(2, ng // 3), ng).astype(np.int64)
        qg = rs.randint(0, max(2, ng // 3), nq).astype(np.int64)
    rp = FR.rank_of_positives(dev(s), mode, dev(ql), dev(gl), largest_first=largest, jaccard_threshold=thr,
                              self_offset=off, drop_self=drop, q_group=None if qg is None else dev(qg),
                              g_group=None if gg is None else dev(gg), ties=ties)
    kap = (1, 5, 10)
    st = {k: v.cpu().numpy() for k, v in FR.ap_from_ranks(rp, kap, drop).items()}
    aps = FR.ap_sklearn_from_ranks(rp).cpu().numpy() if ties else None
    pr, npos, nrk = rp["pos_ranks"].cpu().numpy(), rp["npos"].cpu().numpy(), rp["nranked"].cpu().numpy()
    for q in range(nq):
        dropped = np.zeros(ng, bool)
        if drop and 0 <= q + off < ng:
            dropped[q + off] = True
        if group:
            dropped |= gg == qg[q]
        r = rel[q] & ~dropped
        pos, ge, tg, n, ngr = RO.rank_of_positives(s[q], r, largest, dropped)
        assert npos[q] == len(pos) and nrk[q] == n
        assert np.array_equal(pr[q, :len(pos)], pos)
        if ties:
            assert np.array_equal(rp["pos_ge"][q, :len(pos)].cpu().numpy(), ge)
            assert np.array_equal(rp["pos_tgroup"][q, :len(pos)].cpu().numpy(), tg)
            assert int(rp["ngroups"][q]) == ngr
        ppos = list(pos) + ([n] if drop else [])                      # self_last_positive appends rank n
        if ppos:
            assert RM.compute_ap(ppos, len(ppos)) == st["ap_trapz"][q]          # test.py:58-92
            pp = np.asarray(ppos) + 1
            for t, k in enumerate(kap):
                kq = min(max(pp), k)
                assert (pp <= kq).sum() / kq == st["prs"][q, t]              # test.py:137-140
        else:
            assert np.isnan(st["ap_trapz"][q])
        ps = 0.0
        for c, rnk in enumerate(pos):
            ps += (c + 1) / (rnk + 1)                                          # test.py:974-981
        assert ps == st["prec_sum"][q] and st["first"][q] == (pos[0] + 1 if len(pos) else 0)
        assert all(st["hits_at"][q, t] == (pos < k).sum() for t, k in enumerate(kap))
        if ties:
            rows = np.flatnonzero(~dropped)
            order = rows[RO.order_row(s[q][rows], largest)]
            if r.any():
                assert RM.average_precision_ranked(s[q][order], r[order]) == aps[q]   # sklearn AP, bit for bit
            else:
                assert np.isnan(aps[q])


def test_all_equal_scores_rank_by_row(knn):
    FR = knn.fullrank
    n = 20000
    lab = torch.zeros(n, dtype=torch.int64).cuda()
    rp = FR.rank_of_positives(torch.zeros((2, n), device="cuda"), 0, lab[:2], lab, ties=True)
    assert np.array_equal(rp["pos_ranks"][0].cpu().numpy(), np.arange(n))
    assert int(rp["ngroups"][0]) == 1 and bool((rp["pos_ge"][0] == n).all()) and bool((rp["pos_tgroup"][0] == 0).all())


def _dense_single(knn, e, lab, metric="cosine", normalize=True):
    """the dense path these functions used before: full self-excluded ranking (dense scores + knn_rank_rows)"""
    M = knn.metrics
    n = e.shape[0]
    _, idx = knn.search(e, e, n - 1, metric, normalize=normalize, exclude_self=True)
    rel, _ = M.relevance_single(idx, lab, lab)
    return idx, rel


@pytest.mark.parametrize("n,d", [(700, 96), (3000, 64), (8192, 32)])
def test_embeddings_in_metrics_equal_the_dense_path(knn, n, d):
    M = knn.metrics
    x, lab = synth.clustered(n, d, 5, seed=n, noise=3.0)
    x[7] = x[3]                                                     # duplicate rows: exact ties
    x[n // 2] = x[n // 2 + 1]
    e, labd = dev(x), dev(lab)
    # ---- single label: D6 (train.py:399-441), D11 (fusion_eval/metrics.py:41-94), D2 (test.py:95-146)
    idx, rel = _dense_single(knn, e, labd)
    hits, first, _, ps = M.ranked_stats(rel)
    hits_np, psn, first_np = hits.cpu().numpy(), ps.cpu().numpy(), first.cpu().numpy()
    aps = np.where(hits_np > 0, psn / np.maximum(hits_np, 1), 0.0)
    want = {"mAP": float(np.mean(aps) * 100.0)}
    for k in (1, 5, 10):
        want[f"R@{k}"] = float(np.mean(((first_np > 0) & (first_np <= k)).astype(np.float32))) * 100.0
    assert M._compute_single_label_retrieval_metrics(e, labd) == want
    assert M.evaluate_retrieval_metrics(e, lab.tolist(), None, (1, 5, 10)) == \
        M.retrieval_metrics_from_ranking(idx, lab.tolist(), (1, 5, 10))
    for metric in ("cosine", "l2"):
        en = knn.normalize(e)
        dists = knn.scores_dense(en, en, metric, self_mode="exclude")
        dists = -dists if metric == "l2" else dists
        ranks = knn.rank_rows(dists.t().contiguous(), largest_first=True)
        a = M.compute_map(ranks.t(), labd, [1, 5, 10])
        b = M.compute_map_from_embeddings(en, labd, [1, 5, 10], metric=metric)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    # ---- multi label: D4 (test.py:941-985), D7 (train.py:444-487), D8 (nih_multilabel_training.py:66-99)
    ml = synth.multihot(n, seed=n + 1)
    emb = dev(synth.labelset_clustered(ml, d, n + 2, 1.0))
    mld = dev(ml)
    m = M.pack_multihot(mld)
    vals, idx = knn.search(emb, emb, n - 1, "cosine", normalize=True, exclude_self=True)
    for thr in (0.25, 0.4, 0.5):
        relj, _ = M.relevance_multilabel(idx, m, m, thr, arith="fp32")
        hits, first, ap, _ = M.ranked_stats(relj)
        hn = hits.cpu().numpy()
        a = ap.cpu().numpy()[hn > 0]
        assert M.compute_map_multilabel_from_embeddings(emb, mld, thr) == (float(np.mean(a)) if len(a) else 0)
    relj, _ = M.relevance_multilabel(idx, m, m, 0.4, arith="fp32")
    hits, first, _, _ = M.ranked_stats(relj)
    hn, fn = hits.cpu().numpy(), first.cpu().numpy()
    a = M.ap_sklearn(vals, relj).cpu().numpy()[hn > 0]
    want = {"mAP": float(np.mean(a) * 100.0) if len(a) else 0.0}
    for k in (1, 5, 10):
        want[f"R@{k}"] = float(np.mean(((fn > 0) & (fn <= k)).astype(np.float64)) * 100.0)
    assert M._compute_multilabel_retrieval_metrics(emb, mld) == want
    vals, idx = knn.search(emb, emb, n, "cosine", normalize=True, self_mode="minus1")
    relj, _ = M.relevance_multilabel(idx, m, m, 0.4, arith="fp32")
    hn = M.ranked_stats(relj)[0].cpu().numpy()
    a = M.ap_sklearn(vals, relj).cpu().numpy()[hn > 0]
    assert M.evaluate_map_embeddings(emb, mld, 0.4) == float(np.mean(a) * 100.0)
    # chunking over the queries does not change a bit
    st1 = knn.fullrank.full_ranking_stats(emb, emb, 1, m, m, normalize=True, self_mode="exclude", drop_self=True,
                                          jaccard_threshold=0.4, sklearn_ap=True, rows_per_chunk=n)
    st2 = knn.fullrank.full_ranking_stats(emb, emb, 1, m, m, normalize=True, self_mode="exclude", drop_self=True,
                                          jaccard_threshold=0.4, sklearn_ap=True, rows_per_chunk=257)
    for key in st1:
        assert torch.equal(st1[key], st2[key]) or (st1[key].dtype.is_floating_point and
                                                   torch.equal(torch.nan_to_num(st1[key]), torch.nan_to_num(st2[key])))


def test_fusion_metrics_with_shared_image_paths_match_the_reference(knn):
    with open(os.path.join(GOLDEN, "golden_fullrank.json")) as fh:
        g = json.load(fh)
    x, labels, paths = mgf.dup_inputs()
    assert len(set(paths)) == g["n_unique_paths"] < len(paths)
    got = knn.metrics.evaluate_retrieval_metrics(x, labels, paths, (1, 3, 5, 10, 20))
    assert set(got) == set(g["dup_fusion"])
    for k, v in g["dup_fusion"].items():
        assert got[k] == pytest.approx(v, rel=1e-12), k


def test_nih_scale_full_ranking_map_without_an_n_by_n_matrix(knn):
    """BASELINE config 3's gallery (112 k x 1024, 14 NIH-like labels) as SELF-retrieval: the reference's evaluate_map
    (nih_multilabel_training.py:66-99) would build a 50 GB similarity matrix and call sklearn 112 k times."""
    M, FR = knn.metrics, knn.fullrank
    n, d = 112_000, 1024
    ml = synth.multihot(n, seed=3)
    mld = dev(ml)
    rs = np.random.RandomState(33)
    mu = rs.standard_normal((14, d)).astype(np.float32)
    emb = torch.from_numpy(ml).cuda() @ torch.from_numpy(mu).cuda()
    gen = torch.Generator(device="cuda").manual_seed(7)
    emb = knn.normalize(emb / dev(np.maximum(ml.sum(1, keepdims=True), 1.0))
                        + 1.0 * torch.randn((n, d), generator=gen, device="cuda"))
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    t0 = time.perf_counter()
    value = M.evaluate_map_embeddings(emb, mld, 0.4)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    peak = torch.cuda.max_memory_allocated() - base
    print(f"\nevaluate_map_embeddings 112k x 1024: mAP {value:.4f} % in {dt:.3f} s, peak transient memory {peak / 2**30:.2f} GiB")
    assert 0.0 < value < 100.0
    assert peak < 12 * 2**30, "an N x N int64 ranking would be 100 GB"
    assert dt < 5.0
    # a 2 k-query slice against the oracle: exact-fp32 scores restated in C, stable ranking, restated sklearn AP
    s0, nqs = 40_000, 2048
    m = M.pack_multihot(mld)
    st = FR.full_ranking_stats(emb[s0:s0 + nqs], emb, FR.REL_JACCARD_F32, m[s0:s0 + nqs], m, metric="cosine",
                               self_mode="minus1", drop_self=False, query_offset=s0, jaccard_threshold=0.4,
                               sklearn_ap=True, kappas=(1, 10))
    got, npos = st["ap_sklearn"].cpu().numpy(), st["npos"].cpu().numpy()
    eh = emb.cpu().numpy()
    sc = oracle.scores(eh[s0:s0 + nqs], eh, "cosine", "minus1", s0)
    for q in range(nqs):
        rel = RM.jaccard_fp32(ml[s0 + q], ml) > np.float32(0.4)
        assert npos[q] == rel.sum()
        order = RO.order_row(sc[q])
        if rel.any():
            assert RM.average_precision_ranked(sc[q][order], rel[order]) == got[q], q
        else:
            assert np.isnan(got[q])


@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_tensor_core_dense_scores(knn, metric):
    """knn_scores_dense on the tcgen05 kernel: bf16 rows = fp32 accumulation of the bf16-rounded inputs; bf16x3 = the
    error-free split, |error| <= the filter bound of the exact engine (8.04 * 2^-18 + accumulation) |q||g|."""
    rs = np.random.RandomState(5)
    nq, ng, d = 300, 5000, 200                                  # odd block count, ragged tile, d not a multiple of 64
    q = oracle.normalize(rs.standard_normal((nq, d)).astype(np.float32))
    g = oracle.normalize(rs.standard_normal((ng, d)).astype(np.float32))
    exact = knn.scores_dense(dev(q), dev(g), metric, self_mode="exclude", query_offset=17).cpu().numpy()
    assert np.array_equal(exact, oracle.scores(q, g, metric, "exclude", 17))
    x3 = knn.scores_dense(dev(q), dev(g), metric, self_mode="exclude", query_offset=17, precision="bf16x3").cpu().numpy()
    fin = np.isfinite(exact)
    assert np.array_equal(np.isinf(x3), ~fin)                   # the masked self entries
    tol = 3e-5 if metric == "cosine" else 2e-3                  # d = sqrt(d^2): the error of d^2 (~4e-5) over 2 d
    assert np.abs(x3[fin] - exact[fin]).max() < tol
    b = knn.scores_dense(dev(q), dev(g), metric, self_mode="minus1" if metric == "cosine" else "keep",
                         query_offset=17, precision="bf16").cpu().numpy()
    qb, gb = oracle.bf16_round(q), oracle.bf16_round(g)
    want = qb.astype(np.float64) @ gb.astype(np.float64).T
    if metric == "l2":
        want = np.sqrt(np.maximum((qb.astype(np.float64) ** 2).sum(1)[:, None] + (gb.astype(np.float64) ** 2).sum(1)[None, :]
                                  - 2 * want, 0))
        assert np.abs(b - want).max() < 2e-3
    else:
        r = np.arange(nq)
        assert np.all(b[r, r + 17] == -1.0)
        b[r, r + 17] = want[r, r + 17]
        assert np.abs(b - want).max() < 1e-5


def test_full_ranking_metrics_on_the_tensor_cores_track_the_exact_ones(knn):
    """precision="bf16x3": the score block comes from the tensor cores; only scores closer than ~1e-5 may swap ranks, so
    the metrics agree with the exact mode to ~1e-6."""
    M = knn.metrics
    n, d = 6000, 128
    ml = synth.multihot(n, seed=12)
    emb = dev(synth.labelset_clustered(ml, d, 13, 1.0))
    mld = dev(ml)
    a = M.evaluate_map_embeddings(emb, mld, 0.4)
    b = M.evaluate_map_embeddings(emb, mld, 0.4, precision="bf16x3")
    c = M.evaluate_map_embeddings(emb, mld, 0.4, precision="bf16")
    assert abs(a - b) < 1e-4 and abs(a - c) < 0.05, (a, b, c)
    x, lab = synth.clustered(n, d, 5, seed=14, noise=3.0)
    ea = M._compute_single_label_retrieval_metrics(dev(x), dev(lab))
    eb = M._compute_single_label_retrieval_metrics(dev(x), dev(lab), precision="bf16x3")
    for k in ea:
        assert abs(ea[k] - eb[k]) < 1e-3, (k, ea[k], eb[k])
