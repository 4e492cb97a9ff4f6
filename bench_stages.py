"""Per-stage timings of the hot path (SURVEY 8(d) "CPU baseline": normalise / distance / ranking / each metric
function, separately), GPU next to the reference algorithm on the host cores.  Part of bench.py's cpu_baseline leg
(``python bench.py --stages``): the only place besides the tests where the CPU restatement under oracle/ is executed,
and only as the thing the GPU numbers are printed beside.

One JSON line per (config, stage): {"config", "stage", "gpu_ms", "cpu_ms", "cpu_kind", "cpu_sample", "cores"}.
CPU stages run the reference's own torch primitives (F.normalize, mm, cdist, topk, argsort) and the restated
reference metric functions of oracle/reference_metrics.py (python loops / sklearn calls as in the reference), best of
3 for sub-second stages, once otherwise; large configs are timed on a stated sample and extrapolated linearly.
"""
from __future__ import annotations

import json
import time


def _gpu_ms(fn, reps=20, warmup=3):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def _cpu_ms(fn, budget_s=1.0):
    t0 = time.perf_counter()
    fn()
    first = time.perf_counter() - t0
    if first > budget_s:
        return first * 1e3
    best = first
    for _ in range(2):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def run_stages():
    import torch
    import torch.nn.functional as F

    import b200knn
    from b200knn import metrics as M
    from oracle import cpu_baseline, synth
    from oracle import reference_metrics as rm

    cores = cpu_baseline.host_threads()
    torch.set_num_threads(cores)
    dev = torch.device("cuda", 0)

    def emit(config, stage, gpu_ms, cpu_ms, sample="full size", factor=1.0):
        print(json.dumps({"config": config, "stage": stage, "gpu_ms": round(gpu_ms, 4),
                          "cpu_ms": round(cpu_ms * factor, 3), "speedup": round(cpu_ms * factor / gpu_ms, 1),
                          "cpu_kind": "port", "cpu_sample": sample, "cores": cores}), flush=True)

    # ---------------------------------------------------------------- config 1: 400 x 400 x 1024 self-retrieval, cosine
    x, lab = synth.clustered(400, 1024, 3, seed=0, noise=2.0)
    xt, labt = torch.from_numpy(x), torch.from_numpy(lab)
    xd, labd = xt.to(dev), labt.to(dev)
    name = "c1: 400 x 400 x 1024 fp32 self-retrieval, cosine"
    e_cpu = F.normalize(xt, p=2, dim=1)
    e_dev = b200knn.normalize(xd)
    emit(name, "normalise (F.normalize, test.py:1005)", _gpu_ms(lambda: b200knn.normalize(xd)),
         _cpu_ms(lambda: F.normalize(xt, p=2, dim=1)))

    def cpu_dists():
        d = torch.mm(e_cpu, e_cpu.t())
        d.fill_diagonal_(-float("inf"))
        return d

    dists = cpu_dists()
    emit(name, "similarity + top-10 (mm, fill_diagonal_, topk; test.py:1006-1007,44)",
         _gpu_ms(lambda: b200knn.search(e_dev, e_dev, 10, "cosine", exclude_self=True)),
         _cpu_ms(lambda: cpu_dists().topk(10, 1, True, True)))
    dense = b200knn.scores_dense(e_dev, e_dev, "cosine", self_mode="exclude")
    emit(name, "full column ranking (argsort dim=0, test.py:1090)",
         _gpu_ms(lambda: b200knn.rank_rows(dense.t().contiguous())),
         _cpu_ms(lambda: torch.argsort(dists, dim=0, descending=True)))
    ranks_dev = b200knn.rank_rows(dense.t().contiguous())
    ranks_np = ranks_dev.cpu().numpy()
    _, top20 = b200knn.search(e_dev, e_dev, 20, "cosine", exclude_self=True)
    top20_np = top20.cpu().numpy()
    emit(name, "R@1/5/10 (retrieval_accuracy, test.py:38-54)",
         _gpu_ms(lambda: M.recall_at_k_from_topk(top20, labd, labd, (1, 5, 10))),
         _cpu_ms(lambda: rm.retrieval_accuracy(top20_np, lab, lab, (1, 5, 10))))
    emit(name, "mAP + mP@k over the full ranking (compute_map, test.py:95-146)",
         _gpu_ms(lambda: M.map_full(ranks_dev, labd, labd, (1, 5, 10)), reps=10),
         _cpu_ms(lambda: rm.compute_map(ranks_np, lab, lab, (1, 5, 10))))
    emit(name, "majority-vote classification metrics k=1..20 (test.py:164-223)",
         _gpu_ms(lambda: M.classification_metrics_from_topk(top20, labd, labd), reps=10),
         _cpu_ms(lambda: rm.compute_classification_metrics(top20_np, lab, lab)))

    # ---------------------------------------------------------------- config 2: 600 x 2000 x 256, L2, test_ath metrics
    x, lab = synth.clustered(2600, 256, 3, seed=2, noise=2.0, priors=[0.67, 0.17, 0.16])
    e_cpu = F.normalize(torch.from_numpy(x), p=2, dim=1)
    q_cpu, g_cpu, ql, gl = e_cpu[:600], e_cpu[600:], lab[:600], lab[600:]
    q_dev, g_dev = q_cpu.to(dev), g_cpu.to(dev)
    name = "c2: 600 x 2000 x 256 fp32, L2"
    emit(name, "distance + ranking (cdist + argsort dim=1, test_ath.py:87,100)",
         _gpu_ms(lambda: b200knn.search(q_dev, g_dev, 10, "l2")),
         _cpu_ms(lambda: torch.argsort(torch.cdist(q_cpu, g_cpu, p=2), dim=1)))
    order = torch.argsort(torch.cdist(q_cpu, g_cpu, p=2), dim=1).numpy()
    emit(name, "mHR / mAP@k / mRR / mP@k / R@k / vote (compute_metrics, test_ath.py:90-172), incl. the search",
         _gpu_ms(lambda: M.compute_metrics(q_dev, torch.from_numpy(ql), g_dev, torch.from_numpy(gl)), reps=10),
         _cpu_ms(lambda: rm.ath_compute_metrics(order, ql, gl, (1, 5, 10))))

    # ---------------------------------------------------------------- config 3: 25 000 x 112 000 x 1024, top-50, NIH metrics
    nq, ng, d, k = 25_000, 112_000, 1024, 50
    mlab = synth.multihot(nq + ng, 3)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    g_dev = b200knn.normalize(torch.randn((ng, d), generator=gen, device=dev))
    q_dev = b200knn.normalize(torch.randn((nq, d), generator=gen, device=dev))
    qlab, glab = torch.from_numpy(mlab[:nq]).to(dev), torch.from_numpy(mlab[nq:]).to(dev)
    name = "c3: 25 000 x 112 000 x 1024, top-50, multilabel"
    sq = 1024
    sec, _ = cpu_baseline.time_reference(sq, ng, d, k)
    index = b200knn.FlatIndex(d, "cosine", "fp32").adopt(g_dev)
    emit(name, "normalise + similarity + top-50, exact fp32 (F.normalize, mm, topk)",
         _gpu_ms(lambda: index.search(q_dev, k), reps=5), sec * 1e3, f"{sq} of {nq} queries", nq / sq)
    vals, idx = index.search(q_dev, k)
    sm = 2048
    v_np, i_np = vals[:sm].cpu().numpy(), idx[:sm].cpu().numpy()
    emit(name, "mAP / P@k / R@k over the hit lists, Jaccard > 0.4 (evaluate_results, evaluate_nih_zilliz.py:34-64)",
         _gpu_ms(lambda: M.evaluate_results_from_topk(vals, idx, qlab, glab, 0.4, (1, 5, 10, 20, 50)), reps=5),
         _cpu_ms(lambda: rm.evaluate_results(v_np, i_np, mlab[:sm], mlab[nq:], 0.4, (1, 5, 10, 20, 50))),
         f"{sm} of {nq} queries", nq / sm)
    emit(name, "Precision@K / hit-rate, shares >= 1 label (test.py:1031-1056)",
         _gpu_ms(lambda: M.multilabel_hit_rate_from_topk(idx, qlab, glab, (1, 5, 10, 15, 20)), reps=5),
         _cpu_ms(lambda: rm.multilabel_precision_recall_at_k(i_np, mlab[:sm], mlab[nq:], (1, 5, 10, 15, 20))),
         f"{sm} of {nq} queries", nq / sm)
