#!/usr/bin/env python
"""bench.py -- queries/sec @ top-100 of the exact k-NN hot path on 1/2/4/8 B200 (BASELINE.json metric).

Workload (BASELINE config 5): 8192 queries x 50M-vector 512-d bf16 gallery, top-100.  The 50M-row gallery
(51.2 GB) fits one B200, so the SAME total work runs at every N ("strong" scaling): the gallery is row-sharded
over the N ranks (rank r keeps rows [r*N/W, (r+1)*N/W) resident in its HBM, generated on the device), queries
are replicated, every rank runs the fused tcgen05 distance + top-k over its shard, then ONE all-gather of the
[Q,100] candidates + an on-device k-way merge.  A step = one full search of the query batch.

  value : queries/s with the (normalised bf16) queries already resident in HBM
  e2e   : queries/s through the public API from HOST buffers: pinned fp32 queries -> H2D -> fused
          normalise+cast -> search -> (all-gather + merge) -> D2H of distances + indices
  roofline : tensor-pipe roofline of the dominant kernel (search_bf16_pair_kernel, csrc/search_tc2.cu): 2*Q*N_shard*D
             FLOP per launch over its CUDA-event duration (events recorded inside knn_search on the launching stream)
  parity_check : after the timed regions, recall@k of the timed search against the exact-fp32 engine run over the WHOLE
             gallery for 32 sampled queries (and, at N > 1, that every rank holds the identical merged result)
  secondary : the same measurement for c4 (64 x 10 M x 768 bf16: HBM roofline) and c3-fp32 (25 000 x 112 000 x 1024, exact
             fp32 mode), shorter runs in the same process
  cpu_baseline / --impl reference : the reference's own CPU path (F.normalize -> mm -> topk, all host threads)
             on a bounded sample, extrapolated linearly in queries x gallery rows.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (queries, gallery rows, dim, k)
    "c5": (8192, 50_000_000, 512, 100),
    "c4": (64, 10_000_000, 768, 100),
    "c3": (25_000, 112_000, 1024, 50),
    # exact fp32 mode (BASELINE config 3 as quoted: bit-exact indices / metrics), tensor-core exact engine
    "c3-fp32": (25_000, 112_000, 1024, 50),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--gallery-rows", type=int, default=0, help="override the gallery size (debug only)")
    ap.add_argument("--queries", type=int, default=0, help="override the query batch (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the two secondary workloads (c4: HBM-bound small batch, c3-fp32: exact mode) that the "
                         "default c5 run measures after the headline one")
    ap.add_argument("--stages", action="store_true",
                    help="instead of the headline line: per-stage GPU vs host-CPU timings (normalise / distance / "
                         "ranking / each metric) for BASELINE configs 1-3, one JSON line each (bench_stages.py)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "allgather"],
                    help="N > 1: candidate exchange -- merge kernel reading the peers' symmetric-memory buffers over "
                         "NVLink (falls back to the all-gather if symmetric memory is unavailable), or NCCL all-gather")
    ap.add_argument("--pipeline", default="on", choices=["on", "off"],
                    help="N > 1, peer exchange: run the exchange of step i (wait for the slowest shard + merge over NVLink) "
                         "on a side stream under the local search of step i + 1 (ShardedFlatIndex.search_async); every "
                         "step's result is still taken -- one step later -- inside the timed region")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"tflops": float(p["bf16_tflops_sustained"]), "hbm_gbs": float(p["hbm_gbs"]),
                "source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


def traffic_bytes(workload, nq, rows, d):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture of this very configuration (profiles/traffic.json); None when it was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        table = json.load(fh)
    ent = table.get(f"{workload}:{nq}x{rows}x{d}")
    return None if ent is None else ent["dram_bytes_per_launch"]


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML every 100 ms while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_numbers(nq_total, ng_total, d, k, steps=None, warmup=1):
    """Bounded sample of the workload on the host cores -> (q/s extrapolated to the full gallery, dict).
    steps=None: about 10 s of CPU work (the sample pass repeated, at most 20 times) -- the default arm's cpu_baseline;
    the --impl reference arm passes its own --steps / --warmup."""
    from oracle import cpu_baseline

    sq, sg = min(nq_total, 1024), min(ng_total, 1_000_000)
    if steps is None:
        first, threads = cpu_baseline.time_reference(sq, sg, d, k, steps=1, warmup=warmup)
        steps = max(1, min(20, int(10.0 / max(first, 1e-3))))
        warmup = 0
    sec, threads = cpu_baseline.time_reference(sq, sg, d, k, steps=steps, warmup=warmup)
    factor = ng_total / sg
    qps = sq / (sec * factor)
    info = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{sq} queries x {sg} rows x {d}-d fp32, F.normalize+mm+topk({k}) (test.py:1005-1006,44): "
                      f"{sec:.3f} s per pass, mean of {steps} passes; x{factor:g} rows extrapolated linearly"}
    return qps, sec, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nq, ng, d, k = WORKLOADS[args.workload]
    nq, ng = args.queries or nq, args.gallery_rows or ng
    t0 = time.perf_counter()
    qps, sec, info = cpu_reference_numbers(nq, ng, d, k, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": "queries/sec @top-100", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": nq / qps * 1e3, "ms_per_step_is_extrapolated": True, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {nq} queries x {ng} gallery x {d}-d, top-{k}",
                   "note": "reference CPU path (torch F.normalize + mm + topk) on a bounded sample, extrapolated"},
        "cpu_baseline": info,
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


GALLERY_SEED = 1234567
GEN_CHUNK = 1 << 20


def _source_rows(gen, start, s, e, d, dev):
    """The fp32 rows [start+s, start+e) of the synthetic gallery BEFORE normalisation / the cast to the search dtype
    (one seeded stream per 1 Mi-row chunk, so any chunk can be regenerated)."""
    import torch

    gen.manual_seed(GALLERY_SEED + start + s)
    return torch.randn((e - s, d), generator=gen, device=dev, dtype=torch.float32)


def parity_check(b200knn, dist, world, rank, dev, rows, start, count, d, k, q_src, q_dev, result, exact_mode):
    """Correctness of the TIMED configuration, inside the bench run.  For 32 sampled queries the exact-fp32 engine
    (FFMA kernel, bit-exact against the CPU oracle in the test-suite) searches the whole gallery chunk by chunk:
      * over the STORED rows upcast to fp32 with the stored queries -> what an exact selection over the kernel's own
        inputs returns (gate: recall@k >= 0.999; differences can only be fp32 summation-order near-ties);
      * over the fp32 SOURCE rows (regenerated) with the fp32 queries -> recall against the fp32 reference including
        the bf16 rounding of the stored gallery (reported; on i.i.d. Gaussian rows the scores at rank k are ~1e-4 apart).
    At N > 1 every rank searches its shard, the [32, k] lists are all-gathered and merged, and the merged result of the
    timed search must be identical on every rank (checksum)."""
    import torch

    nq = q_dev.shape[0]
    ns = min(32, nq)
    os.environ["KNN_EXACT_ENGINE"] = "ffma"     # the referee is the FFMA kernel, never the engine under test
    sel = torch.linspace(0, nq - 1, ns, device=dev).long()
    val, idx = result
    stored_q = q_dev[sel].float()
    source_q = b200knn.normalize(q_src[sel])
    gen = torch.Generator(device=dev)
    best = {"stored": None, "source": None}

    def fold(key, v, i):
        if best[key] is None:
            best[key] = (v, i)
        else:
            best[key] = b200knn.merge_topk(torch.stack([best[key][0], v]), torch.stack([best[key][1], i]), "cosine")

    for s in range(0, count, GEN_CHUNK):
        e = min(count, s + GEN_CHUNK)
        chunk = rows[s:e].float()                                   # exact upcast of the stored rows
        ix = b200knn.FlatIndex(d, "cosine", "fp32", index_base=start + s, device=dev).adopt(chunk)
        fold("stored", *ix.search(stored_q, min(k, e - s)) if e - s >= k else _pad(ix.search(stored_q, e - s), k))
        if not exact_mode:
            src = b200knn.normalize(_source_rows(gen, start, s, e, d, dev))
            ix = b200knn.FlatIndex(d, "cosine", "fp32", index_base=start + s, device=dev).adopt(src)
            fold("source", *ix.search(source_q, k) if e - s >= k else _pad(ix.search(source_q, e - s), k))
        del chunk, ix
    out = {"queries_sampled": ns, "k": k, "gate": 0.999}
    for key in ("stored", "source"):
        if best[key] is None:
            continue
        v, i = best[key]
        if world > 1:                                               # merge the per-shard exact lists
            gv = torch.empty((world,) + tuple(v.shape), dtype=v.dtype, device=dev)
            gi = torch.empty((world,) + tuple(i.shape), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(gv, v.contiguous())
            dist.all_gather_into_tensor(gi, i.contiguous())
            v, i = b200knn.merge_topk(gv, gi, "cosine")
        got_i, got_v = idx[sel], val[sel]
        hit = (got_i.unsqueeze(2) == i.unsqueeze(1)).any(dim=2)     # [ns, k]: returned row is in the exact top-k
        recall = hit.float().mean(dim=1)
        name = "exact_fp32_same_inputs" if key == "stored" else "fp32_source_rows"
        out[f"recall_at_k_{name}"] = float(recall.mean().item())
        out[f"min_query_recall_{name}"] = float(recall.min().item())
        if key == "stored":
            out["identical_indices_queries"] = int((got_i == i).all(dim=1).sum().item())
            out["max_abs_score_diff"] = float((got_v - v).abs().max().item())
    ordered = bool((val[:, 1:] <= val[:, :-1]).all().item())
    uniq = bool(all(len(set(r)) == len(r) for r in idx[sel].tolist()))
    out["sorted_best_first"], out["unique_indices"] = ordered, uniq
    if world > 1:
        h = torch.stack([idx.double().sum(), (idx.double() * torch.arange(1, idx.shape[1] + 1, device=dev)).sum(),
                         val.double().sum()])
        allh = torch.empty((world, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allh, h)
        out["cross_rank_identical"] = bool((allh == allh[0]).all().item())
    out["ok"] = bool(out["recall_at_k_exact_fp32_same_inputs"] >= out["gate"] and ordered and uniq
                     and out.get("cross_rank_identical", True))
    os.environ.pop("KNN_EXACT_ENGINE", None)
    return out


class PendingNow:
    """A result that is already complete in stream order (same interface as sharded.PendingSearch)."""

    def __init__(self, vals, idx):
        self.vals, self.idx = vals, idx

    def result(self):
        return self.vals, self.idx


def _pad(res, k):
    import torch

    v, i = res
    pv = torch.full((v.shape[0], k), float("-inf"), dtype=v.dtype, device=v.device)
    pi = torch.full((i.shape[0], k), -1, dtype=i.dtype, device=i.device)
    pv[:, : v.shape[1]], pi[:, : i.shape[1]] = v, i
    return pv, pi


def run_workload(name, args, ctx, steps, warmup, primary):
    """One workload on this rank's shard -> the JSON object (rank 0 prints the primary one; the others ride in its
    "secondary" field)."""
    import torch

    import b200knn
    from b200knn import _lib
    from b200knn.sharded import ShardedFlatIndex, shard_rows

    dist, world, rank, dev, lib, exchange = ctx["dist"], ctx["world"], ctx["rank"], ctx["dev"], ctx["lib"], ctx["exchange"]
    nq, ng, d, k = WORKLOADS[name]
    if primary:
        nq, ng = args.queries or nq, args.gallery_rows or ng
    start, count = shard_rows(ng, world)[rank]
    exact = name.endswith("-fp32")
    precision = "fp32" if exact else "bf16"
    store_dtype = torch.float32 if exact else torch.bfloat16

    # ---- gallery shard: generated on the device chunk by chunk, fused normalise + cast ---------------------
    rows = torch.empty((count, d), dtype=store_dtype, device=dev)
    gen = torch.Generator(device=dev)
    for s in range(0, count, GEN_CHUNK):
        e = min(count, s + GEN_CHUNK)
        rows[s:e] = b200knn.normalize(_source_rows(gen, start, s, e, d, dev), out_dtype=store_dtype)
    index = b200knn.FlatIndex(d, "cosine", precision, normalize=True, index_base=start, device=dev).adopt(rows)
    pipeline = bool(ctx.get("pipeline")) and world > 1 and exchange == "peer"
    sharded = ShardedFlatIndex(index, exchange=exchange, pipeline=pipeline)

    # ---- queries: i.i.d. Gaussian rows (a flat score distribution is the worst case for the fused selection),
    # replicated on every rank -------------------------------------------------------------------------------
    gen.manual_seed(99)
    qsrc = torch.randn((nq, d), generator=gen, device=dev, dtype=torch.float32)
    q_host = qsrc.cpu().pin_memory()                      # the user's host buffer (fp32)
    q_dev = b200knn.normalize(qsrc, out_dtype=store_dtype)
    index_prepared = b200knn.FlatIndex(d, "cosine", precision, normalize=False, index_base=start, device=dev).adopt(rows)
    sharded_prepared = ShardedFlatIndex(index_prepared, exchange=exchange, pipeline=pipeline)
    out_val_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    out_idx_host = torch.empty((nq, k), dtype=torch.int64).pin_memory()

    # Pipelined exchange (N > 1): a step enqueues its search and TAKES the result of the previous step (the consumer
    # runs one step behind); `drain` takes the last one, inside the timed region.
    pending = []

    last_copy = [None]

    def take(p, to_host):
        if to_host and hasattr(p, "to_host"):
            # pipelined exchange: the device->host copies ride on the side stream behind the merge (copy engine, under
            # the next local search); `drain` makes the timed stream wait for the last of them
            last_copy[0] = p.to_host(out_val_host, out_idx_host)
            return p.vals, p.idx
        v, i = p.result()
        if to_host:
            out_val_host.copy_(v, non_blocking=True)
            out_idx_host.copy_(i, non_blocking=True)
        return v, i

    def step_resident():
        # queries already normalised / cast in HBM: the index over the same rows with normalize=False skips the
        # normalisation FlatIndex.search would otherwise repeat; N > 1 adds the candidate exchange + shard merge
        if not pipeline:
            return sharded_prepared.search(q_dev, k)
        pending.append((sharded_prepared.search_async(q_dev, k), False))
        return take(*pending.pop(0)) if len(pending) > 1 else None

    def step_e2e():
        # public API on a HOST batch: this rank copies its 1/N slice, normalises + casts it, the prepared slices are
        # all-gathered over NVLink; then search, candidate exchange, shard merge, results back to the host
        if not pipeline:
            return take(PendingNow(*sharded.search_host(q_host, k)), True)
        pending.append((sharded.search_host_async(q_host, k), True))
        return take(*pending.pop(0)) if len(pending) > 1 else None

    def drain():
        res = None
        while pending:
            res = take(*pending.pop(0))
        if last_copy[0] is not None:      # the closing event of the timed region comes after the last host copy
            torch.cuda.current_stream(dev).wait_event(last_copy[0])
            last_copy[0] = None
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            fn()
        drain()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(warmup, 3)):
        step_resident()
    drain()
    for _ in range(2):
        step_e2e()
    drain()

    # ---- timed region: device-resident, clocks sampled; the library records CUDA events around its kernels on the
    # launching stream during these very steps (no synchronisation), read back after the region has ended ----------
    lib.knn_profile_enable(1)
    sharded_prepared.profile(world > 1)
    launches0 = lib.knn_launch_count()
    with ClockSampler(ctx["local_rank"]) as clocks:
        total_ms = timed(step_resident, steps)
    xprof = sharded_prepared.profile_read()[-steps:] if world > 1 else []
    sharded_prepared.profile(False)
    launches = lib.knn_launch_count() - launches0            # counted by the library at every launch site
    kern_ms = []
    n_prof = lib.knn_profile_count()
    for i in range(max(0, n_prof - min(steps, 48)), n_prof):     # the library keeps the last 64 calls
        sd, a, b = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        _lib.check(lib.knn_profile_read(i, ctypes.byref(sd), ctypes.byref(a), ctypes.byref(b)), "knn_profile_read")
        kern_ms.append((a.value, b.value, sd.value))
    lib.knn_profile_enable(0)
    e2e_ms = timed(step_e2e, steps)

    ms_per_step = total_ms / steps
    value = nq / (ms_per_step / 1e3)
    e2e_value = nq / (e2e_ms / steps / 1e3)
    peaks = load_peaks()
    dist_ms = sum(a for a, _, _ in kern_ms) / len(kern_ms)     # the dominant kernel alone
    merge_ms = sum(b for _, b, _ in kern_ms) / len(kern_ms)
    seed_ms = sum(c for _, _, c in kern_ms) / len(kern_ms)     # threshold-seeding pre-pass + seeding merge
    flops = 2.0 * nq * count * d                               # algorithmic: every (q, g, d) product counted once
    achieved = flops / (dist_ms / 1e3) / 1e12
    esize = 4 if exact else 2
    gallery_gbs = count * d * esize / (dist_ms / 1e3) / 1e9
    # regime: arithmetic intensity 2Q/2 FLOP/B for bf16 against the measured ridge
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    hbm_bound = nq < ridge
    algo_bytes = count * d * esize + nq * d * esize + nq * k * 12
    # dispatch rule of knn_search (csrc/api.cu): one 128-row query block and d <= 768 -> TMEM-resident-query kernel,
    # several query blocks -> CTA-pair kernel
    kernel_name = ("search_bf16_ts_kernel" if d <= 768 else "search_bf16_kernel") if nq <= 128 else "search_bf16_pair_kernel"
    if exact:
        kernel_name = "search_bf16_pair_kernel<kSplit> (bf16 split filter of the exact mode)" if nq > 128 else kernel_name

    line = {
        "metric": "queries/sec @top-100", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32" if exact else "bf16",
        "data": "synthetic",
        "config": {
            "workload": f"{name}: {nq} queries x {ng} gallery x {d}-d {precision}, top-{k}, cosine",
            "gallery_rows_per_gpu": count, "sharding": f"rows/{world}",
            "exchange": "none" if world == 1 else exchange,
            "pipeline": ("exchange of step i on a side stream under the local search of step i+1; results taken one step "
                         "later, the last one before the closing event") if pipeline else "off", "l2_flush": "inputs larger than L2 "
            f"({count * d * esize / 1e9:.1f} GB gallery shard streamed per step)",
        },
        "roofline": ({
            "bound": "hbm", "achieved": algo_bytes / (dist_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": algo_bytes / (dist_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            "traffic": traffic_bytes(name, nq, count, d), "algorithmic_bytes": algo_bytes,
            "kernel": kernel_name, "kernel_ms": dist_ms, "seeding_ms": seed_ms,
            "merge_kernel_ms": merge_ms, "tensor_TFLOPs": achieved, "peak_source": "MEASURED_PEAKS.json hbm_gbs",
        } if hbm_bound else {
            "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tflops"], "traffic": traffic_bytes(name, nq, count, d),
            "kernel": kernel_name, "seeding_ms": seed_ms,
            "kernel_ms": dist_ms, "merge_kernel_ms": merge_ms, "gallery_stream_GBps": gallery_gbs,
            "algorithmic_bytes": algo_bytes,
            "peak_source": peaks["source"],
        }),
        # every rank copies its 1/N slice of the fp32 batch in and the whole result out
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": ((nq + world - 1) // world) * d * 4,
                "d2h_bytes_per_step": nq * k * 12, "ms_per_step": e2e_ms / steps,
                "note": "bytes per rank; the prepared query slices travel between the GPUs over NVLink (all-gather)"},
        # per step: seeding pre-pass + seeding merge + distance/selection kernel + unit merge
        # (+ k-way shard merge after the exchange)
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if not exact:   # work decomposition of this rank's search (host arithmetic of the library, knn_search_geometry)
        geo = (ctypes.c_int64 * 8)()
        if lib.knn_search_geometry(nq, count, d, 1, k, geo) == 0:
            line["roofline"]["decomposition"] = {
                "query_blocks": geo[0], "gallery_splits": geo[1], "lists_per_row": geo[1] * geo[2],
                "seeding_sample_rows": geo[5] * geo[6],
                "seeding": ("none" if geo[5] == 0 else "chunk maxima" if geo[7] > 0 else "selecting pre-pass")}
    if exact:
        import importlib

        S = importlib.import_module("b200knn.search")
        nprod = 2 if (S._two_product_filter(nq, count, "keep") and k * 2 + 16 <= 256) else 3
        line["roofline"]["note"] = (
            f"exact fp32 mode on the tensor cores: the dominant kernel is the bf16 split filter ({nprod} MMAs per product, "
            "csrc/search_tc2.cu kSplit; two products = q_hi.g_hi + q_lo.g_hi under the wide bound |q| max|g_lo|, queries "
            "it cannot prove are re-run with three); 'achieved' counts every product ONCE (SURVEY 8d), mma_TFLOPs is the "
            "bf16 tensor work actually issued; the exact fp32 re-scoring + proof kernel runs after it; kernel_ms averages "
            "the profiled knn_search calls of the timed steps (re-runs of unproven queries, if any, included)")
        line["roofline"]["mma_TFLOPs"] = nprod * achieved
        line["roofline"]["mma_frac_of_peak"] = nprod * achieved / peaks["tflops"]
        line["config"]["filter_products"] = nprod
        line["config"]["two_product_queries_rerun_with_three"] = S._search_exact_tensor.last_two_product_rerun
        line["config"]["unverified_queries_rerun_on_ffma"] = S._search_exact_tensor.last_unverified
    if world > 1:  # per-rank view of the same timed region: which rank the max-over-ranks step time waits for
        n = max(len(xprof), 1)
        local_ms = sum(x[0] for x in xprof) / n            # local search (seeding + kernel + unit merge) of a step
        wait_ms = sum(x[1] for x in xprof) / n             # all-gather exchange: the NCCL all-gather; peer exchange: ~0
        xmerge_ms = sum(x[2] for x in xprof) / n           # peer exchange: wait for the slowest shard + merge over NVLink
        mine = torch.tensor([dist_ms, seed_ms, merge_ms, float(clocks.summary()["sm_mhz"] or 0), local_ms, wait_ms,
                             xmerge_ms], device=dev)
        allr = torch.empty((world, 7), device=dev)
        dist.all_gather_into_tensor(allr, mine)
        line["per_rank"] = {"kernel_ms": [round(x, 3) for x in allr[:, 0].tolist()],
                            "seeding_ms": [round(x, 3) for x in allr[:, 1].tolist()],
                            "unit_merge_ms": [round(x, 3) for x in allr[:, 2].tolist()],
                            "sm_mhz": [int(x) for x in allr[:, 3].tolist()],
                            "local_search_ms": [round(x, 3) for x in allr[:, 4].tolist()],
                            "allgather_ms": [round(x, 3) for x in allr[:, 5].tolist()],
                            "exchange_wait_plus_merge_ms": [round(x, 3) for x in allr[:, 6].tolist()]}
    # ---- correctness of this very configuration (after the timed regions) ------------------------------------
    try:
        result = sharded_prepared.search(q_dev, k)
        line["parity_check"] = parity_check(b200knn, dist, world, rank, dev, rows, start, count, d, k, qsrc, q_dev,
                                            result, exact)
    except Exception as exc:  # noqa: BLE001 - reported, never hidden
        line["parity_check"] = {"ok": False, "error": repr(exc)}
    if primary and rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, _, info = cpu_reference_numbers(nq, ng, d, k)
            line["cpu_baseline"] = info
        except Exception as exc:  # the GPU number stands on its own
            line["cpu_baseline"] = {"error": repr(exc)}
    del rows, index, index_prepared, sharded, sharded_prepared
    torch.cuda.empty_cache()
    return line


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.stages:
        import bench_stages

        bench_stages.run_stages()
        return

    import torch
    import torch.distributed as dist

    import b200knn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = b200knn.load_library()
    exchange = args.exchange
    if world > 1 and exchange == "peer":
        try:  # collective: every rank succeeds or every rank falls back (the probe result is all-reduced)
            from b200knn.sharded import PeerExchange

            probe = PeerExchange(None, dev)
            probe.slot(8, 8)
            ok = torch.ones(1, device=dev)
        except Exception as exc:  # noqa: BLE001 - any failure means "no symmetric memory here"
            if rank == 0:
                print(f"[bench] symmetric memory unavailable ({exc!r}); using the all-gather exchange", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            exchange = "allgather"
    ctx = {"dist": dist, "world": world, "rank": rank, "local_rank": local_rank, "dev": dev, "lib": lib,
           "exchange": exchange, "pipeline": args.pipeline == "on"}

    line = run_workload(args.workload, args, ctx, args.steps, args.warmup, primary=True)
    # the other two regimes the metric names, measured in the same driver-run process: HBM-bound small batch (c4) and
    # exact fp32 (c3-fp32); shorter runs, same timing rules
    if not args.no_secondary and args.workload == "c5" and not (args.queries or args.gallery_rows):
        line["secondary"] = {}
        for name in ("c4", "c3-fp32"):
            try:
                sec = run_workload(name, args, ctx, max(5, min(args.steps, 20)), 3, primary=False)
                line["secondary"][name] = {key: sec[key] for key in ("value", "unit", "ms_per_step", "dtype", "config",
                                                                     "roofline", "e2e", "gpu_launches", "clocks",
                                                                     "parity_check", "steps", "warmup") if key in sec}
            except Exception as exc:  # noqa: BLE001
                line["secondary"][name] = {"error": repr(exc)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
