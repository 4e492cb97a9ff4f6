#!/usr/bin/env python
"""Stall-cycle breakdown of the tcgen05 search kernels (diagnostics; needs KNN_PAIR_STATS=1 in the environment).

    KNN_PAIR_STATS=1 python tools/diag_stalls.py --queries 8192 --rows 20000000 --dim 512 --k 100

Prints, per kernel role, the share of its lifetime spent waiting on the other roles:
  MMA issuer   : waiting for a free TMEM stage (selection behind) / for a filled smem slot (TMA behind)
  TMA producer : waiting for a free smem slot
  selection    : waiting for an accumulator tile; time inside chunks that left the fast path; final compaction
"""
import argparse
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=8192)
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch

    import b200knn
    from b200knn import _lib
    from b200knn.search import _search_prepared

    lib = b200knn.load_library()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    rows = torch.empty((args.rows, args.dim), dtype=torch.bfloat16, device=dev)
    for s in range(0, args.rows, 1 << 20):
        e = min(args.rows, s + (1 << 20))
        gen.manual_seed(1234567 + s)
        rows[s:e] = b200knn.normalize(torch.randn((e - s, args.dim), generator=gen, device=dev), out_dtype=torch.bfloat16)
    gen.manual_seed(99)
    q = b200knn.normalize(torch.randn((args.queries, args.dim), generator=gen, device=dev), out_dtype=torch.bfloat16)

    def step():
        return _search_prepared(q, None, rows, None, args.k, "cosine", "keep", 0, 0)

    for _ in range(2):
        step()
    out = (ctypes.c_ulonglong * 32)()
    have_stats = os.environ.get("KNN_PAIR_STATS") == "1"
    if have_stats:
        _lib.check(lib.knn_debug_stats(out, 1), "knn_debug_stats")
    torch.cuda.synchronize()
    import threading

    import pynvml

    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            stop.wait(0.02)

    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    stop.set()
    th.join()
    if samples:
        mhz = sorted(c for c, _ in samples)
        pw = sorted(w for _, w in samples)
        print(f"  nvml: SM clock median {mhz[len(mhz) // 2]} MHz (min {mhz[0]}, max {mhz[-1]}), power median {pw[len(pw) // 2]:.0f} W "
              f"(max {pw[-1]:.0f} W), {len(samples)} samples")
    if have_stats:
        _lib.check(lib.knn_debug_stats(out, 1), "knn_debug_stats")
    s = [int(x) for x in out]
    mma_total, mma_tmem, mma_full, epi_total, epi_wait, epi_slow, n_slow, prod_empty, n_mma, n_epi, n_chunks, fin = s[:12]
    tf = 2.0 * args.queries * args.rows * args.dim / (ms / 1e3) / 1e12
    gbs = args.rows * args.dim * 2 / (ms / 1e3) / 1e9
    print(f"q={args.queries} rows={args.rows} d={args.dim} k={args.k}: {ms:.3f} ms/step  {tf:.1f} TFLOP/s  {gbs:.0f} GB/s gallery")
    if n_mma:
        print(f"  MMA issuer ({n_mma} threads): lifetime {mma_total / n_mma / 1e6:.2f} Mcyc each; "
              f"wait TMEM stage {100 * mma_tmem / mma_total:.1f}%  wait smem slot {100 * mma_full / mma_total:.1f}%")
        print(f"  TMA producer: wait free slot {100 * prod_empty / max(1, mma_total):.1f}% (of MMA lifetime, both CTAs summed)")
    if n_epi:
        print(f"  selection ({n_epi} warps): wait accumulator {100 * epi_wait / epi_total:.1f}%  "
              f"slow chunks {100 * epi_slow / epi_total:.1f}% of time, {100 * n_slow / max(1, n_chunks):.2f}% of chunks, "
              f"{epi_slow / max(1, n_slow):.0f} cyc each; final compaction {100 * fin / epi_total:.1f}% extra")


if __name__ == "__main__":
    main()
