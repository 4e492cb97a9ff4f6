#!/usr/bin/env python
"""Measurement of the paths bench.py's headline line does not cover (one JSON line each, CUDA-event timed):

  exact-fp32 mode   C3 (25000 q x 112000 g x 1024, top-50)  and C1 / C2 (launch-latency bound)  -- FP32-FFMA bound
  hamming           1024 q x 10M codes x 64 bit, top-100                                         -- issue / HBM bound
  score fusion      4096 x 4096 self-retrieval over two embedding sets, zscore                   -- fp32 path + stats

    python tools/bench_side_paths.py [--steps 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, steps, warmup=3):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import pynvml
    import torch

    import b200knn

    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    sms = torch.cuda.get_device_properties(0).multi_processor_count

    def clock():
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)

    # ---- exact fp32 ------------------------------------------------------------------------------------------
    import importlib

    S = importlib.import_module("b200knn.search")
    for name, nq, ng, d, k, metric, engine in (("c1", 400, 400, 1024, 10, "cosine", "ffma"),
                                               ("c2", 600, 2000, 256, 10, "l2", "ffma"),
                                               ("c3", 25000, 112000, 1024, 50, "cosine", "ffma"),
                                               ("c3", 25000, 112000, 1024, 50, "cosine", "tensor"),
                                               ("c3", 25000, 112000, 1024, 50, "l2", "tensor")):
        g = b200knn.normalize(torch.randn((ng, d), generator=gen, device=dev))
        q = g if name == "c1" else b200knn.normalize(torch.randn((nq, d), generator=gen, device=dev))
        os.environ["KNN_EXACT_ENGINE"] = engine
        assert S.exact_engine(nq, ng, d, k) == engine
        index = b200knn.FlatIndex(d, metric, "fp32").adopt(g)     # the tensor engine keeps its split rows here
        mhz = []

        def step():
            out = index.search(q, k, exclude_self=(name == "c1"))
            mhz.append(clock())
            return out

        ms = timed(step, args.steps if name == "c3" else 50)
        flops = 2.0 * nq * ng * d
        clk = sorted(mhz)[len(mhz) // 2]
        if engine == "ffma":
            peak = sms * 128 * 2 * clk * 1e6 / 1e12          # FFMA lanes x 2 FLOP x observed SM clock
            roof = {"bound": "fp32 FFMA", "achieved": flops / (ms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "frac": flops / (ms / 1e3) / 1e12 / peak,
                    "peak_source": f"{sms} SMs x 128 lanes x 2 FLOP x {clk} MHz observed (nominal; no measured fp32 peak)"}
            kern, note = "search_f32_kernel", "whole call (normalised fp32 rows resident): distance+select kernel + unit merge"
        else:
            peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                               "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
            roof = {"bound": "tensor", "achieved": flops / (ms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "frac": flops / (ms / 1e3) / 1e12 / peak, "mma_TFLOPs": 3 * flops / (ms / 1e3) / 1e12,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained; every product counted once, the filter "
                                   "issues 3 bf16 MMAs per product"}
            kern = "search_bf16_pair_kernel<kSplit> + rescore_exact_stream_kernel"
            note = ("whole call (normalised fp32 rows + their bf16x3 split resident): split of the queries, filter, "
                    f"exact re-scoring + proof; {S._search_exact_tensor.last_unverified} queries re-run on FFMA")
        line = {"path": f"exact-fp32/{engine}", "workload": f"{name}: {nq} x {ng} x {d} fp32, top-{k}, {metric}",
                "ms_per_call": ms, "queries_per_s": nq / (ms / 1e3), "kernel": kern, "sm_mhz": clk, "roofline": roof,
                "note": note}
        print(json.dumps(line), flush=True)
    os.environ.pop("KNN_EXACT_ENGINE", None)

    # ---- hamming ---------------------------------------------------------------------------------------------
    nq, ng, bits, k = 1024, 10_000_000, 64, 100
    gw = torch.randint(-(2 ** 62), 2 ** 62, (ng, 1), generator=gen, device=dev, dtype=torch.int64)
    qw = torch.randint(-(2 ** 62), 2 ** 62, (nq, 1), generator=gen, device=dev, dtype=torch.int64)
    for nqq in (1, 128, 1024):
        for method in ("popc", "mma"):
            ms = timed(lambda: b200knn.search_hamming(qw[:nqq], gw, k, packed=True, method=method), args.steps)
            print(json.dumps({
                "path": f"hamming/{method}", "workload": f"{nqq} x {ng} x {bits}-bit codes, top-{k}", "ms_per_call": ms,
                "queries_per_s": nqq / (ms / 1e3),
                "kernel": "search_hamming_kernel" if method == "popc" else "unpack_pm1 + tcgen05 bf16 search",
                "pair_rate_G_per_s": nqq * ng / (ms / 1e3) / 1e9,
                "note": ("integer xor+popc over packed words, one thread per query row; issue-bound for >= 128 queries"
                         if method == "popc" else
                         "codes as +-1 bf16 rows (16x the gallery bytes, expanded inside the timed call): "
                         "<q, g> = bits - 2 d exactly")}), flush=True)

    # ---- score fusion ----------------------------------------------------------------------------------------
    n = 4096
    conv = torch.randn((n, 1024), generator=gen, device=dev)
    dino = torch.randn((n, 768), generator=gen, device=dev)
    for mode in ("none", "zscore"):
        ms = timed(lambda: b200knn.fusion.score_fusion_search(conv, dino, 0.5, 10, mode), args.steps)
        print(json.dumps({"path": "score-fusion", "workload": f"{n} x {n} self-retrieval, 1024-d + 768-d, top-10, {mode}",
                          "ms_per_call": ms, "queries_per_s": n / (ms / 1e3),
                          "note": "normalise x2 (+ score statistics x2) + one fused fp32 search over 1792-d"}), flush=True)


if __name__ == "__main__":
    main()
