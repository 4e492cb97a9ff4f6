"""The reference's CPU search path, timed on the host cores -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Restates what the reference executes for a query batch against a gallery (its own torch primitives, nothing
faster substituted):  F.normalize (test.py:1005) -> `q @ g.T` (test.py:1006 / train.py:405) ->
`S.topk(k, dim=1, largest=True, sorted=True)` (test.py:44), all fp32 on CPU with every host thread.
bench.py calls this for the `cpu_baseline` object and for the `--impl reference` arm, on a BOUNDED sample of
the workload, and extrapolates linearly in (queries x gallery rows).
"""
from __future__ import annotations

import os
import time

import torch
import torch.nn.functional as F


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_sample(nq: int, ng: int, d: int, seed: int = 5):
    gen = torch.Generator().manual_seed(seed)
    q = torch.randn((nq, d), generator=gen, dtype=torch.float32)
    g = F.normalize(torch.randn((ng, d), generator=gen, dtype=torch.float32), p=2, dim=1)
    return q, g


def reference_search(q: torch.Tensor, g_normalized: torch.Tensor, k: int):
    """One pass of the reference hot path on CPU tensors -> (values, indices)."""
    qn = F.normalize(q, p=2, dim=1)
    sim = torch.mm(qn, g_normalized.t())
    return sim.topk(k, 1, True, True)


def time_reference(nq: int, ng: int, d: int, k: int, steps: int = 1, warmup: int = 1):
    """-> (seconds per pass over the sample, threads used)."""
    threads = host_threads()
    torch.set_num_threads(threads)
    q, g = make_sample(nq, ng, d)
    for _ in range(warmup):
        reference_search(q, g, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_search(q, g, k)
    return (time.perf_counter() - t0) / steps, threads
