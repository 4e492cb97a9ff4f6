"""Build the C-ABI shared library (libb200knn.so) in-tree with nvcc for sm_100a.

    python "image-retrieval---thesis-2026_b200/build.py" [--force]

The library is the product: there is no Python/torch fallback for any kernel.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libb200knn.so")
SOURCES = ["api.cu", "normalize.cu", "search_f32.cu", "search_tc.cu", "search_tc2.cu", "search_ts.cu", "search_hamming.cu", "exact_tc.cu", "rerank.cu", "merge.cu", "metrics.cu", "rank_positives.cu", "pairwise.cu", "small.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_mtime() -> float:
    newest = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in os.listdir(root):
            if name.endswith((".cu", ".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


CHECK_LIB = os.path.join(HERE, "libb200knn_check.so")


def build(force: bool = False, verbose: bool = False, check: bool = False) -> str:
    """check=True builds libb200knn_check.so with -DKNN_BOUNDS_CHECK (device-side asserts on the candidate-list
    invariants; select it with KNN_LIB=<path>): compute-sanitizer is not available on the GPU pool."""
    if check:
        return _build_variant(CHECK_LIB, os.path.join(CSRC, "build_check"), ["-DKNN_BOUNDS_CHECK"], verbose)
    if not force and not needs_build():
        return LIB
    return _build_variant(LIB, BUILD, [], verbose, force)


def _build_variant(LIB: str, BUILD: str, extra_flags, verbose: bool, force: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    hdr_mtime = max(
        os.path.getmtime(os.path.join(r, n))
        for r in (CSRC, os.path.join(HERE, "..", "include"))
        for n in os.listdir(r)
        if n.endswith((".cuh", ".h"))
    )

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), hdr_mtime):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", srcp, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, check="--check" in sys.argv)
    print(path)
