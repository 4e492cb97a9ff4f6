"""CPU oracle for the retrieval hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package, and only as the checker / the timed CPU baseline.  The product package
(``image-retrieval---thesis-2026_b200``) never imports it and has no CPU fallback.

``knn_oracle.c`` restates normalise -> similarity/distance -> self-mask -> stable top-k (reference file:line in
its header); ``reference_metrics.py`` restates the reference's metric functions.  Both are pinned against the
REAL reference (imported from /root/reference through ``ref_shim.py``) by the golden vectors under
``tests/golden/`` that ``make_golden.py`` generated in the build container.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "knn_oracle.c")
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libknn_oracle.so")

METRIC = {"cosine": 0, "ip": 1, "l2": 2}
SELF = {"keep": 0, "exclude": 1, "minus1": 2}
EPS_MODE = {"clamp": 0, "none": 1, "add": 2, "cast": 3}


def build(force: bool = False) -> str:
    """gcc the C restatement (native flags: it is rebuilt on whichever host runs the tests)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["gcc", "-O2", "-march=native", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True, capture_output=True)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        try:
            lib = C.CDLL(build())
        except OSError:  # built on another host with -march=native
            lib = C.CDLL(build(force=True))
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bf16_round(x: np.ndarray) -> np.ndarray:
    x = _f32(x)
    y = np.empty_like(x)
    _load().oracle_bf16_round(_fp(x), _fp(y), C.c_int64(x.size))
    return y


def row_sqnorm(x: np.ndarray) -> np.ndarray:
    x = _f32(x)
    sq = np.empty((x.shape[0],), dtype=np.float32)
    _load().oracle_row_sqnorm(_fp(x), _fp(sq), C.c_int64(x.shape[0]), C.c_int(x.shape[1]))
    return sq


def normalize(x: np.ndarray, eps: float = 1e-12, eps_mode: str = "clamp", to_bf16: bool = False,
              return_sqnorm: bool = False):
    x = _f32(x)
    y = np.empty_like(x)
    sq = np.empty((x.shape[0],), dtype=np.float32) if return_sqnorm else None
    _load().oracle_normalize(_fp(x), _fp(y), _fp(sq), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), C.c_float(eps),
                             C.c_int(EPS_MODE[eps_mode]), C.c_int(1 if to_bf16 else 0))
    return (y, sq) if return_sqnorm else y


def scores(q: np.ndarray, g: np.ndarray, metric: str = "cosine", self_mode: str = "keep", self_offset: int = 0):
    """Dense [nq, ng] similarity (cosine/ip) or distance (l2) with the engine's fp32 arithmetic."""
    q, g = _f32(q), _f32(g)
    qsq = row_sqnorm(q) if metric == "l2" else None
    gsq = row_sqnorm(g) if metric == "l2" else None
    out = np.empty((q.shape[0], g.shape[0]), dtype=np.float32)
    _load().oracle_scores(_fp(q), _fp(g), _fp(qsq), _fp(gsq), C.c_int64(q.shape[0]), C.c_int64(g.shape[0]),
                          C.c_int(q.shape[1]), C.c_int(METRIC[metric]), C.c_int(SELF[self_mode]),
                          C.c_int64(self_offset), _fp(out))
    return out


def topk(score_matrix: np.ndarray, k: int, largest_first: bool = True, drop_masked: bool = True):
    s = _f32(score_matrix)
    val = np.empty((s.shape[0], k), dtype=np.float32)
    idx = np.empty((s.shape[0], k), dtype=np.int64)
    _load().oracle_topk(_fp(s), C.c_int64(s.shape[0]), C.c_int64(s.shape[1]), C.c_int(k),
                        C.c_int(1 if largest_first else 0), C.c_int(1 if drop_masked else 0), _fp(val), _fp(idx))
    return val, idx


def rank_rows(score_matrix: np.ndarray, largest_first: bool = True) -> np.ndarray:
    return topk(score_matrix, score_matrix.shape[1], largest_first, drop_masked=False)[1]


def search(q: np.ndarray, g: np.ndarray, k: int, metric: str = "cosine", self_mode: str = "keep",
           self_offset: int = 0, index_base: int = 0):
    """(distances [nq,k] fp32, indices [nq,k] int64): the oracle of b200knn.search on prepared rows."""
    q, g = _f32(q), _f32(g)
    qsq = row_sqnorm(q) if metric == "l2" else None
    gsq = row_sqnorm(g) if metric == "l2" else None
    val = np.empty((q.shape[0], k), dtype=np.float32)
    idx = np.empty((q.shape[0], k), dtype=np.int64)
    _load().oracle_search(_fp(q), _fp(g), _fp(qsq), _fp(gsq), C.c_int64(q.shape[0]), C.c_int64(g.shape[0]),
                          C.c_int(q.shape[1]), C.c_int(k), C.c_int(METRIC[metric]), C.c_int(SELF[self_mode]),
                          C.c_int64(self_offset), C.c_int64(index_base), _fp(val), _fp(idx))
    return val, idx


def merge_topk(vals: np.ndarray, idx: np.ndarray, metric: str = "cosine"):
    """k-way merge of [parts, nq, k] candidate lists: best first, ties by ascending global index."""
    parts, nq, k = vals.shape
    out_v = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    for r in range(nq):
        cand = []
        for p in range(parts):
            for j in range(k):
                i = int(idx[p, r, j])
                v = float(vals[p, r, j])
                key = (-v if metric != "l2" else v)
                cand.append((1 if i < 0 else 0, key, i if i >= 0 else 1 << 62, p, j, v, i))
        cand.sort(key=lambda c: c[:5])
        for j in range(k):
            out_v[r, j], out_i[r, j] = cand[j][5], cand[j][6]
    return out_v, out_i


def split_bf16x3(x: np.ndarray, role: str) -> np.ndarray:
    """Restatement of knn_split_bf16x3 (csrc/exact_tc.cu): hi = bf16(x), lo = bf16(x - hi); rows [hi|lo|hi] for
    role "queries", [hi|hi|lo] for "gallery", each part zero-padded to a multiple of 8 columns.  -> fp32 array holding
    bf16-representable values."""
    x = _f32(x)
    n, d = x.shape
    dpad = (d + 7) // 8 * 8
    hi = bf16_round(x)
    lo = bf16_round(x - hi)
    parts = (hi, lo, hi) if role == "queries" else (hi, hi, lo)
    out = np.zeros((n, 3 * dpad), dtype=np.float32)
    for p, part in enumerate(parts):
        out[:, p * dpad:p * dpad + d] = part
    return out
