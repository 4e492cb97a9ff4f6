"""ctypes binding of libb200knn.so (the C ABI declared in include/b200knn.h).

There is deliberately NO fallback: if the shared library is missing or a kernel call fails, the caller
gets an exception.  Nothing in this package computes a distance, a ranking or a metric in torch/numpy.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# KNN_LIB selects another build of the SAME library (the checked build of build.py --check); never a fallback
LIB_PATH = os.environ.get("KNN_LIB") or os.path.join(_HERE, "libb200knn.so")

# constants mirrored from include/b200knn.h
KNN_F32, KNN_BF16, KNN_BF16X3, KNN_F32_PACKED, KNN_BF16X2 = 0, 1, 2, 3, 4
KNN_COSINE, KNN_IP, KNN_L2 = 0, 1, 2
KNN_EPS_CLAMP, KNN_EPS_NONE, KNN_EPS_ADD, KNN_CAST_ONLY = 0, 1, 2, 3
KNN_SELF_KEEP, KNN_SELF_EXCLUDE, KNN_SELF_MINUS1 = 0, 1, 2
MAX_FUSED_K = 256

_p, _i, _i64, _sz, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol include/b200knn.h declares (tests/test_abi.py checks the two agree)
SIGNATURES = {
    "knn_version": (_i, []),
    "knn_last_error": (C.c_char_p, []),
    "knn_normalize": (_i, [_p, _p, _p, _i64, _i, _i, _i, _f, _i, _p]),
    "knn_row_sqnorm": (_i, [_p, _p, _i64, _i, _i, _p]),
    "knn_search": (_i, [_p, _p, _p, _p, _i64, _i64, _i, _i, _i, _i, _i, _i64, _i64, _p, _p, _p, _sz, _p]),
    "knn_search_workspace": (_sz, [_i64, _i64, _i, _i, _i]),
    "knn_split_bf16x3": (_i, [_p, _i64, _i, _i, _p, _p]),
    "knn_max_sqnorm": (_i, [_p, _i64, _p, _p]),
    "knn_filter_error_bound": (_i, [_p, _i64, _p, _i, _i, _p, _p]),
    "knn_filter_error_bound2": (_i, [_p, _i64, _p, _p, _i, _i, _p, _p]),
    "knn_split_lo_max_sqnorm": (_i, [_p, _i64, _i, _p, _p]),
    "knn_rescore_exact": (_i, [_p, _p, _p, _p, _i64, _i64, _i, _i, _i, _i64, _i64, _p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "knn_launch_count": (C.c_longlong, []),
    "knn_profile_enable": (_i, [_i]),
    "knn_profile_count": (_i, []),
    "knn_profile_read": (_i, [_i, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "knn_profile_last": (_i, [C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "knn_debug_stats": (_i, [C.POINTER(C.c_ulonglong), _i]),
    "knn_search_geometry": (_i, [_i64, _i64, _i, _i, _i, C.POINTER(C.c_int64)]),
    "knn_pack_bits": (_i, [_p, _i64, _i, _i, _p, _p]),
    "knn_unpack_bits_pm1": (_i, [_p, _i64, _i, _p, _p]),
    "knn_hamming_from_scores": (_i, [_p, _i64, _i, _p, _p]),
    "knn_search_hamming": (_i, [_p, _p, _i64, _i64, _i, _i, _i, _i64, _i64, _p, _p, _p, _sz, _p]),
    "knn_search_hamming_workspace": (_sz, [_i64, _i64, _i]),
    "knn_score_stats": (_i, [_p, _p, _p, _p, _i64, _i64, _i, _i, _i, _i, _i64, _p, _p, _sz, _p]),
    "knn_score_stats_workspace": (_sz, [_i64, _i64]),
    "knn_rescore_topk": (_i, [_p, _p, _i64, _i, _p, _i64, _i, _p, _f, _f, _i, _i64, _i, _p, _p]),
    "knn_lesion_rerank": (_i, [_p, _p, _i64, _i, _i, _p, _p, _p, _p, _i64, _i, _i, _d, _p, _p, _p, _p]),
    "knn_sort_topk": (_i, [_p, _p, _i64, _i, _i, _p, _p, _p]),
    "knn_pack_f32_bytes": (_sz, [_i64, _i]),
    "knn_pack_f32": (_i, [_p, _i64, _i, _p, _p]),
    "knn_scores_dense": (_i, [_p, _p, _p, _p, _i64, _i64, _i, _i, _i, _i, _i64, _p, _p]),
    "knn_rank_rows": (_i, [_p, _i64, _i64, _i, _p, _p, _sz, _p]),
    "knn_rank_rows_workspace": (_sz, [_i64, _i64]),
    "knn_merge_topk": (_i, [_p, _p, _i, _i64, _i, _i, _p, _p, _p]),
    "knn_merge_topk_parts": (_i, [_p, _p, _i, _i64, _i, _i, _p, _p, _p]),
    "knn_merge_topk_parts_sync": (_i, [_p, _p, _i, _i64, _i, _i, _p, _i, _i, _p, _p, _p]),
    "knn_merge_topk_parts_wait": (_i, [_p, _p, _i, _i64, _i, _i, _p, _i, _i, _p, _p, _p]),
    "knn_peer_publish": (_i, [_p, _i, _i, _i, _p]),
    "knn_relevance_single": (_i, [_p, _i64, _i, _p, _p, _i64, _p, _p, _p]),
    "knn_relevance_multilabel": (_i, [_p, _i64, _i, _p, _p, _i64, _d, _i, _p, _p, _p]),
    "knn_ranked_stats": (_i, [_p, _i64, _i, _i, _p, _p, _p, _p, _p]),
    "knn_ranked_stats_multi": (_i, [_p, _i64, _i, _p, _i, _p, _p, _p, _p, _p]),
    "knn_recall_counts": (_i, [_p, _i64, _i, _p, _p, _i64, _p, _i, _p, _p]),
    "knn_majority_vote": (_i, [_p, _i64, _i, _i, _i, _p, _p]),
    "knn_majority_vote_multi": (_i, [_p, _i64, _i, _p, _i, _i, _p, _p]),
    "knn_map_full": (_i, [_p, _i64, _i64, _p, _p, _p, _i, _p, _p, _p, _p]),
    "knn_ap_sklearn": (_i, [_p, _p, _i64, _i, _p, _p, _sz, _p]),
    "knn_ap_sklearn_workspace": (_sz, [_i64, _i]),
    "knn_rank_of_positives": (_i, [_p, _i64, _i64, _i64, _i, _i, _p, _p, _d, _i64, _i, _p, _p, _p, _i64, _p, _p, _p,
                                   _p, _p, _p, _sz, _p]),
    "knn_rank_of_positives_workspace": (_sz, [_i64, _i64]),
    "knn_ap_from_ranks": (_i, [_p, _i64, _p, _p, _i64, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "knn_ap_sklearn_from_ranks": (_i, [_p, _p, _i64, _p, _p, _i64, _p, _p, _sz, _p]),
    "knn_ap_sklearn_from_ranks_workspace": (_sz, [_i64, _i64]),
    "knn_first_relevant_rank": (_i, [_p, _i64, _i64, _i64, _i, _p, _p, _p, _p]),
    "knn_triplet_mine": (_i, [_p, _p, _i64, _f, _p, _p, _p, _p, _p]),
    "knn_jaccard_matrix": (_i, [_p, _p, _i64, _i64, _f, _p, _p]),
    "knn_class_means": (_i, [_p, _p, _i64, _i, _p, _i, _p, _p, _p]),
    "knn_centroid_min_dist": (_i, [_p, _p, _i64, _i, _i, _p, _p]),
}

_lock = threading.Lock()
_lib = None


class KnnError(RuntimeError):
    """A b200knn C-ABI call returned a negative status."""


def load() -> C.CDLL:
    """Load the shared library once; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise KnnError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). b200knn has no CPU / torch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().knn_last_error()
        raise KnnError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
