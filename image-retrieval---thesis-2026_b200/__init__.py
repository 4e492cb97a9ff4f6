"""b200knn: B200-native exact k-NN retrieval + retrieval metrics.

Drop-in for the query x gallery similarity-search hot path of CrispyChillies/Image-Retrieval---Thesis-2026
(L2-normalise -> cosine / L2 -> top-k -> R@K / P@K / mAP / multilabel hit-rate).  The package directory is
``image-retrieval---thesis-2026_b200/``; it is importable as ``b200knn`` through the shim next to it.
"""
from ._lib import KnnError, LIB_PATH, load as load_library
from .search import (FlatIndex, merge_topk, merge_topk_parts, normalize, pack_bits, rank_rows, row_sqnorm,
                     scores_dense, search, search_hamming, split_bf16x3, unpack_bits_pm1)
from . import metrics
from . import fusion
from . import collection
from . import formats
from . import analysis
from . import fullrank
from . import pairwise
from . import sources
from . import nih
from . import chestmir
from .sharded import ShardedFlatIndex

__all__ = [
    "KnnError", "LIB_PATH", "load_library", "FlatIndex", "ShardedFlatIndex", "merge_topk", "normalize",
    "merge_topk_parts", "pack_bits", "rank_rows", "row_sqnorm", "scores_dense", "search", "search_hamming",
    "split_bf16x3", "unpack_bits_pm1", "metrics", "fusion", "collection", "formats", "analysis", "fullrank", "pairwise",
    "sources", "nih", "chestmir",
]
