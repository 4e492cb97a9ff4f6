// Region (lesion) re-ranking of the first m candidates of every query -- ChestMIR/chestmir_eval.py:507-650
// (rerank_with_specific_lesion / rerank_with_adaptive_lesion) on top-k lists instead of N x N rank matrices.
//
// Every image carries a ragged set of L2-normalised region vectors per lesion type, stored CSR over (image, slot):
// vectors[offsets[img * n_slots + slot] .. offsets[img * n_slots + slot + 1]).  Query q re-ranks with ONE vector (qvec[q])
// of lesion slot qslot[q] (-1: the query has none -> its list is left in global order).  For candidate j:
//   region = max over the candidate's vectors of that slot of <qvec, v>      (-1.0 when it has none)
//   score  = gw * base + (1 - gw) * region                                   (IEEE double, no contraction)
// and the first m candidates are re-ordered by (score, base) descending, equal pairs in their previous order -- the
// `combined_scores.sort(key=lambda x: (x[1], x[2]), reverse=True)` of the reference (a stable sort).  If no candidate
// has a matching region the list is left as it is.  Entries from m on are the reference's `tail`: unchanged.
#include "common.cuh"
#include "kernels.h"

namespace knn {
namespace {

constexpr int kMaxRerank = 1024;

__global__ void __launch_bounds__(128)
lesion_rerank_kernel(const float* __restrict__ cand_val, const int64_t* __restrict__ cand_idx, int K, int m,
                     const float* __restrict__ qvec, const int32_t* __restrict__ qslot,
                     const int64_t* __restrict__ offsets, const float* __restrict__ vectors, int64_t n_img,
                     int n_slots, int dl, double gw, int64_t* __restrict__ out_idx, double* __restrict__ out_score,
                     int32_t* __restrict__ matched) {
  __shared__ double s_score[kMaxRerank];
  __shared__ float s_base[kMaxRerank];
  __shared__ int s_matched, s_valid;
  const int64_t r = blockIdx.x;
  const int64_t* idx = cand_idx + r * K;
  const float* val = cand_val + r * K;
  const int slot = __ldg(qslot + r);
  if (threadIdx.x == 0) { s_matched = 0; s_valid = 0; }
  __syncthreads();
  // candidates are valid up to the first empty entry (lists are filled best-first)
  int nv = 0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) nv += (idx[j] >= 0 && idx[j] < n_img) ? 1 : 0;
  if (nv) atomicAdd(&s_valid, nv);
  __syncthreads();
  const int mm = s_valid;
  bool rerank = slot >= 0 && slot < n_slots && mm > 0;
  if (rerank) {
    const float* qv = qvec + r * (int64_t)dl;
    const double gr = __dsub_rn(1.0, gw);
    int hit = 0;
    for (int j = threadIdx.x; j < mm; j += blockDim.x) {
      const int64_t img = idx[j];
      const int64_t b = __ldg(offsets + img * n_slots + slot), e = __ldg(offsets + img * n_slots + slot + 1);
      float region = -1.0f;
      for (int64_t v = b; v < e; ++v) {
        const float* vec = vectors + v * (int64_t)dl;
        float dot = 0.0f;
        for (int t = 0; t < dl; ++t) dot = fmaf(__ldg(qv + t), __ldg(vec + t), dot);
        region = (v == b) ? dot : fmaxf(region, dot);
      }
      if (region >= 0.0f) ++hit;
      const float base = val[j];
      s_base[j] = base;
      s_score[j] = __dadd_rn(__dmul_rn(gw, (double)base), __dmul_rn(gr, (double)region));
    }
    if (hit) atomicAdd(&s_matched, hit);
  }
  __syncthreads();
  const int nm = s_matched;
  rerank = rerank && nm > 0;
  if (threadIdx.x == 0) matched[r] = (slot >= 0 && slot < n_slots && mm > 0) ? nm : -1;
  if (rerank) {
    // rank counting = stable descending sort on (score, base)
    for (int j = threadIdx.x; j < mm; j += blockDim.x) {
      const double sj = s_score[j];
      const float bj = s_base[j];
      int pos = 0;
      for (int i = 0; i < mm; ++i) {
        const double si = s_score[i];
        const float bi = s_base[i];
        const bool before = si > sj || (si == sj && (bi > bj || (bi == bj && i < j)));
        pos += before ? 1 : 0;
      }
      out_idx[r * K + pos] = idx[j];
      if (out_score) out_score[r * m + pos] = sj;
    }
    for (int j = mm + threadIdx.x; j < K; j += blockDim.x) out_idx[r * K + j] = idx[j];
    if (out_score)
      for (int j = mm + threadIdx.x; j < m; j += blockDim.x) out_score[r * m + j] = -INFINITY;
  } else {
    for (int j = threadIdx.x; j < K; j += blockDim.x) out_idx[r * K + j] = idx[j];
    if (out_score)
      for (int j = threadIdx.x; j < m; j += blockDim.x) out_score[r * m + j] = j < mm ? (double)val[j] : -INFINITY;
  }
}

}  // namespace
}  // namespace knn

using namespace knn;

extern "C" int knn_lesion_rerank(const float* cand_val, const int64_t* cand_idx, int64_t nq, int k, int first_m,
                                 const float* qvec, const int32_t* qslot, const int64_t* offsets, const float* vectors,
                                 int64_t n_img, int n_slots, int dl, double global_weight, int64_t* out_idx,
                                 double* out_score, int32_t* matched, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && first_m >= 0 && first_m <= k, "bad sizes nq=%lld k=%d first_m=%d", (long long)nq, k,
              first_m);
  KNN_REQUIRE(first_m <= kMaxRerank, "first_m=%d exceeds the re-ranking limit %d", first_m, kMaxRerank);
  KNN_REQUIRE(n_img >= 0 && n_slots >= 1 && dl >= 1, "bad region index n_img=%lld n_slots=%d dl=%d", (long long)n_img,
              n_slots, dl);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(cand_val && cand_idx && qvec && qslot && offsets && out_idx && matched, "null pointer");
  KNN_REQUIRE(cand_idx != out_idx, "in-place re-ranking is not supported");
  lesion_rerank_kernel<<<(unsigned)nq, 128, 0, (cudaStream_t)stream>>>(cand_val, cand_idx, k, first_m, qvec, qslot,
                                                                       offsets, vectors, n_img, n_slots, dl,
                                                                       global_weight, out_idx, out_score, matched);
  KNN_LAUNCHED();
  return KNN_OK;
}
