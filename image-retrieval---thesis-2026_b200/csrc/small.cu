// Small problems (the sizes the reference's own scripts run: a few hundred to a few thousand rows, SURVEY configs 1-2).
// The 128 x 128 tiles of search_f32.cu leave most of the 148 SMs idle there (400 x 400: 16 CTAs), and the candidate-list
// machinery of the fused selection is pure overhead when a whole score row fits shared memory.  Three kernels:
//   dense_small_kernel   32 x 32 score tiles (400 x 400: 169 CTAs), the SAME arithmetic definition as search_f32.cu: one
//                        fmaf chain per score over d ascending from +0.0f, so the two kernels return identical bits;
//   topk_dense_kernel    one CTA per query: the row's <= 4096 scores -> 64-bit keys -> bitonic sort in shared memory ->
//                        the best k in knn_search's output format;
//   first_relevant_rank_kernel  one warp per query: 1-based rank of the best-scoring row with the query's label over a
//                        dense score row -- all that R@K (`retrieval_accuracy`, test.py:38-54) reads off the top-k.
#include "select.cuh"
#include "kernels.h"

namespace knn {
namespace {

constexpr int TS = 32;

template <bool kL2>
__global__ void __launch_bounds__(256) dense_small_kernel(SearchParams p) {
  __shared__ float Qs[TS][TS + 1];
  __shared__ float Gs[TS][TS + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // column of the tile; rows ty, ty + 8, ty + 16, ty + 24
  const int64_t r0 = (int64_t)blockIdx.y * TS, c0 = (int64_t)blockIdx.x * TS;
  const float* __restrict__ Q = reinterpret_cast<const float*>(p.q);
  const float* __restrict__ G = reinterpret_cast<const float*>(p.g);
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int k0 = 0; k0 < p.d; k0 += TS) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int rr = ty + h * 8;
      const int64_t qr = r0 + rr, gr = c0 + rr;
      const int kk = k0 + tx;
      Qs[rr][tx] = (qr < p.nq && kk < p.d) ? __ldg(Q + qr * p.d + kk) : 0.0f;
      Gs[rr][tx] = (gr < p.ng && kk < p.d) ? __ldg(G + gr * p.d + kk) : 0.0f;
    }
    __syncthreads();
    const int kn = p.d - k0 < TS ? p.d - k0 : TS;   // no padded terms: fmaf(0, 0, acc) would turn a -0.0 sum into +0.0
    for (int kk = 0; kk < kn; ++kk) {
      const float g = Gs[tx][kk];
#pragma unroll
      for (int h = 0; h < 4; ++h) acc[h] = fmaf(Qs[ty + h * 8][kk], g, acc[h]);
    }
    __syncthreads();
  }
  const int64_t c = c0 + tx;
  if (c >= p.ng) return;
  const float gn = kL2 ? __ldg(p.gsq + c) : 0.0f;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int64_t r = r0 + ty + h * 8;
    if (r >= p.nq) continue;
    float v = acc[h];
    if (kL2) {
      const float x = __ldg(p.qsq + r) + gn;
      v = __fsqrt_rn(fmaxf(-fmaf(2.0f, v, -x), 0.0f));
    }
    if (p.self_mode != KNN_SELF_KEEP && c == p.self_offset + r) {
      if (p.self_mode == KNN_SELF_EXCLUDE) v = kL2 ? INFINITY : -INFINITY;
      else if (p.self_mode == KNN_SELF_MINUS1) v = kL2 ? 1.0f : -1.0f;
    }
    p.dense_out[r * p.ng + c] = v;
  }
}

__device__ __forceinline__ void small_bitonic_desc(uint64_t* s, int n) {
  for (int size = 2; size <= n; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = s[lo], b = s[hi];
        if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(256) topk_dense_kernel(const float* __restrict__ dense, int64_t nq, int64_t ng, int k,
                                                        int l2, int self_mode, int64_t self_offset, int64_t index_base,
                                                        int npad, float* __restrict__ out_val,
                                                        int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) uint8_t small_smem[];
  uint64_t* s = reinterpret_cast<uint64_t*>(small_smem);
  const int64_t r = blockIdx.x;
  const float* row = dense + r * ng;
  const int64_t self = self_mode == KNN_SELF_EXCLUDE ? self_offset + r : -1;
  for (int j = threadIdx.x; j < npad; j += blockDim.x) {
    uint64_t key = 0ull;
    if (j < ng && j != self) {
      const float v = row[j];
      key = make_key(l2 ? (0.0f - v) : v, (uint32_t)j);
    }
    s[j] = key;
  }
  __syncthreads();
  small_bitonic_desc(s, npad);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = j < npad ? s[j] : 0ull;
    float v;
    int64_t id;
    if (key == 0ull) { v = l2 ? INFINITY : -INFINITY; id = -1; }
    else {
      const float sc = key_score(key);
      v = l2 ? (0.0f - sc) : sc;
      id = (int64_t)key_row(key) + index_base;
    }
    out_val[r * k + j] = v;
    out_idx[r * k + j] = id;
  }
}

__global__ void first_relevant_rank_kernel(const float* __restrict__ scores, int64_t ld, int64_t nq, int64_t ng,
                                           int largest, const int64_t* __restrict__ qlab,
                                           const int64_t* __restrict__ glab, int32_t* __restrict__ first) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  const float* row = scores + q * ld;
  const int64_t ql = qlab[q];
  unsigned long long best = 0ull;   // best key among the rows with the query's label (0 = none)
  for (int64_t g = lane; g < ng; g += 32)
    if (glab[g] == ql) {
      const float v = row[g];
      const unsigned long long key = make_key(largest ? v : (0.0f - v), (uint32_t)g);
      best = key > best ? key : best;
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, best, off);
    best = o > best ? o : best;
  }
  int above = 0;
  if (best != 0ull)
    for (int64_t g = lane; g < ng; g += 32) {
      const float v = row[g];
      above += make_key(largest ? v : (0.0f - v), (uint32_t)g) > best;
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) above += __shfl_xor_sync(0xFFFFFFFFu, above, off);
  if (lane == 0) first[q] = best != 0ull ? above + 1 : 0;
}

}  // namespace

int launch_dense_small(const SearchParams& p, cudaStream_t stream) {
  dim3 grid((unsigned)((p.ng + TS - 1) / TS), (unsigned)((p.nq + TS - 1) / TS));
  if (p.metric == KNN_L2) dense_small_kernel<true><<<grid, 256, 0, stream>>>(p);
  else dense_small_kernel<false><<<grid, 256, 0, stream>>>(p);
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_topk_dense(const float* dense, int64_t nq, int64_t ng, int k, int metric, int self_mode, int64_t self_offset,
                      int64_t index_base, float* out_val, int64_t* out_idx, cudaStream_t stream) {
  int npad = 32;
  while (npad < ng) npad <<= 1;
  const size_t smem = (size_t)npad * sizeof(uint64_t);
  topk_dense_kernel<<<(unsigned)nq, npad >= 512 ? 256 : 64, smem, stream>>>(dense, nq, ng, k, metric == KNN_L2 ? 1 : 0,
                                                                          self_mode, self_offset, index_base, npad, out_val,
                                                                          out_idx);
  KNN_LAUNCHED();
  return KNN_OK;
}

}  // namespace knn

using namespace knn;

extern "C" int knn_first_relevant_rank(const float* scores, int64_t ld_scores, int64_t nq, int64_t ng, int largest_first,
                                       const int64_t* q_labels, const int64_t* g_labels, int32_t* first, void* stream) {
  KNN_REQUIRE(nq >= 0 && ng >= 0 && ld_scores >= ng && ng < 0xFFFFFFFEll, "knn_first_relevant_rank: bad sizes");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(q_labels && first && (ng == 0 || (scores && g_labels)), "knn_first_relevant_rank: null pointer");
  first_relevant_rank_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      scores, ld_scores, nq, ng, largest_first ? 1 : 0, q_labels, g_labels, first);
  KNN_LAUNCHED();
  return KNN_OK;
}
