// Exact fp32 search on the TENSOR CORES: filter on tcgen05, decide in fp32.
//
// The exact mode's definition of a score is the fp32 fmaf chain of search_f32.cu (restated by oracle/knn_oracle.c).
// Computing every one of the Q x N scores that way is FFMA-bound (~42 TFLOP/s); but only the k best per query are
// ever returned, so the chain has to run for a handful of candidates per query only -- if a cheap, *bounded-error*
// score can name those candidates.  The bf16 tcgen05 kernels provide exactly that through an error-free split:
//
//   x = hi + lo + e,   hi = bf16_rn(x),  lo = bf16_rn(x - hi)  (x - hi is exact in fp32),  |e| <= 2^-18 |x|
//   q.g ~= qhi.ghi + qlo.ghi + qhi.glo     (dropped: qlo.glo, qe.g, q.ge  <=  3.02 * 2^-18 * |q||g|)
//
// and the three partial products are ONE inner product of the concatenated rows
//   queries  [ hi | lo | hi ]      gallery  [ hi | hi | lo ]      (3 * dpad wide, dpad = d rounded up to 8)
// so the unmodified bf16 distance + top-k kernels run the filter (knn_split_bf16x3 builds the rows).
//
// knn_rescore_exact then (i) recomputes the kc > k candidates of every query with the exact fp32 chain, (ii) sorts
// them by the usual (score, row) key and emits the best k, (iii) PROVES that no row outside the candidate set can
// belong to the answer: every such row has an approximate score <= m (the kc-th best approximate score), hence an
// exact score <= m + eps, where eps[q] bounds |approximate - exact| for that query (split error + tensor-core
// accumulation + the fp32 chain's own rounding; computed by the host layer from the row norms).  If the k-th best
// exact score t satisfies t > m + eps the emitted top-k is the exact one, bit for bit; otherwise the query is
// flagged and the host layer re-runs it through the FFMA kernel (ties / near-duplicates wider than the slack).
#include "common.cuh"
#include "kernels.h"

namespace knn {
namespace {

// ---------------------------------------------------------------------------------------------- split
// One warp per row.  role 0 (queries): parts = hi, lo, hi;  role 1 (gallery): parts = hi, hi, lo.
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ x, int64_t n, int d, int dpad,
                                                           int role, __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* row = x + r * (int64_t)d;
  __nv_bfloat16* o = out + r * (int64_t)(3 * dpad);
  const bool vec = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (int e = lane * 4; e < dpad; e += 128) {
    float v[4];
    if (vec && e + 4 <= d) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row + e));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = (e + c < d) ? __ldg(row + e + c) : 0.0f;
    }
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      hi[c] = __float2bfloat16_rn(v[c]);
      lo[c] = __float2bfloat16_rn(__fsub_rn(v[c], __bfloat162float(hi[c])));  // v - hi is exact
    }
    const uint2 H = *reinterpret_cast<const uint2*>(hi);
    const uint2 Lo = *reinterpret_cast<const uint2*>(lo);
    *reinterpret_cast<uint2*>(o + e) = H;                                  // dpad % 8 == 0: 8-byte aligned
    *reinterpret_cast<uint2*>(o + dpad + e) = role == 0 ? Lo : H;
    *reinterpret_cast<uint2*>(o + 2 * dpad + e) = role == 0 ? H : Lo;
  }
}

// ---------------------------------------------------------------------------------------------- exact re-scoring
constexpr int kRescoreThreads = 128;

__device__ __forceinline__ void bitonic_desc(uint64_t* s, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int lowmask = stride - 1;
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = ((i & ~lowmask) << 1) | (i & lowmask);
        const bool desc = (pos & size) == 0;
        const uint64_t a = s[pos], b = s[pos + stride];
        if (desc ? (a < b) : (a > b)) { s[pos] = b; s[pos + stride] = a; }
      }
      __syncthreads();
    }
  }
}

// One CTA per query.  Thread t re-scores candidates t, t + 128, ... with the exact chain
//   dot = fmaf(q[0], g[0], +0) ... fmaf(q[d-1], g[d-1], dot)      (search_f32.cu's definition)
template <bool kL2, bool kVec>
__global__ void __launch_bounds__(kRescoreThreads)
rescore_exact_kernel(const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ qsq,
                     const float* __restrict__ gsq, int64_t ng, int d, int self_mode, int64_t self_offset,
                     int64_t index_base, const float* __restrict__ cand_val, const int64_t* __restrict__ cand_idx,
                     int kc, int k, int npad, const float* __restrict__ eps, float* __restrict__ out_val,
                     int64_t* __restrict__ out_idx, int32_t* __restrict__ unverified) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(rs_smem);             // [npad]
  float* qs = reinterpret_cast<float*>(rs_smem + (size_t)npad * 8);  // [d rounded up to 4]
  __shared__ int s_nvalid;
  const int64_t r = blockIdx.x;
  const float* qrow = q + r * (int64_t)d;
  if (threadIdx.x == 0) s_nvalid = 0;
  for (int e = threadIdx.x; e < ((d + 3) & ~3); e += blockDim.x) qs[e] = e < d ? __ldg(qrow + e) : 0.0f;
  __syncthreads();

  const int64_t self_row = (self_mode != KNN_SELF_KEEP) ? self_offset + r : -1;  // local gallery row of this query
  const float qn = kL2 ? __ldg(qsq + r) : 0.0f;
  int nvalid = 0;
  for (int j = threadIdx.x; j < npad; j += blockDim.x) {
    uint64_t key = 0ull;
    if (j < kc) {
      const int64_t gi = cand_idx[r * kc + j];
      const int64_t row = gi - index_base;
      if (gi >= 0 && row >= 0 && row < ng) {
        ++nvalid;
        const float* grow = g + row * (int64_t)d;
        float dot = 0.0f;
        if (kVec) {
          const float4* g4 = reinterpret_cast<const float4*>(grow);
          const float4* q4 = reinterpret_cast<const float4*>(qs);
#pragma unroll 4
          for (int e = 0; e < (d >> 2); ++e) {
            const float4 b = __ldg(g4 + e);
            const float4 a = q4[e];
            dot = fmaf(a.x, b.x, dot);
            dot = fmaf(a.y, b.y, dot);
            dot = fmaf(a.z, b.z, dot);
            dot = fmaf(a.w, b.w, dot);
          }
        } else {
          for (int e = 0; e < d; ++e) dot = fmaf(qs[e], __ldg(grow + e), dot);
        }
        float s;
        if (kL2) {
          const float f = fmaf(2.0f, dot, -(qn + __ldg(gsq + row)));
          s = -__fsqrt_rn(fmaxf(-f, 0.0f));
        } else {
          s = dot;
        }
        bool take = true;
        if (row == self_row) {
          if (self_mode == KNN_SELF_EXCLUDE) take = false;
          else if (self_mode == KNN_SELF_MINUS1) s = -1.0f;
        }
        if (take) key = make_key(s, (uint32_t)row);
      }
    }
    keys[j] = key;
  }
  if (nvalid) atomicAdd(&s_nvalid, nvalid);
  __syncthreads();
  bitonic_desc(keys, npad);

  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = j < npad ? keys[j] : 0ull;
    float v;
    int64_t id;
    if (key == 0ull) {
      v = kL2 ? INFINITY : -INFINITY;
      id = -1;
    } else {
      const float sc = key_score(key);
      v = kL2 ? (0.0f - sc) : sc;
      id = (int64_t)key_row(key) + index_base;
    }
    out_val[r * k + j] = v;
    out_idx[r * k + j] = id;
  }
  if (threadIdx.x == 0) {
    bool ok;
    if (s_nvalid < kc) {
      ok = true;  // the filter returned every admissible row of the gallery: nothing is outside the candidate set
    } else {
      const uint64_t kk = keys[k - 1];
      const double e = (double)__ldg(eps + r);
      const double m = (double)__ldg(cand_val + r * kc + (kc - 1));  // worst approximate value inside the set
      if (kk == 0ull) {
        ok = false;
      } else if (!kL2) {
        ok = (double)key_score(kk) > m + e;
      } else {
        // rows outside the set: approximate distance >= m, i.e. approximate d^2 >= m^2 (1 - 2^-22) (sqrt rounding),
        // exact d^2 >= that - eps, exact distance >= sqrt(...) (1 - 2^-23)
        const double lb2 = m * m * (1.0 - 2.384185791015625e-07) - e;
        const double lb = lb2 > 0.0 ? sqrt(lb2) * (1.0 - 1.1920928955078125e-07) : 0.0;
        ok = (double)(0.0f - key_score(kk)) < lb;
      }
    }
    unverified[r] = ok ? 0 : 1;
  }
}

}  // namespace
}  // namespace knn

using namespace knn;

extern "C" int knn_split_bf16x3(const float* x, int64_t n, int d, int role, void* out, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1, "bad shape n=%lld d=%d", (long long)n, d);
  KNN_REQUIRE(role == 0 || role == 1, "role must be 0 (queries) or 1 (gallery), got %d", role);
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(x && out, "null pointer");
  KNN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  const int dpad = (d + 7) & ~7;
  split_bf16x3_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      x, n, d, dpad, role, reinterpret_cast<__nv_bfloat16*>(out));
  KNN_CHECK_CUDA(cudaGetLastError());
  return KNN_OK;
}

extern "C" int knn_rescore_exact(const float* q, const float* g, const float* q_sqnorm, const float* g_sqnorm,
                                 int64_t nq, int64_t ng, int d, int metric, int self_mode, int64_t self_offset,
                                 int64_t index_base, const float* cand_val, const int64_t* cand_idx, int kc, int k,
                                 const float* eps, float* out_val, int64_t* out_idx, int32_t* unverified,
                                 void* stream) {
  KNN_REQUIRE(nq >= 0 && ng >= 0 && d >= 1, "bad shape nq=%lld ng=%lld d=%d", (long long)nq, (long long)ng, d);
  KNN_REQUIRE(metric == KNN_COSINE || metric == KNN_IP || metric == KNN_L2, "bad metric %d", metric);
  KNN_REQUIRE(self_mode >= KNN_SELF_KEEP && self_mode <= KNN_SELF_MINUS1, "bad self_mode %d", self_mode);
  KNN_REQUIRE(!(metric == KNN_L2 && self_mode == KNN_SELF_MINUS1), "KNN_SELF_MINUS1 is a similarity convention");
  KNN_REQUIRE(k >= 1 && kc >= k && kc <= 4096, "need 1 <= k <= kc <= 4096, got k=%d kc=%d", k, kc);
  KNN_REQUIRE(d <= 16384, "d=%d too large for the shared-memory query row", d);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(q && g && cand_val && cand_idx && eps && out_val && out_idx && unverified, "null pointer");
  KNN_REQUIRE(metric != KNN_L2 || (q_sqnorm && g_sqnorm), "KNN_L2 needs q_sqnorm and g_sqnorm");
  int npad = 2;
  while (npad < kc) npad <<= 1;
  const size_t smem = (size_t)npad * 8 + (size_t)((d + 3) & ~3) * 4;
  const bool vec = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  const int64_t self_local = self_offset - index_base;
  cudaStream_t s = (cudaStream_t)stream;
#define KNN_LAUNCH_RESCORE(L2, VEC)                                                                              \
  do {                                                                                                           \
    auto kern = rescore_exact_kernel<L2, VEC>;                                                                   \
    if (smem > 48 * 1024)                                                                                        \
      KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    kern<<<(unsigned)nq, kRescoreThreads, smem, s>>>(q, g, q_sqnorm, g_sqnorm, ng, d, self_mode, self_local,     \
                                                     index_base, cand_val, cand_idx, kc, k, npad, eps, out_val,  \
                                                     out_idx, unverified);                                       \
  } while (0)
  if (metric == KNN_L2) {
    if (vec) KNN_LAUNCH_RESCORE(true, true); else KNN_LAUNCH_RESCORE(true, false);
  } else {
    if (vec) KNN_LAUNCH_RESCORE(false, true); else KNN_LAUNCH_RESCORE(false, false);
  }
#undef KNN_LAUNCH_RESCORE
  KNN_CHECK_CUDA(cudaGetLastError());
  return KNN_OK;
}
