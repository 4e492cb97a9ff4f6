// Exact fp32 search on the TENSOR CORES: filter on tcgen05, decide in fp32.
//
// The exact mode's definition of a score is the fp32 fmaf chain of search_f32.cu (restated by oracle/knn_oracle.c).
// Computing every one of the Q x N scores that way is FFMA-bound (~42 TFLOP/s); but only the k best per query are
// ever returned, so the chain has to run for a handful of candidates per query only -- if a cheap, *bounded-error*
// score can name those candidates.  The bf16 tcgen05 kernels provide exactly that through an error-free split:
//
//   x = hi + lo + e,   hi = bf16_rn(x),  lo = bf16_rn(x - hi)  (x - hi is exact in fp32)
//   bf16 keeps 8 significant bits: |x - hi| <= 2^-8 |x|, and lo sits at least one binade lower: |e| <= 2^-17 |x|
//   q.g ~= qhi.ghi + qlo.ghi + qhi.glo     (dropped: qlo.glo <= 2^-16, qe.g <= 2^-17, (qhi+qlo).ge <= 2^-17 (1 + 2^-7)
//                                           -- together <= 8.04 * 2^-18 * |q||g|; round 1 used u = 2^-9 and got 3.02,
//                                           which rows sitting just below the bf16 rounding midpoints exceed)
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
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
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

// Exact-mode score of one candidate from its fp32 chain `dot` -> sortable key (0 = not a candidate).
template <bool kL2>
__device__ __forceinline__ uint64_t exact_key(float dot, float qn, const float* __restrict__ gsq, int64_t row,
                                              int64_t self_row, int self_mode) {
  float s;
  if (kL2) {
    const float f = fmaf(2.0f, dot, -(qn + __ldg(gsq + row)));
    s = -__fsqrt_rn(fmaxf(-f, 0.0f));
  } else {
    s = dot;
  }
  if (row == self_row) {
    if (self_mode == KNN_SELF_EXCLUDE) return 0ull;
    if (self_mode == KNN_SELF_MINUS1) s = -1.0f;
  }
  return make_key(s, (uint32_t)row);
}

// Run-time check of the hardware model behind eps (the one assumption of the proof: the accumulation error of the tensor
// cores): every candidate carries BOTH its filter value and its exact value, so |filter - exact| is observed on kc rows
// per query for free.  A candidate whose difference exceeds half the bound flags the query as unverified (it is re-run
// through the FFMA engine): the engine never relies on the bound where the data contradicts it.
template <bool kL2>
__device__ __forceinline__ bool filter_value_off(float approx, float dot, float qn, const float* __restrict__ gsq,
                                                 int64_t row, float eps) {
  if (!kL2) return fabsf(approx - dot) > 0.5f * eps;
  const float e2 = fmaxf(-fmaf(2.0f, dot, -(qn + __ldg(gsq + row))), 0.0f);   // exact d^2 (the bound is on d^2)
  const float a2 = approx * approx;
  return fabsf(a2 - e2) > 0.5f * eps + 4.8e-7f * (a2 + e2);                   // + the sqrt rounding of the filter value
}

// Sort the re-scored keys, emit the best k, and decide whether the candidate set provably contains the answer.
template <bool kL2>
__device__ __forceinline__ void emit_and_verify(uint64_t* keys, int npad, int nvalid_cands, int64_t r, int kc, int k,
                                                int64_t index_base, const float* __restrict__ cand_val,
                                                const float* __restrict__ eps, float* __restrict__ out_val,
                                                int64_t* __restrict__ out_idx, int32_t* __restrict__ unverified,
                                                bool model_off) {
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
    if (nvalid_cands < kc) {
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
    unverified[r] = (ok && !model_off) ? 0 : 1;
  }
}

// Pruning inside the candidate set (similarity metrics): the filter's values are sorted best first, and each of its k best
// has an exact score >= f_k - eps (f_k = the k-th best filter value).  A candidate whose filter value lies below
// f_k - 2 eps has an exact score < f_k - eps: at least k rows beat it strictly, it cannot be in the answer and need not
// be re-scored.  With the narrow bound of the three-product filter this never fires; under the wide bound of the
// two-product filter (kc = 2k + slack candidates) it removes ~45 % of the gathered rows.  Rounded down: prunes less.
template <bool kL2>
__device__ __forceinline__ float prune_threshold(const float* __restrict__ vals, int kc, int k, float eps) {
  if (kL2 || k > kc) return -INFINITY;
  return __fmaf_rd(-2.0f, eps, __ldg(vals + (k - 1)));
}

// Generic form: one CTA per query, thread t re-scores candidates t, t + 128, ... reading its rows straight from
// global memory.  Any d, any kc <= 4096.  The exact chain (search_f32.cu's definition):
//   dot = fmaf(q[0], g[0], +0) ... fmaf(q[d-1], g[d-1], dot)
template <bool kL2>
__global__ void __launch_bounds__(kRescoreThreads)
rescore_exact_kernel(const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ qsq,
                     const float* __restrict__ gsq, int64_t ng, int d, int self_mode, int64_t self_offset,
                     int64_t index_base, const float* __restrict__ cand_val, const int64_t* __restrict__ cand_idx,
                     int kc, int k, int npad, const float* __restrict__ eps, float* __restrict__ out_val,
                     int64_t* __restrict__ out_idx, int32_t* __restrict__ unverified) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(rs_smem);             // [npad]
  float* qs = reinterpret_cast<float*>(rs_smem + (size_t)npad * 8);  // [d]
  __shared__ int s_nvalid;
  const int64_t r = blockIdx.x;
  const float* qrow = q + r * (int64_t)d;
  if (threadIdx.x == 0) s_nvalid = 0;
  for (int e = threadIdx.x; e < d; e += blockDim.x) qs[e] = __ldg(qrow + e);
  __syncthreads();
  const int64_t self_row = (self_mode != KNN_SELF_KEEP) ? self_offset + r : -1;  // local gallery row of this query
  const float qn = kL2 ? __ldg(qsq + r) : 0.0f;
  int nvalid = 0;
  int off = 0;
  const float e_r = __ldg(eps + r);
  const float thr = prune_threshold<kL2>(cand_val + r * kc, kc, k, e_r);
  for (int j = threadIdx.x; j < npad; j += blockDim.x) {
    uint64_t key = 0ull;
    if (j < kc) {
      const int64_t row = cand_idx[r * kc + j] - index_base;
      if (cand_idx[r * kc + j] >= 0 && row >= 0 && row < ng) {
        ++nvalid;
        if (j >= k && __ldg(cand_val + r * kc + j) < thr) { keys[j] = 0ull; continue; }   // provably outside the top k
        const float* grow = g + row * (int64_t)d;
        float dot = 0.0f;
        for (int e = 0; e < d; ++e) dot = fmaf(qs[e], __ldg(grow + e), dot);
        key = exact_key<kL2>(dot, qn, gsq, row, self_row, self_mode);
        if (row != self_row) off |= filter_value_off<kL2>(__ldg(cand_val + r * kc + j), dot, qn, gsq, row, e_r) ? 1 : 0;
      }
    }
    keys[j] = key;
  }
  if (nvalid) atomicAdd(&s_nvalid, nvalid);
  const bool model_off = __syncthreads_or(off) != 0;
  emit_and_verify<kL2>(keys, npad, s_nvalid, r, kc, k, index_base, cand_val, eps, out_val, out_idx, unverified,
                       model_off);
}

// Streaming form (d % 4 == 0, kc <= 256): one CTA per query, ONE THREAD PER CANDIDATE.  The candidates' gallery rows
// are scattered over HBM; read 16 bytes at a time by 64 independent threads they arrive as 64-byte fragments of 64
// different DRAM pages (measured 1.2 TB/s).  Here every thread asks the TMA engine for `cf` floats (256 B at kc = 64)
// of ITS row per step -- cp.async.bulk, three steps in flight, all landing on one mbarrier per stage -- and runs the
// chain from shared memory (row pitch cf + 4 floats: conflict-free float4 reads).
constexpr int kRsStages = 3;

template <bool kL2>
__global__ void __launch_bounds__(256)
rescore_exact_stream_kernel(const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ qsq,
                            const float* __restrict__ gsq, int64_t ng, int d, int self_mode, int64_t self_offset,
                            int64_t index_base, const float* __restrict__ cand_val,
                            const int64_t* __restrict__ cand_idx, int kc, int k, int cf, const float* __restrict__ eps,
                            float* __restrict__ out_val, int64_t* __restrict__ out_idx,
                            int32_t* __restrict__ unverified) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const int npad = blockDim.x;                                        // pow2 >= max(kc, 32)
  const int pitch = cf + 4;                                           // floats
  uint64_t* keys = reinterpret_cast<uint64_t*>(rs_smem);              // [npad]
  float* qs = reinterpret_cast<float*>(rs_smem + (size_t)npad * 8);   // [d]
  float* ring = qs + d;                                               // [kRsStages][npad][pitch]
  __shared__ uint64_t bars[kRsStages];
  const int64_t r = blockIdx.x;
  const int c = threadIdx.x;
  if (c == 0) {
    for (int s = 0; s < kRsStages; ++s) ptx::mbar_init(&bars[s], (uint32_t)npad);
    ptx::fence_barrier_init();
  }
  int64_t row = -1;
  bool listed = false;                                                // a row of the gallery (counts for "set = gallery")
  if (c < kc) {
    const int64_t gi = cand_idx[r * kc + c];
    if (gi >= 0 && gi - index_base >= 0 && gi - index_base < ng) {
      listed = true;
      // candidates that provably cannot reach the top k are not re-scored (see prune_threshold)
      if (c < k || !(__ldg(cand_val + r * kc + c) < prune_threshold<kL2>(cand_val + r * kc, kc, k, __ldg(eps + r))))
        row = gi - index_base;
    }
  }
  const bool active = row >= 0;
  const float* grow = g + (active ? row : 0) * (int64_t)d;
  const float* qrow = q + r * (int64_t)d;
  for (int e = c * 4; e < d; e += npad * 4)
    *reinterpret_cast<float4*>(qs + e) = __ldg(reinterpret_cast<const float4*>(qrow + e));
  const int nvalid = __syncthreads_count(listed ? 1 : 0);            // also: barriers initialised, qs complete

  const int nchunks = (d + cf - 1) / cf;
  const uint32_t bar0 = ptx::smem_u32(&bars[0]);
  const uint32_t my_slot0 = ptx::smem_u32(ring + (size_t)c * pitch);
  const uint32_t stage_bytes = (uint32_t)npad * (uint32_t)pitch * 4u;
  auto issue = [&](int i, int s) {
    const uint32_t bar = bar0 + (uint32_t)s * 8u;
    if (active) {
      const int rem = d - i * cf;
      const uint32_t bytes = (uint32_t)(rem < cf ? rem : cf) * 4u;
      ptx::mbar_arrive_expect_tx_u32(bar, bytes);
      ptx::bulk_load_1d(my_slot0 + (uint32_t)s * stage_bytes, grow + (size_t)i * cf, bytes, bar);
    } else {
      ptx::mbar_arrive_u32(bar);
    }
  };
  for (int s = 0; s < kRsStages && s < nchunks; ++s) issue(s, s);

  float dot = 0.0f;
  int s = 0;
  uint32_t phase = 0;
  for (int i = 0; i < nchunks; ++i) {
    ptx::mbar_wait(&bars[s], phase);
    if (active) {
      const float4* b4 = reinterpret_cast<const float4*>(ring + ((size_t)s * npad + c) * pitch);
      const float4* a4 = reinterpret_cast<const float4*>(qs + (size_t)i * cf);
      const int rem = d - i * cf;
      const int n4 = (rem < cf ? rem : cf) >> 2;
#pragma unroll 8
      for (int e = 0; e < n4; ++e) {
        const float4 b = b4[e];
        const float4 a = a4[e];
        dot = fmaf(a.x, b.x, dot);
        dot = fmaf(a.y, b.y, dot);
        dot = fmaf(a.z, b.z, dot);
        dot = fmaf(a.w, b.w, dot);
      }
    }
    __syncthreads();                                   // every reader of stage s is done: it may be refilled
    if (i + kRsStages < nchunks) issue(i + kRsStages, s);
    if (++s == kRsStages) { s = 0; phase ^= 1; }
  }
  const int64_t self_row = (self_mode != KNN_SELF_KEEP) ? self_offset + r : -1;
  const float qn = kL2 ? __ldg(qsq + r) : 0.0f;
  keys[c] = active ? exact_key<kL2>(dot, qn, gsq, row, self_row, self_mode) : 0ull;
  const int off = (active && row != self_row &&
                   filter_value_off<kL2>(__ldg(cand_val + r * kc + c), dot, qn, gsq, row, __ldg(eps + r))) ? 1 : 0;
  const bool model_off = __syncthreads_or(off) != 0;
  emit_and_verify<kL2>(keys, npad, nvalid, r, kc, k, index_base, cand_val, eps, out_val, out_idx, unverified,
                       model_off);
}

// ---------------------------------------------------------------------------------------------- error bound
// max of non-negative floats (squared norms): their bit patterns order like unsigned integers.  NaN-free input assumed.
__global__ void __launch_bounds__(256) max_nonneg_kernel(const float* __restrict__ x, int64_t n, uint32_t* out) {
  uint32_t m = 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(fmaxf(__ldg(x + i), 0.0f)));
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// eps[q] >= |filter value - exact-mode value| for every gallery row; see knn_filter_error_bound in b200knn.h.
__global__ void __launch_bounds__(256) error_bound_kernel(const float* __restrict__ qsq, const float* __restrict__ gmax,
                                                          int64_t nq, int d, int l2, float* __restrict__ eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int dpad = (d + 7) & ~7;
  const double steps = (double)(3 * dpad / 16 + 1);
  // split error + tensor-core accumulation + rounding of the exact fp32 chain + rounding of the fp32 squared norms the
  // bound itself is built from ((d/32 + 6) * 2^-24: the per-lane chains and the butterfly of normalize.cu)
  const double u = 8.04 * 3.814697265625e-06 /*2^-18*/ + steps * 4.76837158203125e-07 /*2^-21*/ * 1.012 +
                   (double)d * 5.9604644775390625e-08 /*2^-24*/ * 1.001 +
                   ((double)d / 32.0 + 6.0) * 5.9604644775390625e-08;
  const double qn2 = (double)__ldg(qsq + i), gn2 = (double)__ldg(gmax);
  double e = u * sqrt(qn2 * gn2) * (1.0 + 1e-6) + 1e-30;
  if (l2) e = 2.0 * e + 4.76837158203125e-07 * 1.01 * (qn2 + gn2);
  eps[i] = __double2float_ru(e);
}

// Two-product filter (q_hi.g_hi + q_lo.g_hi): the dropped product is q.(g - g_hi), bounded by |q| * |g - g_hi|.  The lo
// part of a gallery split row is bf16(g - g_hi): max over the rows of its squared norm (one warp per row, fp32).
__global__ void __launch_bounds__(256) lo_max_sqnorm_kernel(const __nv_bfloat16* __restrict__ rows, int64_t n, int dpad,
                                                            uint32_t* out) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float acc = 0.0f;
  if (r < n) {
    const __nv_bfloat16* lo = rows + r * (int64_t)(3 * dpad) + 2 * dpad;   // gallery rows are [hi | hi | lo]
    for (int c = lane; c < dpad; c += 32) {
      const float v = __bfloat162float(lo[c]);
      acc = fmaf(v, v, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if (lane == 0 && acc > 0.0f) atomicMax(out, __float_as_uint(acc));
}

// eps of the two-product filter: |q| * (max|g_lo| + 2^-16 max|g|)  [dropped product: g - g_hi = g_lo + e, |e| <= 2^-17|g|;
// residual of the query split: |q - q_hi - q_lo| <= 2^-17 |q|] + accumulation of 2 * dpad / 16 + 1 MMA steps + the chain
// and norm roundings of error_bound_kernel.  max|g_lo|^2 is an fp32 sum: (1 + 1e-4) covers its rounding.
__global__ void __launch_bounds__(256) error_bound2_kernel(const float* __restrict__ qsq, const float* __restrict__ gmax,
                                                           const float* __restrict__ glomax, int64_t nq, int d, int l2,
                                                           float* __restrict__ eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int dpad = (d + 7) & ~7;
  const double steps = (double)(2 * dpad / 16 + 1);
  const double u = 4.04 * 3.814697265625e-06 /*2^-18*/ + steps * 4.76837158203125e-07 /*2^-21*/ * 1.012 +
                   (double)d * 5.9604644775390625e-08 /*2^-24*/ * 1.001 +
                   ((double)d / 32.0 + 6.0) * 5.9604644775390625e-08;
  const double qn2 = (double)__ldg(qsq + i), gn2 = (double)__ldg(gmax), lo2 = (double)__ldg(glomax);
  double e = (u * sqrt(qn2 * gn2) + sqrt(qn2 * lo2) * (1.0 + 1e-4)) * (1.0 + 1e-6) + 1e-30;
  if (l2) e = 2.0 * e + 4.76837158203125e-07 * 1.01 * (qn2 + gn2);
  eps[i] = __double2float_ru(e);
}

}  // namespace
}  // namespace knn

using namespace knn;

extern "C" int knn_split_lo_max_sqnorm(const void* gallery_split_rows, int64_t n, int d, float* out, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1 && out, "knn_split_lo_max_sqnorm: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  KNN_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(gallery_split_rows, "knn_split_lo_max_sqnorm: null pointer");
  const int dpad = (d + 7) & ~7;
  lo_max_sqnorm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(gallery_split_rows), n, dpad, reinterpret_cast<uint32_t*>(out));
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_filter_error_bound2(const float* q_sqnorm, int64_t nq, const float* g_sqnorm_max,
                                       const float* g_lo_sqnorm_max, int d, int metric, float* eps, void* stream) {
  KNN_REQUIRE(nq >= 0 && d >= 1, "bad shape nq=%lld d=%d", (long long)nq, d);
  KNN_REQUIRE(metric == KNN_COSINE || metric == KNN_IP || metric == KNN_L2, "bad metric %d", metric);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(q_sqnorm && g_sqnorm_max && g_lo_sqnorm_max && eps, "null pointer");
  error_bound2_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      q_sqnorm, g_sqnorm_max, g_lo_sqnorm_max, nq, d, metric == KNN_L2 ? 1 : 0, eps);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_split_bf16x3(const float* x, int64_t n, int d, int role, void* out, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1, "bad shape n=%lld d=%d", (long long)n, d);
  KNN_REQUIRE(role == 0 || role == 1, "role must be 0 (queries) or 1 (gallery), got %d", role);
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(x && out, "null pointer");
  KNN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  const int dpad = (d + 7) & ~7;
  split_bf16x3_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      x, n, d, dpad, role, reinterpret_cast<__nv_bfloat16*>(out));
  KNN_LAUNCHED();
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
  const int64_t self_local = self_offset - index_base;
  cudaStream_t s = (cudaStream_t)stream;
  const bool l2 = metric == KNN_L2;
  const bool stream_ok = (d & 3) == 0 && kc <= 256 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(q)) & 15) == 0;
  if (stream_ok) {
    int npad = 32;
    while (npad < kc) npad <<= 1;
    // floats per row and step: 256 B per row (measured at 25 000 x 64 candidates x 1024-d: 5.2 TB/s of gathered rows;
    // 512 B: 4.4 TB/s -- fewer CTAs per SM to hide each CTA's prologue / sort; 128 B: 4.7 TB/s), less for wide CTAs
    static const int cf_max = [] {
      const char* e = getenv("KNN_RESCORE_CF");  // experiment knob
      const int v = e ? atoi(e) : 0;
      return (v == 32 || v == 64 || v == 128) ? v : 64;
    }();
    int cf = cf_max;
    while (cf > 32 && (size_t)kRsStages * npad * (cf + 4) * 4 > 104 * 1024) cf >>= 1;
    const size_t smem = (size_t)npad * 8 + (size_t)d * 4 + (size_t)kRsStages * npad * (cf + 4) * 4;
    auto kern = l2 ? rescore_exact_stream_kernel<true> : rescore_exact_stream_kernel<false>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)nq, npad, smem, s>>>(q, g, q_sqnorm, g_sqnorm, ng, d, self_mode, self_local, index_base, cand_val,
                                          cand_idx, kc, k, cf, eps, out_val, out_idx, unverified);
  } else {
    int npad = 2;
    while (npad < kc) npad <<= 1;
    const size_t smem = (size_t)npad * 8 + (size_t)d * 4;
    auto kern = l2 ? rescore_exact_kernel<true> : rescore_exact_kernel<false>;
    if (smem > 48 * 1024)
      KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)nq, kRescoreThreads, smem, s>>>(q, g, q_sqnorm, g_sqnorm, ng, d, self_mode, self_local, index_base,
                                                    cand_val, cand_idx, kc, k, npad, eps, out_val, out_idx, unverified);
  }
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_max_sqnorm(const float* sqnorm, int64_t n, float* out, void* stream) {
  KNN_REQUIRE(n >= 0 && out, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  KNN_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(sqnorm, "null pointer");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  max_nonneg_kernel<<<blocks, 256, 0, s>>>(sqnorm, n, reinterpret_cast<uint32_t*>(out));
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_filter_error_bound(const float* q_sqnorm, int64_t nq, const float* g_sqnorm_max, int d, int metric,
                                      float* eps, void* stream) {
  KNN_REQUIRE(nq >= 0 && d >= 1, "bad shape nq=%lld d=%d", (long long)nq, d);
  KNN_REQUIRE(metric == KNN_COSINE || metric == KNN_IP || metric == KNN_L2, "bad metric %d", metric);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(q_sqnorm && g_sqnorm_max && eps, "null pointer");
  error_bound_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_sqnorm, g_sqnorm_max, nq, d,
                                                                                     metric == KNN_L2 ? 1 : 0, eps);
  KNN_LAUNCHED();
  return KNN_OK;
}
