// Row-wise L2 normalisation fused with the cast to the search dtype, and squared row norms.
//
// Summation order (restated by oracle/knn_oracle.c so the oracle is bit-exact):
//   element e of a row belongs to lane (e / 4) % 32; each lane accumulates acc = fmaf(x, x, acc) over its
//   elements in ascending e starting from +0.0f; the 32 partials are combined by the xor butterfly
//   (offsets 16, 8, 4, 2, 1: acc += shfl_xor(acc, off)).
//   y = x / denom with IEEE division, denom by eps mode (F.normalize clamps: max(||x||, eps)).
#include "common.cuh"

namespace knn {
namespace {

constexpr int kWarpsPerBlock = 8;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float butterfly_sum(float acc) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, off);
  return acc;
}

// 4 consecutive elements starting at e (vectorised when the row pointer allows it)
template <typename T, bool kVec>
__device__ __forceinline__ void load4(const T* __restrict__ row, int e, int d, float (&v)[4]) {
  if (kVec) {
    if (sizeof(T) == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row + e));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(row + e));
      const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
      const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
      v[0] = __low2float(lo); v[1] = __high2float(lo); v[2] = __low2float(hi); v[3] = __high2float(hi);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = (e + c < d) ? to_f<T>(row[e + c]) : 0.0f;
  }
}

template <typename TI, typename TO, bool kVec>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_kernel(const TI* __restrict__ x, TO* __restrict__ y, float* __restrict__ sqnorm, int64_t n, int d,
                 float eps, int eps_mode) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  const TI* row = x + r * (int64_t)d;
  float acc = 0.0f;
  for (int e = lane * 4; e < d; e += 128) {
    float v[4];
    load4<TI, kVec>(row, e, d, v);
#pragma unroll
    for (int c = 0; c < 4; ++c) acc = fmaf(v[c], v[c], acc);
  }
  acc = butterfly_sum(acc);
  if (y == nullptr) {  // norms only
    if (lane == 0) sqnorm[r] = acc;
    return;
  }
  const float nrm = __fsqrt_rn(acc);
  float denom = nrm;
  if (eps_mode == KNN_EPS_CLAMP) denom = fmaxf(nrm, eps);
  else if (eps_mode == KNN_EPS_ADD) denom = nrm + eps;
  else if (eps_mode == KNN_CAST_ONLY) denom = 1.0f;  // x / 1 == x: pure cast to the search dtype
  TO* yrow = y + r * (int64_t)d;
  float acc2 = 0.0f;
  for (int e = lane * 4; e < d; e += 128) {
    float v[4];
    load4<TI, kVec>(row, e, d, v);
    TO o[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      o[c] = from_f<TO>(__fdiv_rn(v[c], denom));
      const float back = to_f<TO>(o[c]);
      if (e + c < d) acc2 = fmaf(back, back, acc2);
    }
    if (kVec) {
      if (sizeof(TO) == 4) {
        *reinterpret_cast<float4*>(yrow + e) = *reinterpret_cast<const float4*>(o);
      } else {
        *reinterpret_cast<uint2*>(yrow + e) = *reinterpret_cast<const uint2*>(o);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (e + c < d) yrow[e + c] = o[c];
    }
  }
  if (sqnorm != nullptr) {
    acc2 = butterfly_sum(acc2);
    if (lane == 0) sqnorm[r] = acc2;
  }
}

template <typename TI, typename TO>
int launch(const void* x, void* y, float* sqnorm, int64_t n, int d, float eps, int eps_mode, cudaStream_t s) {
  const unsigned grid = (unsigned)((n + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) % (4 * sizeof(TI))) == 0) &&
                   (y == nullptr || (reinterpret_cast<uintptr_t>(y) % (4 * sizeof(TO))) == 0);
  if (vec)
    normalize_kernel<TI, TO, true><<<grid, kWarpsPerBlock * 32, 0, s>>>(
        reinterpret_cast<const TI*>(x), reinterpret_cast<TO*>(y), sqnorm, n, d, eps, eps_mode);
  else
    normalize_kernel<TI, TO, false><<<grid, kWarpsPerBlock * 32, 0, s>>>(
        reinterpret_cast<const TI*>(x), reinterpret_cast<TO*>(y), sqnorm, n, d, eps, eps_mode);
  KNN_LAUNCHED();
  return KNN_OK;
}

}  // namespace
}  // namespace knn

using namespace knn;

extern "C" int knn_normalize(const void* x, void* y, float* sqnorm, int64_t n, int d, int in_dtype, int out_dtype,
                             float eps, int eps_mode, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1, "knn_normalize: bad shape n=%lld d=%d", (long long)n, d);
  KNN_REQUIRE(eps_mode >= KNN_EPS_CLAMP && eps_mode <= KNN_CAST_ONLY, "knn_normalize: bad eps_mode %d", eps_mode);
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(x && y, "knn_normalize: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (in_dtype == KNN_F32 && out_dtype == KNN_F32) return launch<float, float>(x, y, sqnorm, n, d, eps, eps_mode, s);
  if (in_dtype == KNN_F32 && out_dtype == KNN_BF16)
    return launch<float, __nv_bfloat16>(x, y, sqnorm, n, d, eps, eps_mode, s);
  if (in_dtype == KNN_BF16 && out_dtype == KNN_BF16)
    return launch<__nv_bfloat16, __nv_bfloat16>(x, y, sqnorm, n, d, eps, eps_mode, s);
  if (in_dtype == KNN_BF16 && out_dtype == KNN_F32)
    return launch<__nv_bfloat16, float>(x, y, sqnorm, n, d, eps, eps_mode, s);
  set_error("knn_normalize: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
  return KNN_E_UNSUPPORTED;
}

extern "C" int knn_row_sqnorm(const void* x, float* sqnorm, int64_t n, int d, int dtype, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1, "knn_row_sqnorm: bad shape n=%lld d=%d", (long long)n, d);
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(x && sqnorm, "knn_row_sqnorm: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == KNN_F32) return launch<float, float>(x, nullptr, sqnorm, n, d, 0.f, KNN_EPS_NONE, s);
  if (dtype == KNN_BF16) return launch<__nv_bfloat16, __nv_bfloat16>(x, nullptr, sqnorm, n, d, 0.f, KNN_EPS_NONE, s);
  set_error("knn_row_sqnorm: unsupported dtype %d", dtype);
  return KNN_E_UNSUPPORTED;
}
