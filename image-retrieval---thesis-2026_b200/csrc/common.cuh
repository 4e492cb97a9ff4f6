// Shared device/host helpers for the b200knn kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b200knn.h"

namespace knn {

// ----------------------------------------------------------------------------------------------
// error plumbing (thread-local message, see api.cu)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define KNN_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      knn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return KNN_E_CUDA;                                                                       \
    }                                                                                          \
  } while (0)

// After every kernel launch of the library: surface the launch error and count the launch (knn_launch_count).
void count_launch();
#define KNN_LAUNCHED()                   \
  do {                                   \
    KNN_CHECK_CUDA(cudaGetLastError());  \
    knn::count_launch();                 \
  } while (0)

#define KNN_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      knn::set_error(__VA_ARGS__);        \
      return KNN_E_INVALID;               \
    }                                     \
  } while (0)

// ----------------------------------------------------------------------------------------------
// 64-bit sortable candidate keys.
//   key = ord(score) << 32 | (0xFFFFFFFF - local_gallery_row)
// ord() is the usual order-preserving map of an fp32 onto uint32, so a LARGER key is a BETTER
// candidate: higher score first, then (equal score) lower gallery row first.  Keys are unique per
// gallery row, which makes every selection below deterministic.  "score" is always larger=better:
// similarity for cosine/ip, minus the distance for L2.  key 0 is the empty slot.
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  score = score + 0.0f;  // canonicalise -0.0 -> +0.0 so that equal floats give equal ord()
  return ((uint64_t)f2ord(score) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) { return ord2f((uint32_t)(key >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)key; }

constexpr int kRowsPerUnit = 128;  // query rows owned by one CTA (= UMMA M, = TMEM lanes)

// Candidate-list geometry for a given k: KP = pow2 >= k (>= 32), list capacity L = 2*KP.
inline int kpad_for(int k) {
  int kp = 32;
  while (kp < k) kp <<= 1;
  return kp;
}
constexpr int kMaxFusedK = 256;  // larger k goes through the dense + rank path

struct SearchGeom {
  int qblocks;        // ceil(nq / 128)
  int splits;         // gallery splits S
  int groups;         // candidate lists per (split, row)
  int64_t split_len;  // gallery rows per split (multiple of the column tile)
  int kp;             // padded k
  int L;              // list capacity per row
  int seed_splits;    // units per query block of the threshold-seeding pre-pass (0 = none)
  int64_t seed_len;   // gallery rows per seeding unit
  int seed_stride;    // > 0: the pre-pass collects chunk maxima (SearchParams::seed_stride) instead of selecting
};

}  // namespace knn
