// Hamming distance over packed binary codes + fused top-k.
//
// Replaces `(q[:, None, :] != g[None, :, :]).sum(dim=2).float()` + `argsort(dim=1)` of test_ath.py:80-100 /
// train_ath.py:162-175, which materialises a [Q, N, bits] int16 tensor.  Codes are packed 64 bits per word
// (knn_pack_bits); distance = sum_w popc(q_w ^ g_w), an exact small integer.  Integer work, HBM / issue bound: no
// tensor cores.  One CTA owns 128 query rows (one thread each, its code words in registers) and walks its gallery
// split 32 rows at a time through shared memory (coalesced 128-bit loads); the 32 distances of a chunk go through the
// same threshold filter / candidate lists as the floating-point kernels with score = -distance, so ties are broken by
// ascending gallery row and the unit merge is shared (csrc/merge.cu).
#include "select.cuh"
#include "kernels.h"

namespace knn {

namespace {

constexpr int kRows = 128;      // query rows per CTA
constexpr int kChunk = 32;      // gallery rows per step
constexpr int kMaxWords = 8;    // codes up to 512 bits
constexpr int LDS = kChunk + 1;

template <int E, int W>
__global__ void __launch_bounds__(kRows, 4) search_hamming_kernel(SearchParams p) {
  __shared__ uint64_t gtile[kChunk * W];
  __shared__ float scores[kRows * LDS];
  const int tid = threadIdx.x, lane = tid & 31;
  const int qb = blockIdx.x, sp = blockIdx.y;
  const int64_t row0 = (int64_t)qb * kRows;
  const int64_t c_begin = (int64_t)sp * p.split_len;
  const int64_t c_end = (c_begin + p.split_len < p.ng) ? c_begin + p.split_len : p.ng;
  const uint64_t* __restrict__ Q = reinterpret_cast<const uint64_t*>(p.q);
  const uint64_t* __restrict__ G = reinterpret_cast<const uint64_t*>(p.g);
  constexpr int L = 32 * E;

  const bool row_valid = row0 + tid < p.nq;
  uint64_t qw[W];
#pragma unroll
  for (int w = 0; w < W; ++w) qw[w] = row_valid ? __ldg(Q + (row0 + tid) * W + w) : 0ull;
  RowState st;
  const int64_t unit = (int64_t)sp * p.qblocks + qb;
  rowstate_init(st, p.lists + ((unit * kRows + tid) * (int64_t)L));
  uint32_t self_row = 0xFFFFFFFFu;
  uint32_t* tau_row = nullptr;
  if (row_valid) {
    const int64_t sr = p.self_offset + row0 + tid;
    if (p.self_mode != KNN_SELF_KEEP && sr >= 0 && sr < p.ng) self_row = (uint32_t)sr;
    tau_row = p.tau_global + row0 + tid;
  }
  float* srow = scores + tid * LDS;

  for (int64_t col0 = c_begin; col0 < c_end; col0 += kChunk) {
    __syncthreads();  // previous chunk's readers of gtile are done
    for (int i = tid; i < kChunk * W; i += kRows) {
      const int64_t r = col0 + i / W;
      gtile[i] = r < c_end ? __ldg(G + r * W + (i % W)) : 0ull;
    }
    __syncthreads();
    // Fast path in integers: 32 distances in registers, their minimum against the row's threshold distance.  Only a
    // chunk that can contribute (min d <= d_k; always, until k candidates exist) takes the float / shared-memory
    // route of the common selection code.
    int dist[kChunk];
    int dmin = 1 << 30;
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      int d = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) d += __popcll(qw[w] ^ gtile[j * W + w]);
      dist[j] = d;
      dmin = min(dmin, d);
    }
    if ((col0 - c_begin) % (16 * kChunk) == 0) refresh_tau<false>(st, tau_row);
    // st.tau = -(k-th best distance) (or -inf): a row qualifies iff -d >= tau
    const bool hit = row_valid && -(float)dmin >= st.ftau;
    if (__any_sync(kFullMask, hit)) {
      if (hit) {
#pragma unroll
        for (int j = 0; j < kChunk; ++j) srow[j] = -(float)dist[j];
        const int64_t rem = c_end - col0;
        const uint32_t nvalid = rem >= kChunk ? (uint32_t)kChunk : (uint32_t)rem;
        auto fv = [&](int j) -> float { return srow[j]; };
        select_chunk_mem<false>(st, fv, srow, 4, 0.f, nullptr, (uint32_t)col0, nvalid, self_row, p.self_mode, row_valid);
      }
      warp_compact_if_needed<E, 32, false>(st, p.k, lane, tau_row);
    }
  }
  p.counts[unit * kRows + tid] = row_valid ? st.cnt : 0;
}

template <int E>
int launch_w(const SearchParams& p, int words, cudaStream_t stream) {
  dim3 grid((unsigned)p.qblocks, (unsigned)p.splits);
  switch (words) {
    case 1: search_hamming_kernel<E, 1><<<grid, kRows, 0, stream>>>(p); break;
    case 2: search_hamming_kernel<E, 2><<<grid, kRows, 0, stream>>>(p); break;
    case 3: search_hamming_kernel<E, 3><<<grid, kRows, 0, stream>>>(p); break;
    case 4: search_hamming_kernel<E, 4><<<grid, kRows, 0, stream>>>(p); break;
    case 8: search_hamming_kernel<E, 8><<<grid, kRows, 0, stream>>>(p); break;
    default: set_error("hamming search: %d words per code not instantiated (1, 2, 3, 4, 8)", words); return KNN_E_UNSUPPORTED;
  }
  KNN_LAUNCHED();
  return KNN_OK;
}

// one thread per (row, word): bit b of word w = (x[row][64*w + b] != 0)
template <typename T>
__global__ void pack_bits_kernel(const T* __restrict__ x, int64_t n, int bits, int words, uint64_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * words) return;
  const int64_t r = t / words;
  const int w = (int)(t % words);
  uint64_t v = 0ull;
  const int b0 = w * 64;
  const int nb = bits - b0 < 64 ? bits - b0 : 64;
  for (int b = 0; b < nb; ++b) v |= (uint64_t)(x[r * bits + b0 + b] != T(0)) << b;
  out[t] = v;
}

// packed words -> +-1 rows in bf16 (bit set -> +1, clear -> -1, columns bits .. dpad-1 -> 0): the operand form of the
// tensor-core Hamming search, <q, g> = bits - 2 * hamming(q, g) exactly.  One thread per 8 output columns.
__global__ void __launch_bounds__(256) unpack_pm1_kernel(const uint64_t* __restrict__ words, int64_t n, int bits,
                                                         int nwords, int dpad, __nv_bfloat16* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = dpad / 8;
  if (t >= n * groups) return;
  const int64_t r = t / groups;
  const int c0 = (int)(t % groups) * 8;
  const uint64_t w = __ldg(words + r * nwords + (c0 >> 6));
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float x = c < bits ? (((w >> (c & 63)) & 1ull) ? 1.0f : -1.0f) : 0.0f;
    v[j] = __float2bfloat16_rn(x);
  }
  *reinterpret_cast<uint4*>(out + r * (int64_t)dpad + c0) = *reinterpret_cast<const uint4*>(v);
}

__global__ void __launch_bounds__(256) hamming_from_scores_kernel(const float* __restrict__ score, int64_t n, float bits,
                                                                 float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float s = score[t];
  out[t] = s == -INFINITY ? INFINITY : (bits - s) * 0.5f;   // integers: exact
}

}  // namespace

int launch_hamming_from_scores(const float* score, int64_t n, int bits, float* out, cudaStream_t stream) {
  if (n == 0) return KNN_OK;
  hamming_from_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(score, n, (float)bits, out);
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_unpack_pm1(const uint64_t* words, int64_t n, int bits, int nwords, void* out, cudaStream_t stream) {
  const int dpad = (bits + 7) & ~7;
  const int64_t total = n * (dpad / 8);
  if (total == 0) return KNN_OK;
  unpack_pm1_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(words, n, bits, nwords, dpad,
                                                                        reinterpret_cast<__nv_bfloat16*>(out));
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_search_hamming(const SearchParams& p, int words, cudaStream_t stream) {
  switch (p.kp) {
    case 32: return launch_w<2>(p, words, stream);
    case 64: return launch_w<4>(p, words, stream);
    case 128: return launch_w<8>(p, words, stream);
    case 256: return launch_w<16>(p, words, stream);
    default: set_error("unsupported padded k %d", p.kp); return KNN_E_UNSUPPORTED;
  }
}

int launch_pack_bits(const void* x, int dtype, int64_t n, int bits, int words, uint64_t* out, cudaStream_t stream) {
  const int64_t total = n * words;
  if (total == 0) return KNN_OK;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == KNN_F32) pack_bits_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(x), n, bits, words, out);
  else {
    set_error("knn_pack_bits: fp32 codes only");
    return KNN_E_UNSUPPORTED;
  }
  KNN_LAUNCHED();
  return KNN_OK;
}

}  // namespace knn
