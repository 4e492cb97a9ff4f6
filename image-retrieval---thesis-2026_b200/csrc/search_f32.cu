// Exact-fp32 distance + fused top-k (the parity path).
//
// Arithmetic definition (restated bit-for-bit by oracle/knn_oracle.c):
//   dot(q,g) = fmaf chain over d = 0..D-1 in ascending order starting from +0.0f (one FFMA per term,
//              no split-K, no reassociation), i.e. plain fp32 accumulation like the reference's
//              torch.mm / `@` (test.py:1006, train.py:405; MKL's summation order is unspecified).
//   L2:  d = sqrt(max(rn(rn(|q|^2 + |g|^2) - 2*dot), 0))   (GEMM form of torch.cdist, test_ath.py:87)
// Tile: 128 query rows x 128 gallery rows per CTA, 256 threads, 8x8 register micro-tile.  Scores never leave the SM:
// the tile goes through shared memory to the 128 row-owner threads which run the selection of select.cuh.
// Two operand paths feed the same inner product (same fmaf chain, same bits):
//   row-major rows (KNN_F32):  BK = 16, loaded through registers and stored TRANSPOSED into double-buffered shared
//                              memory ([k][row], what the packed FFMA2 operands need);
//   packed rows (KNN_F32_PACKED, knn_pack_f32): the rows were transposed once into 128-row tiles [tile][k][128], so a
//                              k-block of a tile is one contiguous 4 KB piece -- a four-stage ring filled by 1-D bulk
//                              copies (TMA engine, mbarrier completion): no register staging, no transposing stores, no
//                              block-wide barrier in the k loop, loads of the next tile in flight under the epilogue.
#include "select.cuh"
#include "kernels.h"

namespace knn {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int LDA = BM;       // transposed operand tiles [BK][128]; columns XOR-swizzled by the k quad (see store_tile)
constexpr int LDS = BN + 1;   // padded leading dimension of the score tile
constexpr int kThreads = 256;

struct Frag { float4 v[2]; };

template <bool kVec>
__device__ __forceinline__ void load_tile(const float* __restrict__ X, int64_t r0, int64_t nrows, int d, int k0,
                                          int tid, Frag& f) {
  const int kq = (tid & 3) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int64_t r = r0 + (tid >> 2) + h * 64;
    if (r >= nrows) r = nrows - 1;  // clamp: results of padded rows are discarded later
    const float* p = X + r * (int64_t)d + k0 + kq;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kVec) {
      if (k0 + kq < d) v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      if (k0 + kq + 0 < d) v.x = __ldg(p + 0);
      if (k0 + kq + 1 < d) v.y = __ldg(p + 1);
      if (k0 + kq + 2 < d) v.z = __ldg(p + 2);
      if (k0 + kq + 3 < d) v.w = __ldg(p + 3);
    }
    f.v[h] = v;
  }
}

// Transposing store.  A warp writes 8 consecutive rows x 4 k quads per instruction; with an unpadded k-row of 128 floats
// the four quads would hit the same 8 banks, so the column is XOR-ed with 8 * (k quad): the quads land 8 banks apart
// (conflict-free) and every aligned float4 / 16-float / 32-float group of columns stays one contiguous group for the
// LDS.128 reads of the inner loop.
__device__ __forceinline__ int swz(int k) { return ((k >> 2) & 3) << 3; }

__device__ __forceinline__ void store_tile(float* __restrict__ T, int tid, const Frag& f) {
  const int kq = (tid & 3) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = ((tid >> 2) + h * 64) ^ ((tid & 3) << 3);
    T[(kq + 0) * LDA + r] = f.v[h].x;
    T[(kq + 1) * LDA + r] = f.v[h].y;
    T[(kq + 2) * LDA + r] = f.v[h].z;
    T[(kq + 3) * LDA + r] = f.v[h].w;
  }
}

// kMode: 0 = fused top-k selection, 1 = dense score matrix, 2 = per-row score statistics (sum, sum of squares, min,
// max over the unit's gallery rows, accumulated in double by the row owners; fusion_eval/evaluate.py:152-177)
constexpr int kModeSelect = 0, kModeDense = 1, kModeStats = 2;

// operand paths
constexpr int kLoadScalar = 0, kLoadVec = 1, kLoadPacked = 2;
// packed path: ring of kStages stages, each [A k-block | B k-block] = 2 x [BKP][128] floats
// k-block = 16 steps, two stages: the fixed cost of a loop trip (barrier wait, the register moves ptxas places at the
// loop end, the restart of the LDS -> FFMA2 pipeline) is worth ~3.6 k steps -- 8-step blocks ran at 49 TFLOP/s, 16 at 57,
// 32 at 60 (not worth 64 KB of ring); deeper rings measured no gain (8 x 4 -> 8 x 8, 16 x 2 -> 16 x 4)
constexpr int BKP = 16, kStages = 2;
constexpr int kStageFloats = 2 * BKP * 128;
constexpr uint32_t kStageBytes = kStageFloats * sizeof(float);
constexpr int kRingFloats = kStages * kStageFloats > 4 * BK * LDA ? kStages * kStageFloats : 4 * BK * LDA;
static_assert((kStages & (kStages - 1)) == 0 && BK % BKP == 0, "ring geometry");

// 8 x 8 micro-tile step: acc[i][j] += a[i] * (b[2j], b[2j+1]) on FFMA2
__device__ __forceinline__ void fma_step(unsigned long long (&acc2)[8][4], const float4& a0, const float4& a1,
                                         const ulonglong2& b0, const ulonglong2& b1) {
  const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const unsigned long long b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const unsigned long long ai = ptx::dup_f32x2(a[i]);   // folded into the FFMA2 operand (R.F32 broadcast)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = ptx::ffma2(ai, b[j], acc2[i][j]);
  }
}

struct FragAB { float4 a0, a1; ulonglong2 b0, b1; };

__device__ __forceinline__ void load_frag(FragAB& f, const float* a_s, const float* b_s, int kk, int tx, int ty) {
  f.a0 = *reinterpret_cast<const float4*>(a_s + kk * 128 + ty * 4);
  f.a1 = *reinterpret_cast<const float4*>(a_s + kk * 128 + 64 + ty * 4);
  f.b0 = *reinterpret_cast<const ulonglong2*>(b_s + kk * 128 + tx * 4);
  f.b1 = *reinterpret_cast<const ulonglong2*>(b_s + kk * 128 + 64 + tx * 4);
}

template <int E, bool kL2, int kMode, int kLoad>
__global__ void __launch_bounds__(kThreads, 2) search_f32_kernel(SearchParams p) {
  constexpr bool kDense = kMode == kModeDense;
  constexpr bool kStats = kMode == kModeStats;
  constexpr bool kPacked = kLoad == kLoadPacked;
  constexpr bool kVec = kLoad == kLoadVec;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                     // [2][BK][LDA]            (packed: the ring, [kStages][2][BKP][128])
  float* Bs = As + 2 * BK * LDA;        // [2][BK][LDA]
  uint64_t* bars = reinterpret_cast<uint64_t*>(As + kRingFloats);   // packed: full[kStages], empty[kStages]
  float* Ss = As + kRingFloats + 4 * kStages;   // [BM][LDS]   (not in dense mode)
  float* gs = Ss + BM * LDS;            // [BN]

  const int tid = threadIdx.x, lane = tid & 31;
  // a warp covers 16 query rows x 32 gallery columns per half (lanes: 8 tx x 4 ty): its A fragment is 64 contiguous
  // bytes and its B fragment 128 -- one shared-memory wavefront each (16 tx x 2 ty took 2 and ~4)
  const int tx = (lane & 7) | (((tid >> 5) & 1) << 3), ty = (lane >> 3) | ((tid >> 6) << 2);
  const int qb = blockIdx.x, sp = blockIdx.y;
  const int64_t row0 = (int64_t)qb * BM;
  const int64_t c_begin = (int64_t)sp * p.split_len;
  const int64_t c_end = (c_begin + p.split_len < p.ng) ? c_begin + p.split_len : p.ng;
  const float* __restrict__ Q = reinterpret_cast<const float*>(p.q);
  const float* __restrict__ G = reinterpret_cast<const float*>(p.g);
  const int ntk = (p.d + BK - 1) / BK;

  // selection state of the row owners (threads 0..127 own query row row0 + tid)
  constexpr int L = 32 * E;
  RowState st;
  const bool owner = tid < BM;
  const bool row_valid = owner && (row0 + tid < p.nq);
  uint32_t self_row = 0xFFFFFFFFu;
  float qn = 0.f;
  uint32_t* tau_row = nullptr;
  double st_sum = 0.0, st_sq = 0.0;
  float st_min = INFINITY, st_max = -INFINITY;
  if (!kDense && owner) {
    const int64_t unit = (int64_t)sp * p.qblocks + qb;
    if (!kStats) rowstate_init(st, p.lists + ((unit * BM + tid) * (int64_t)L));
    if (row_valid) {
      const int64_t sr = p.self_offset + row0 + tid;
      if (p.self_mode != KNN_SELF_KEEP && sr >= 0 && sr < p.ng) self_row = (uint32_t)sr;
      if (kL2) qn = __ldg(p.qsq + row0 + tid);
      tau_row = p.tau_global + row0 + tid;
    }
  }

  // packed path: thread 0 is the producer of the ring.  Iterations (tile, k-block) are numbered across the tiles of
  // the unit, so the copies of the next tile's first k-blocks are already in flight during a tile's epilogue.
  const int dpad = (p.d + BK - 1) / BK * BK;       // columns of the packed rows (zero padded like the row-major path)
  const int ntkp = dpad / BKP;
  const uint32_t total_it = kPacked ? (uint32_t)(((c_end - c_begin + BN - 1) / BN) * ntkp) : 0u;
  uint32_t it = 0, pit = 0;                         // consumer / producer iteration
  int pkt = 0;                                      // producer: k-block within its tile
  const float* pa = Q + (int64_t)qb * dpad * 128;   // producer: packed query tile, packed gallery tile
  const float* pb = G + (c_begin / BN) * (int64_t)dpad * 128;
  auto issue = [&]() {
    const uint32_t slot = pit & (kStages - 1), use = pit / kStages;
    if (use > 0) ptx::mbar_wait(&bars[kStages + slot], (use - 1) & 1);   // all 8 warps released the slot
    ptx::mbar_arrive_expect_tx(&bars[slot], kStageBytes);
    const uint32_t dst = ptx::smem_u32(As + slot * kStageFloats), bar = ptx::smem_u32(&bars[slot]);
    ptx::bulk_load_1d(dst, pa + (int64_t)pkt * (BKP * 128), kStageBytes / 2, bar);
    ptx::bulk_load_1d(dst + kStageBytes / 2, pb + (int64_t)pkt * (BKP * 128), kStageBytes / 2, bar);
    ++pit;
    if (++pkt == ntkp) { pkt = 0; pb += (int64_t)dpad * 128; }
  };
  if constexpr (kPacked) {
    if (tid == 0) {
      for (int i = 0; i < kStages; ++i) {
        ptx::mbar_init(&bars[i], 1);
        ptx::mbar_init(&bars[kStages + i], kThreads / 32);
      }
      ptx::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0)
      for (int i = 0; i < kStages - 1 && (uint32_t)i < total_it; ++i) issue();
  }

  for (int64_t col0 = c_begin; col0 < c_end; col0 += BN) {
    // 8 x 8 micro-tile held as 8 x 4 packed pairs (columns 2p, 2p+1): the inner product runs on FFMA2
    // (fma.rn.f32x2, two independent IEEE fused multiply-adds per lane and issue slot -- bit-identical to two
    // scalar fmaf).  The kernel is issue-bound; halving the FMA instruction count lifted the FMA pipe from 62 %.
    unsigned long long acc2[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;  // (+0.0f, +0.0f)

    if constexpr (kPacked) {
      for (int kt = 0; kt < ntkp; ++kt) {
        if (tid == 0 && it + (kStages - 1) < total_it) issue();   // keep kStages - 1 k-blocks in flight
        const uint32_t slot = it & (kStages - 1);
        ptx::mbar_wait(&bars[slot], (it / kStages) & 1);
        const float* a_s = As + slot * kStageFloats;
        const float* b_s = a_s + BKP * 128;
#pragma unroll
        for (int kk = 0; kk < BKP; ++kk) {
          FragAB f;
          load_frag(f, a_s, b_s, kk, tx, ty);
          fma_step(acc2, f.a0, f.a1, f.b0, f.b1);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars[kStages + slot]);   // this warp has read the whole slot
        ++it;
      }
    } else {
      Frag fa, fb;
      load_tile<kVec>(Q, row0, p.nq, p.d, 0, tid, fa);
      load_tile<kVec>(G, col0, p.ng, p.d, 0, tid, fb);
      __syncthreads();  // previous tile's readers of As/Bs/Ss are done
      store_tile(As, tid, fa);
      store_tile(Bs, tid, fb);
      __syncthreads();

      for (int kt = 0; kt < ntk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ntk) {
          load_tile<kVec>(Q, row0, p.nq, p.d, (kt + 1) * BK, tid, fa);
          load_tile<kVec>(G, col0, p.ng, p.d, (kt + 1) * BK, tid, fb);
        }
        const float* a_s = As + buf * BK * LDA;
        const float* b_s = Bs + buf * BK * LDA;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          const float4 a0 = *reinterpret_cast<const float4*>(a_s + kk * LDA + ((ty * 4) ^ swz(kk)));
          const float4 a1 = *reinterpret_cast<const float4*>(a_s + kk * LDA + 64 + ((ty * 4) ^ swz(kk)));
          const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(b_s + kk * LDA + ((tx * 4) ^ swz(kk)));
          const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(b_s + kk * LDA + 64 + ((tx * 4) ^ swz(kk)));
          fma_step(acc2, a0, a1, b0, b1);
        }
        if (kt + 1 < ntk) {
          store_tile(As + (buf ^ 1) * BK * LDA, tid, fa);
          store_tile(Bs + (buf ^ 1) * BK * LDA, tid, fb);
        }
        __syncthreads();
      }
    }

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][2 * j] = __uint_as_float((uint32_t)acc2[i][j]);
        acc[i][2 * j + 1] = __uint_as_float((uint32_t)(acc2[i][j] >> 32));
      }

    if (kDense) {
      // write the tile (score definition identical to the selection path)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + ty * 4 + (i & 3) + (i >> 2) * 64;
        if (r >= p.nq) continue;
        float qni = 0.f;
        if (kL2) qni = __ldg(p.qsq + r);
        const int64_t sr = (p.self_mode != KNN_SELF_KEEP) ? p.self_offset + r : -1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t c = col0 + tx * 4 + (j & 3) + (j >> 2) * 64;
          if (c >= c_end) continue;
          float v = acc[i][j];
          if (kL2) {
            const float x = qni + __ldg(p.gsq + c);
            v = __fsqrt_rn(fmaxf(-fmaf(2.0f, v, -x), 0.0f));
          }
          if (c == sr) {
            if (p.self_mode == KNN_SELF_EXCLUDE) v = kL2 ? INFINITY : -INFINITY;
            else if (p.self_mode == KNN_SELF_MINUS1) v = kL2 ? 1.0f : -1.0f;
          }
          p.dense_out[r * p.ng + c] = v;
        }
      }
    } else {
      // stage the score tile for the row owners
      if (kPacked) __syncthreads();   // the owners are done with the previous tile's scores (no barrier in the k loop)
      if (owner) {
        if (kL2) {
          int64_t c = col0 + tid;
          if (c >= p.ng) c = p.ng - 1;
          gs[tid] = __ldg(p.gsq + c);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty * 4 + (i & 3) + (i >> 2) * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = tx * 4 + (j & 3) + (j >> 2) * 64;
          Ss[r * LDS + c] = acc[i][j];
        }
      }
      __syncthreads();
      if (owner && kStats) {
        if (row_valid) {
          const float* srow = Ss + tid * LDS;
          const int64_t rem = c_end - col0;
          const int ncols = rem >= BN ? BN : (int)rem;
          for (int j = 0; j < ncols; ++j) {
            float v = srow[j];
            if (kL2) v = __fsqrt_rn(fmaxf(-fmaf(2.0f, v, -(qn + gs[j])), 0.0f));
            if ((uint32_t)(col0 + j) == self_row) {
              if (p.self_mode == KNN_SELF_EXCLUDE) continue;
              if (p.self_mode == KNN_SELF_MINUS1) v = -1.0f;
            }
            st_sum += (double)v;
            st_sq = fma((double)v, (double)v, st_sq);
            st_min = fminf(st_min, v);
            st_max = fmaxf(st_max, v);
          }
        }
      } else if (owner) {
        refresh_tau<kL2>(st, tau_row);
        const float* srow = Ss + tid * LDS;
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
          const int64_t cg = col0 + cb;
          const int64_t rem = c_end - cg;
          const uint32_t nvalid = rem <= 0 ? 0u : (rem >= 32 ? 32u : (uint32_t)rem);
          auto fv = [&](int j) -> float {
            const float dot = srow[cb + j];
            if (kL2) return fmaf(2.0f, dot, -(qn + gs[cb + j]));
            return dot;
          };
          select_chunk_mem<kL2>(st, fv, srow + cb, 4, qn, gs + cb, (uint32_t)cg, nvalid, self_row, p.self_mode,
                                row_valid);
          warp_compact_if_needed<E, 32, kL2>(st, p.k, lane, tau_row);
        }
      }
      // next tile starts with a __syncthreads() before Ss / gs are overwritten
    }
  }

  if (kStats) {
    if (owner) {
      double* o = p.stats_out + (((int64_t)sp * p.qblocks + qb) * BM + tid) * 4;
      o[0] = st_sum; o[1] = st_sq; o[2] = (double)st_min; o[3] = (double)st_max;
    }
    return;
  }
  // end of unit: the list stays unordered; the unit merge reads `cnt` keys from it
  if (!kDense && owner) p.counts[((int64_t)sp * p.qblocks + qb) * BM + tid] = row_valid ? st.cnt : 0;
}

size_t f32_smem_bytes(bool dense) {
  return sizeof(float) * (size_t)(kRingFloats + 4 * kStages + (dense ? 0 : BM * LDS + BN));
}

template <int E, bool kL2, int kMode, int kLoad>
int launch_one(const SearchParams& p, cudaStream_t stream) {
  auto kern = search_f32_kernel<E, kL2, kMode, kLoad>;
  const size_t smem = f32_smem_bytes(kMode == kModeDense);
  KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)p.qblocks, (unsigned)p.splits);
  kern<<<grid, kThreads, smem, stream>>>(p);
  KNN_LAUNCHED();
  return KNN_OK;
}

template <int E, int kMode>
int launch_e(const SearchParams& p, int load, cudaStream_t stream) {
  const bool l2 = p.metric == KNN_L2;
  if (l2) {
    if (load == kLoadPacked) return launch_one<E, true, kMode, kLoadPacked>(p, stream);
    return load == kLoadVec ? launch_one<E, true, kMode, kLoadVec>(p, stream)
                            : launch_one<E, true, kMode, kLoadScalar>(p, stream);
  }
  if (load == kLoadPacked) return launch_one<E, false, kMode, kLoadPacked>(p, stream);
  return load == kLoadVec ? launch_one<E, false, kMode, kLoadVec>(p, stream)
                          : launch_one<E, false, kMode, kLoadScalar>(p, stream);
}

// rows [n, d] -> 128-row tiles [ceil(n/128)][dpad][128], dpad = d rounded up to 16, zeros beyond n and d
__global__ void __launch_bounds__(256) pack_f32_kernel(const float* __restrict__ x, int64_t n, int d, int dpad,
                                                       float* __restrict__ out) {
  __shared__ float tile[32][129];
  const int tid = threadIdx.x;
  const int64_t t = blockIdx.x;
  const int k0 = blockIdx.y * 32;
  {
    const int k = k0 + (tid & 31);
#pragma unroll 4
    for (int r = tid >> 5; r < 128; r += 8) {
      const int64_t row = t * 128 + r;
      tile[tid & 31][r] = (row < n && k < d) ? __ldg(x + row * (int64_t)d + k) : 0.0f;
    }
  }
  __syncthreads();
  const int col = tid & 127;
#pragma unroll 4
  for (int kk = tid >> 7; kk < 32; kk += 2) {
    const int k = k0 + kk;
    if (k < dpad) out[(t * dpad + k) * 128 + col] = tile[kk][col];
  }
}

}  // namespace

size_t pack_f32_bytes(int64_t n, int d) {
  const int64_t dpad = (d + BK - 1) / BK * BK;
  return (size_t)((n + 127) / 128) * (size_t)dpad * 128 * sizeof(float);
}

int launch_pack_f32(const float* x, int64_t n, int d, float* out, cudaStream_t stream) {
  if (n <= 0) return KNN_OK;
  const int dpad = (d + BK - 1) / BK * BK;
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)((dpad + 31) / 32));
  pack_f32_kernel<<<grid, 256, 0, stream>>>(x, n, d, dpad, out);
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_search_f32(const SearchParams& p, bool dense, cudaStream_t stream) {
  const int vec = p.f32_packed ? kLoadPacked
                  : ((p.d % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.q) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.g) & 15) == 0)) ? kLoadVec : kLoadScalar;
  if (p.stats_out != nullptr) return launch_e<2, kModeStats>(p, vec, stream);
  if (dense) return launch_e<2, kModeDense>(p, vec, stream);
  switch (p.kp) {
    case 32: return launch_e<2, kModeSelect>(p, vec, stream);
    case 64: return launch_e<4, kModeSelect>(p, vec, stream);
    case 128: return launch_e<8, kModeSelect>(p, vec, stream);
    case 256: return launch_e<16, kModeSelect>(p, vec, stream);
    default: set_error("unsupported padded k %d", p.kp); return KNN_E_UNSUPPORTED;
  }
}

}  // namespace knn
