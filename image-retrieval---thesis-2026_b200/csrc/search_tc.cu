// bf16 distance + fused top-k on the 5th-generation tensor cores (sm_100a).
//
//   S[q, g] = sum_d Q[q,d] * G[g,d]      Q, G bf16 row-major ("K-major" for both UMMA operands)
//
// One CTA owns 128 query rows (UMMA M = 128 = the 128 TMEM lanes) and streams its gallery split in
// tiles of 256 rows (UMMA N = 256).  Warp roles:
//   warp 0      TMA producer  : cp.async.bulk.tensor (128B swizzle) of Q and G k-blocks into a 4-stage ring
//   warp 1      MMA issuer    : one thread issues tcgen05.mma.kind::f16, accumulators in TMEM
//                               (2 stages x 256 fp32 columns = all 512 TMEM columns)
//   warps 2..5  epilogue      : tcgen05.ld 32 columns at a time, thread i owns TMEM lane i = query row i and
//                               runs the threshold filter / candidate lists of select.cuh
// The selection of tile t overlaps the MMAs of tile t+1 through the two TMEM stages, so the Q x N score
// matrix never exists outside TMEM.
#include <stdlib.h>
#include "select.cuh"
#include "ptx.cuh"
#include "kernels.h"

namespace knn {

namespace {

constexpr int TM = 128;        // query rows per CTA
constexpr int TN = 256;        // gallery rows per tile
constexpr int BKE = 64;        // bf16 elements per k-block (= 128 bytes = one swizzle row)
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr uint32_t A_STAGE_BYTES = TM * BKE * 2;   // 16 KB
constexpr uint32_t B_STAGE_BYTES = TN * BKE * 2;   // 32 KB
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;

struct alignas(8) TcBarriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr size_t kSmemBytes = 1024 /*alignment slack*/ + (size_t)kStages * STAGE_BYTES + kDumpBytes +
                              sizeof(float) * kAccStages * TN + sizeof(TcBarriers);

template <int E, bool kL2>
__global__ void __launch_bounds__(kThreads, 1)
search_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                   SearchParams p, unsigned long long* stats, int debug, int prefetch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                       // [kStages][128][64] bf16, swizzled
  uint8_t* smem_b = smem + (size_t)kStages * A_STAGE_BYTES;     // [kStages][256][64] bf16, swizzled
  float* dump = reinterpret_cast<float*>(smem + (size_t)kStages * STAGE_BYTES);  // slow-path staging, 16 KB
  float* gs = dump + kDumpBytes / 4;                                             // [kAccStages][TN]
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(gs + kAccStages * TN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, sp = blockIdx.y;
  const int64_t row0 = (int64_t)qb * TM;
  const int64_t c_begin = (int64_t)sp * p.split_len;
  const int64_t c_end = (c_begin + p.split_len < p.ng) ? c_begin + p.split_len : p.ng;
  const int ntiles = c_end > c_begin ? (int)((c_end - c_begin + TN - 1) / TN) : 0;
  const int nkb = (p.d + BKE - 1) / BKE;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_g);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&bars->tmem_full[s], 1);
      ptx::mbar_init(&bars->tmem_empty[s], kEpiThreads / 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      int pf_t = prefetch / nkb, pf_kb = prefetch % nkb;
      if (prefetch > 0) {
        for (int i = 0; i < prefetch && i < ntiles * nkb; ++i)
          ptx::tma_prefetch_2d(&tmap_g, (i % nkb) * BKE, (int32_t)(c_begin + (int64_t)(i / nkb) * TN));
      }
      for (int t = 0; t < ntiles; ++t) {
        const int32_t col0 = (int32_t)(c_begin + (int64_t)t * TN);
        for (int kb = 0; kb < nkb; ++kb) {
          if (prefetch > 0) {
            if (pf_t < ntiles) ptx::tma_prefetch_2d(&tmap_g, pf_kb * BKE, (int32_t)(c_begin + (int64_t)pf_t * TN));
            if (++pf_kb == nkb) { pf_kb = 0; ++pf_t; }
          }
          const long long c0 = stats ? clock64() : 0;
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          if (stats) w_empty += clock64() - c0;
          ptx::mbar_arrive_expect_tx(&bars->full[stage], STAGE_BYTES);
          ptx::tma_load_2d(smem_a + (size_t)stage * A_STAGE_BYTES, &tmap_q, &bars->full[stage], kb * BKE,
                           (int32_t)row0, ptx::kEvictLast);
          ptx::tma_load_2d(smem_b + (size_t)stage * B_STAGE_BYTES, &tmap_g, &bars->full[stage], kb * BKE, col0,
                           ptx::kEvictNormal);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (stats) atomicAdd(stats + 7, (unsigned long long)w_empty);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(TM, TN);
      int stage = 0;
      uint32_t phase = 0;
      long long w_tmem = 0, w_full = 0;
      const long long m_begin = stats ? clock64() : 0;
      for (int t = 0; t < ntiles; ++t) {
        const int as = t & 1;
        const uint32_t aphase = (uint32_t)(t >> 1) & 1u;
        long long c0 = stats ? clock64() : 0;
        ptx::mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        if (stats) w_tmem += clock64() - c0;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * TN);
        for (int kb = 0; kb < nkb; ++kb) {
          c0 = stats ? clock64() : 0;
          ptx::mbar_wait(&bars->full[stage], phase);
          if (stats) w_full += clock64() - c0;
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = ptx::smem_u32(smem_b + (size_t)stage * B_STAGE_BYTES);
          if (!(debug & 1)) {
#pragma unroll
            for (int k = 0; k < BKE / UMMA_K; ++k) {
              const uint64_t da = ptx::make_sw128_kmajor_desc(a_addr + k * UMMA_K * 2);
              const uint64_t db = ptx::make_sw128_kmajor_desc(b_addr + k * UMMA_K * 2);
              ptx::mma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          ptx::tc_commit(&bars->empty[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(&bars->tmem_full[as]);   // accumulator tile complete
      }
      if (stats) {
        atomicAdd(stats + 0, (unsigned long long)(clock64() - m_begin));
        atomicAdd(stats + 1, (unsigned long long)w_tmem);
        atomicAdd(stats + 2, (unsigned long long)w_full);
        atomicAdd(stats + 8, 1ull);
      }
    }
  } else {
    // ===================================================================== epilogue / selection
    constexpr int L = 32 * E;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int rloc = quarter * 32 + lane;         // query row inside the unit
    const bool row_valid = row0 + rloc < p.nq;
    const int64_t unit = (int64_t)sp * p.qblocks + qb;
    RowState st;
    rowstate_init(st, p.lists + ((unit * TM + rloc) * (int64_t)L));
    uint32_t self_row = 0xFFFFFFFFu;
    float qn = 0.f;
    uint32_t* tau_row = nullptr;
    if (row_valid) {
      const int64_t sr = p.self_offset + row0 + rloc;
      if (p.self_mode != KNN_SELF_KEEP && sr >= 0 && sr < p.ng) self_row = (uint32_t)sr;
      if (kL2) qn = __ldg(p.qsq + row0 + rloc);
      tau_row = p.tau_global + row0 + rloc;
    }
    const int et = threadIdx.x - 64;  // 0..127 among the epilogue threads
    long long e_wait = 0, e_slow = 0;
    unsigned long long n_slow = 0;
    const long long e_begin = stats ? clock64() : 0;

    for (int t = 0; t < ntiles; ++t) {
      const int as = t & 1;
      const uint32_t aphase = (uint32_t)(t >> 1) & 1u;
      const int64_t col0 = c_begin + (int64_t)t * TN;
      float* gst = gs + as * TN;
      if (kL2) {
#pragma unroll
        for (int h = 0; h < TN / kEpiThreads; ++h) {
          int64_t c = col0 + et + h * kEpiThreads;
          if (c >= p.ng) c = p.ng - 1;
          gst[et + h * kEpiThreads] = __ldg(p.gsq + c);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      refresh_tau<kL2>(st, tau_row);
      const long long cw = stats ? clock64() : 0;
      ptx::mbar_wait(&bars->tmem_full[as], aphase);
      if (stats) e_wait += clock64() - cw;
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TN);
#pragma unroll 1
      for (int cb = 0; cb < TN; cb += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + (uint32_t)cb, v);
        ptx::tmem_ld_fence(v);
        const int64_t cg = col0 + cb;
        const int64_t rem = c_end - cg;
        const uint32_t nvalid = rem <= 0 ? 0u : (rem >= 32 ? 32u : (uint32_t)rem);
        const int cnt0 = st.cnt;
        const long long c0 = stats ? clock64() : 0;
        select_chunk_regs<kL2>(st, v, dump + et * 4, qn, gst + cb, (uint32_t)cg, nvalid, self_row, p.self_mode,
                               row_valid && !(debug & 2));
        warp_compact_if_needed<E, 32, kL2>(st, p.k, lane, tau_row);
        if (stats && __any_sync(kFullMask, st.cnt != cnt0)) {
          e_slow += clock64() - c0;
          ++n_slow;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[as]);
    }
    if (stats && lane == 0) {
      atomicAdd(stats + 3, (unsigned long long)(clock64() - e_begin));
      atomicAdd(stats + 4, (unsigned long long)e_wait);
      atomicAdd(stats + 5, (unsigned long long)e_slow);
      atomicAdd(stats + 6, n_slow);
      atomicAdd(stats + 9, 1ull);
      atomicAdd(stats + 10, (unsigned long long)ntiles * (TN / 32));
    }
    // end of unit: the list stays unordered; the unit merge reads `cnt` keys from it
    p.counts[unit * TM + rloc] = row_valid ? st.cnt : 0;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    KNN_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return KNN_E_CUDA;
    }
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return KNN_OK;
}

// [rows, d] bf16 row-major -> 2-D tensor map with a {64, box_rows} box and 128-byte swizzle.
int make_tmap_bf16(CUtensorMap* map, const void* base, int64_t rows, int d, int box_rows) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc != KNN_OK) return rc;
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)BKE, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d=%d)", (int)r, (long long)rows, d);
    return KNN_E_CUDA;
  }
  return KNN_OK;
}

int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

template <int E>
int launch_e(const SearchParams& p, cudaStream_t stream) {
  CUtensorMap tq, tg;
  int rc = make_tmap_bf16(&tq, p.q, p.nq, p.d, TM);
  if (rc != KNN_OK) return rc;
  rc = make_tmap_bf16(&tg, p.g, p.ng, p.d, TN);
  if (rc != KNN_OK) return rc;
  dim3 grid((unsigned)p.qblocks, (unsigned)p.splits);
  if (p.metric == KNN_L2) {
    auto kern = search_bf16_kernel<E, true>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    kern<<<grid, kThreads, kSmemBytes, stream>>>(tq, tg, p, debug_stats_buffer(), env_int("KNN_TC_DEBUG"), env_int("KNN_TC_PREFETCH"));
  } else {
    auto kern = search_bf16_kernel<E, false>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    kern<<<grid, kThreads, kSmemBytes, stream>>>(tq, tg, p, debug_stats_buffer(), env_int("KNN_TC_DEBUG"), env_int("KNN_TC_PREFETCH"));
  }
  KNN_CHECK_CUDA(cudaGetLastError());
  return KNN_OK;
}

}  // namespace

int bf16_tile_cols() { return TN; }

int launch_search_bf16(const SearchParams& p, cudaStream_t stream) {
  if (p.d % 8 != 0) {
    set_error("bf16 search needs d %% 8 == 0 (TMA row pitch must be a multiple of 16 bytes), got d=%d", p.d);
    return KNN_E_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(p.q) & 15) || (reinterpret_cast<uintptr_t>(p.g) & 15)) {
    set_error("bf16 search needs 16-byte aligned q and g");
    return KNN_E_INVALID;
  }
  if (p.qblocks > 1) return launch_search_bf16_pair(p, stream);  // one query block: the one-CTA kernel below
  switch (p.kp) {
    case 32: return launch_e<2>(p, stream);
    case 64: return launch_e<4>(p, stream);
    case 128: return launch_e<8>(p, stream);
    case 256: return launch_e<16>(p, stream);
    default: set_error("unsupported padded k %d", p.kp); return KNN_E_UNSUPPORTED;
  }
}

}  // namespace knn
