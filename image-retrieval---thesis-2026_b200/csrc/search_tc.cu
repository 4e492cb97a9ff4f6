// bf16 distance + fused top-k on ONE CTA per unit with the query tile streamed through shared memory.
//
// Serves a single 128-row query block whose rows do not fit tensor memory next to an accumulator (d > 768); the
// common small-batch case (d <= 768) runs search_ts.cu, several query blocks run the CTA-pair kernel search_tc2.cu.
//   S[q, g] = sum_d Q[q,d] * G[g,d]      Q, G bf16 row-major ("K-major" for both UMMA operands)
// UMMA M = 128 (the 128 TMEM lanes = query rows), N = 256 gallery rows per tile, K = 16; a 4-slot ring of
// {Q 16 KB, G 32 KB} k-blocks filled by TMA (128-byte swizzle); fp32 accumulators in TMEM (2 stages x 256 columns).
// Warp roles and the selection are the same as in the other tcgen05 kernels: warp 0 = TMA producer, warp 1 = MMA
// issuer (lean single-lane issue loops), warps 2..9 = selection (two threads per query row taking alternate
// 32-column chunks, separate candidate lists; hits are queued while the accumulator stage is held and appended
// after it has been released).
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
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
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

constexpr size_t kSmemBytes = (size_t)kStages * STAGE_BYTES + sizeof(float) * kAccStages * TN + sizeof(TcBarriers);

// kDiag = false is the production build: the stall counters compile away.
template <int E, bool kL2, bool kDiag>
__global__ void __launch_bounds__(kThreads, 1)
search_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                   SearchParams p, unsigned long long* stats) {
  const bool stats_on = kDiag && stats != nullptr;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;                                                        // [kStages]{B 32 KB, A 16 KB}
  float* gs = reinterpret_cast<float*>(ring + (size_t)kStages * STAGE_BYTES);  // [kAccStages][TN] (L2 metric)
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(gs + kAccStages * TN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, sp = blockIdx.y;
  const int64_t row0 = (int64_t)qb * TM;
  const int64_t c_begin = (int64_t)sp * p.split_len;
  const int64_t c_end = (c_begin + p.split_len < p.ng) ? c_begin + p.split_len : p.ng;
  const int ntiles = c_end > c_begin ? (int)((c_end - c_begin + TN - 1) / TN) : 0;
  const int nkb = (p.d + BKE - 1) / BKE;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("b200knn: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_g);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&bars->tmem_full[s], 1);
      ptx::mbar_init(&bars->tmem_empty[s], kEpiWarps);
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
    // ===================================================================== TMA producer (one thread, lean loop)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      const uint32_t ring_u32 = ptx::smem_u32(ring);
      const uint32_t full0 = ptx::smem_u32(&bars->full[0]);
      int32_t col0 = (int32_t)c_begin;
      for (int t = 0; t < ntiles; ++t, col0 += TN) {
        for (int kb = 0; kb < nkb; ++kb) {
          const long long c0 = stats_on ? clock64() : 0;
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          if (stats_on) w_empty += clock64() - c0;
          const uint32_t dst = ring_u32 + (uint32_t)stage * STAGE_BYTES;
          const uint32_t full_bar = full0 + (uint32_t)stage * 8u;
          ptx::mbar_arrive_expect_tx_u32(full_bar, STAGE_BYTES);
          ptx::tma_load_2d_u32(dst, &tmap_g, full_bar, kb * BKE, col0, ptx::kEvictFirst);
          ptx::tma_load_2d_u32(dst + B_STAGE_BYTES, &tmap_q, full_bar, kb * BKE, (int32_t)row0, ptx::kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (stats_on) atomicAdd(stats + 7, (unsigned long long)w_empty);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp walks, one lane issues)
    constexpr uint32_t idesc = ptx::make_idesc_bf16(TM, TN);
    int stage = 0;
    uint32_t phase = 0;
    long long w_tmem = 0, w_full = 0;
    const long long m_begin = stats_on ? clock64() : 0;
    const uint32_t b_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(ring));
    uint32_t b_lo = b_lo0;
    const bool issuer = ptx::elect_one();
    for (int t = 0; t < ntiles; ++t) {
      const int as = t & 1;
      const uint32_t aphase = (uint32_t)(t >> 1) & 1u;
      long long c0 = stats_on ? clock64() : 0;
      ptx::mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
      if (stats_on) w_tmem += clock64() - c0;
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * TN);
      for (int kb = 0; kb < nkb; ++kb) {
        c0 = stats_on ? clock64() : 0;
        ptx::mbar_wait(&bars->full[stage], phase);
        if (stats_on) w_full += clock64() - c0;
        if (issuer) {
          const uint32_t a_lo = b_lo + (B_STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BKE / UMMA_K; ++k) {
            const uint64_t da = ptx::sw128_desc(a_lo + (uint32_t)(k * UMMA_K * 2 / 16));
            const uint64_t db = ptx::sw128_desc(b_lo + (uint32_t)(k * UMMA_K * 2 / 16));
            ptx::mma_bf16_ss(tmem_d, da, db, idesc, (k != 0 || kb != 0) ? 1u : 0u);
          }
          ptx::tc_commit(&bars->empty[stage]);  // frees the smem slot when these MMAs retire
        }
        b_lo += STAGE_BYTES >> 4;
        if (++stage == kStages) { stage = 0; phase ^= 1; b_lo = b_lo0; }
      }
      if (issuer) ptx::tc_commit(&bars->tmem_full[as]);   // accumulator tile complete
      __syncwarp();
    }
    if (stats_on && issuer) {
      atomicAdd(stats + 0, (unsigned long long)(clock64() - m_begin));
      atomicAdd(stats + 1, (unsigned long long)w_tmem);
      atomicAdd(stats + 2, (unsigned long long)w_full);
      atomicAdd(stats + 8, 1ull);
    }
  } else {
    // ===================================================================== selection (8 warps)
    constexpr int L = 32 * E;
    const int grp = (warp - 2) >> 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int rloc = quarter * 32 + lane;         // query row inside the unit
    const bool row_valid = row0 + rloc < p.nq;
    const int64_t vunit = ((int64_t)sp * 2 + grp) * p.qblocks + qb;
    RowState st;
    rowstate_init(st, p.lists + ((vunit * TM + rloc) * (int64_t)L));
    uint32_t self_row = 0xFFFFFFFFu;
    float qn = 0.f;
    uint32_t* tau_row = nullptr;
    if (row_valid) {
      const int64_t sr = p.self_offset + row0 + rloc;
      if (p.self_mode != KNN_SELF_KEEP && sr >= 0 && sr < p.ng) self_row = (uint32_t)sr;
      if (kL2) qn = __ldg(p.qsq + row0 + rloc);
      tau_row = p.tau_global + row0 + rloc;
    }
    const int et = threadIdx.x - 64;  // 0..255 among the selection threads
    long long e_wait = 0, e_slow = 0;
    unsigned long long n_slow = 0;
    const long long e_begin = stats_on ? clock64() : 0;

    for (int t = 0; t < ntiles; ++t) {
      const int as = t & 1;
      const uint32_t aphase = (uint32_t)(t >> 1) & 1u;
      const int64_t col0 = c_begin + (int64_t)t * TN;
      float* gst = gs + as * TN;
      if (kL2) {
        int64_t c = col0 + et;
        if (c >= p.ng) c = p.ng - 1;
        gst[et] = __ldg(p.gsq + c);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      const uint32_t tau_peek = peek_tau(tau_row);  // L2 round trip hidden behind the barrier wait
      const long long cw = stats_on ? clock64() : 0;
      ptx::mbar_wait(&bars->tmem_full[as], aphase);
      if (stats_on) e_wait += clock64() - cw;
      ptx::tc_fence_after();
      apply_tau<kL2>(st, tau_peek);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TN);
      PendingHits pend;
      pend.n = 0;
      select_tile_tmem<E, kL2>(st, pend, taddr, grp, 2, TN / 32, col0, c_end, gst, qn, self_row, p.self_mode, p.k, lane,
                               tau_row, row_valid, stats_on, e_slow, n_slow);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[as]);
      flush_pending_hits<E, kL2>(st, pend, self_row, p.self_mode, p.k, lane, tau_row, stats_on, e_slow);
    }
    if (stats_on && lane == 0) {
      atomicAdd(stats + 3, (unsigned long long)(clock64() - e_begin));
      atomicAdd(stats + 4, (unsigned long long)e_wait);
      atomicAdd(stats + 5, (unsigned long long)e_slow);
      atomicAdd(stats + 6, n_slow);
      atomicAdd(stats + 9, 1ull);
      atomicAdd(stats + 10, (unsigned long long)ntiles * (TN / 64));
    }
    // end of unit: the list stays unordered; the unit merge reads `cnt` keys from it
    p.counts[vunit * TM + rloc] = row_valid ? st.cnt : 0;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int E>
int launch_e(const SearchParams& p, cudaStream_t stream) {
  CUtensorMap tq, tg;
  int rc = make_tmap_bf16_rows(&tq, p.q, p.nq, p.d, TM);
  if (rc != KNN_OK) return rc;
  rc = make_tmap_bf16_rows(&tg, p.g, p.ng, p.d, TN);
  if (rc != KNN_OK) return rc;
  unsigned long long* stats = debug_stats_buffer();
  dim3 grid((unsigned)p.qblocks, (unsigned)p.splits);
  if (p.metric == KNN_L2) {
    auto kern = search_bf16_kernel<E, true, false>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    kern<<<grid, kThreads, kSmemBytes, stream>>>(tq, tg, p, stats);
  } else if (stats != nullptr && E == 8) {  // diagnostics build exists for the k <= 128, similarity instantiation only
    auto kern = search_bf16_kernel<8, false, true>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    kern<<<grid, kThreads, kSmemBytes, stream>>>(tq, tg, p, stats);
  } else {
    auto kern = search_bf16_kernel<E, false, false>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    kern<<<grid, kThreads, kSmemBytes, stream>>>(tq, tg, p, stats);
  }
  KNN_LAUNCHED();
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
  if (p.groups != 2) {
    set_error("internal: the tcgen05 kernels write 2 candidate lists per row and split");
    return KNN_E_INVALID;
  }
  if (p.qblocks > 1) return launch_search_bf16_pair(p, stream);  // one query block: the one-CTA kernel of this file
  switch (p.kp) {
    case 32: return launch_e<2>(p, stream);
    case 64: return launch_e<4>(p, stream);
    case 128: return launch_e<8>(p, stream);
    case 256: return launch_e<16>(p, stream);
    default: set_error("unsupported padded k %d", p.kp); return KNN_E_UNSUPPORTED;
  }
}

}  // namespace knn
