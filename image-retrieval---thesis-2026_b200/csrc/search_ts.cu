// bf16 distance + fused top-k with the QUERY TILE RESIDENT IN TENSOR MEMORY (tcgen05.mma "TS" form).
//
//   S[q, g] = sum_d Q[q,d] * G[g,d]
//
// The 128 query rows a CTA owns are written ONCE into TMEM (lane = query row, 32-bit column = two consecutive bf16
// of the row) and every tcgen05.mma reads its A operand from there.  Shared memory then holds nothing but the
// gallery stream: a ring of up to 28 k-block slots (~200 KB in flight per SM) instead of the 4-5 slots left over
// when the query tile (or its k-blocks) also lives in shared memory.  That is what both regimes need:
//   * small query batches (HBM-bound): every byte that crosses L2 -> SM is a gallery byte, read once;
//   * large query batches (tensor-bound): the ring rides out the refill latency, the tensor pipe never starves.
// TMEM budget (512 columns): A takes D/2 columns, the rest is two accumulator stages of N = 128 / 64 / 32 columns
// (D <= 512 / 768 / 896).  Larger D falls back to the shared-memory-A kernels (search_tc.cu, search_tc2.cu).
//
// CG = 1: one CTA per unit (single 128-row query block; each SM streams its own gallery split).
// CG = 2: a CTA pair owns 256 query rows; each CTA TMA-loads HALF of every gallery tile, one thread of the leader
//         issues tcgen05.mma.cta_group::2 (M = 256), barriers as in search_tc2.cu.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = selection.  TMEM lane i = query row i is
// owned by TWO selection threads (warp groups 0 and 1), which take alternate 32-column chunks of every tile and
// keep separate candidate lists ("virtual units"); the unit merge sees splits * 2 lists per row.
// The selection slow path re-reads the few surviving columns from TMEM (no shared-memory staging).
#include <stdlib.h>
#include <string.h>
#include "select.cuh"
#include "ptx.cuh"
#include "kernels.h"

namespace knn {

namespace {

constexpr int TM = 128;        // query rows per CTA
constexpr int BKE = 64;        // bf16 elements per k-block (128 B = one swizzle row)
constexpr int UMMA_K = 16;
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kMaxStages = 32;
constexpr int kMaxTileN = 128;
constexpr size_t kSmemBudget = 232448;  // 227 KB opt-in limit per CTA

struct alignas(8) TsBarriers {
  uint64_t full[kMaxStages];   // leader: CG producer arrivals + tx bytes of all CTAs
  uint64_t empty[kMaxStages];  // every CTA: 1 arrival (tcgen05.commit, multicast for CG = 2)
  uint64_t a_full;             // leader: kEpiWarps * CG arrivals (query tile stored to TMEM)
  uint64_t tmem_full[2];       // every CTA: 1 arrival (commit)
  uint64_t tmem_empty[2];      // leader: kEpiWarps * CG arrivals
  uint32_t tmem_base;
  uint32_t pad;
};

struct TsCfg {
  int stages;        // ring depth
  int nkb;           // k-blocks
  int tn;            // gallery rows per tile (whole CTA group)
  int acc_stages;    // accumulator stages in TMEM: 2 when the query tile leaves room, else 1
  uint32_t stage_bytes;  // per CTA: (tn / CG) rows x 128 B
  int debug;         // KNN_TS_DEBUG (timing experiments, results are garbage): bit 0 = no MMA, bit 1 = selection
                     // fast path only, bit 2 = no TMA (MMAs run on whatever the ring holds), bit 3 = no selection
  unsigned long long* stats;
};

// kDiag = false is the production build: the stall counters and timing-experiment switches compile away.
template <int CG, int E, bool kL2, bool kDiag>
__global__ void __launch_bounds__(kThreads, 1)
search_bf16_ts_kernel(const __grid_constant__ CUtensorMap tmap_g, SearchParams p, TsCfg cfg) {
  const bool stats_on = kDiag && cfg.stats != nullptr;
  const int debug = kDiag ? cfg.debug : 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;                                                                      // [stages][stage_bytes]
  float* gs = reinterpret_cast<float*>(ring + (size_t)cfg.stages * cfg.stage_bytes);         // [2][kMaxTileN]
  TsBarriers* bars = reinterpret_cast<TsBarriers*>(gs + 2 * kMaxTileN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int qb = blockIdx.x;  // 128-row query block of THIS CTA
  const int sp = blockIdx.y;
  const int64_t row0 = (int64_t)qb * TM;
  const int TN = cfg.tn;
  const int rows_cta = TN / CG;  // gallery rows this CTA loads per tile
  const int64_t c_begin = (int64_t)sp * p.split_len;
  const int64_t c_end = (c_begin + p.split_len < p.ng) ? c_begin + p.split_len : p.ng;
  const int ntiles = c_end > c_begin ? (int)((c_end - c_begin + TN - 1) / TN) : 0;
  const int nkb = cfg.nkb;
  const int stages = cfg.stages;
  const int ACC = cfg.acc_stages;
  const uint32_t a_col0 = (uint32_t)(ACC * TN);  // TMEM column of the query tile (after the accumulator stages)

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("b200knn: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmap_g);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&bars->full[s], CG);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, kEpiWarps * CG);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bars->tmem_full[s], 1);
      ptx::mbar_init(&bars->tmem_empty[s], kEpiWarps * CG);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      ptx::tmem_alloc_2sm(&bars->tmem_base, kTmemCols);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(&bars->tmem_base, kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer (every CTA)
    // One thread; the loop is kept to a handful of instructions per slot: a single thread's dependent
    // instruction stream costs ~5 cycles per instruction, and that -- not the TMA unit -- bounds the slot rate.
    if (lane == 0 && !(debug & 4)) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      const uint32_t ring_u32 = ptx::smem_u32(ring);
      const uint32_t full0 = CG == 2 ? ptx::mapa(ptx::smem_u32(&bars->full[0]), 0) : ptx::smem_u32(&bars->full[0]);
      const uint32_t tx_bytes = (uint32_t)CG * cfg.stage_bytes;
      int32_t col0 = (int32_t)(c_begin + (int64_t)rank * rows_cta);
      for (int t = 0; t < ntiles; ++t, col0 += TN) {
        for (int kb = 0; kb < nkb; ++kb) {
          const long long c0 = stats_on ? clock64() : 0;
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          if (stats_on) w_empty += clock64() - c0;
          const uint32_t dst = ring_u32 + (uint32_t)stage * cfg.stage_bytes;
          const uint32_t full_bar = full0 + (uint32_t)stage * 8u;
          if (CG == 2) {
            ptx::tma_load_2d_2sm_u32(dst, &tmap_g, full_bar, kb * BKE, col0, ptx::kEvictNormal);
            if (leader) ptx::mbar_arrive_expect_tx_u32(full_bar, tx_bytes);
            else ptx::mbar_arrive_cluster(full_bar);
          } else {
            ptx::mbar_arrive_expect_tx_u32(full_bar, tx_bytes);
            ptx::tma_load_2d_u32(dst, &tmap_g, full_bar, kb * BKE, col0, ptx::kEvictFirst);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (stats_on) atomicAdd(cfg.stats + 7, (unsigned long long)w_empty);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA)
    // The whole warp walks the loop (uniform control flow); one elected lane issues the tcgen05 instructions.
    // Per k-block: one try_wait, four MMAs whose descriptors differ by an immediate, one commit.
    if (leader) {
      const uint32_t idesc = ptx::make_idesc_bf16(CG * TM, TN);
      ptx::mbar_wait(&bars->a_full, 0);
      ptx::tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      long long w_tmem = 0, w_full = 0;
      const long long m_begin = stats_on ? clock64() : 0;
      const uint32_t desc_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(ring));
      const uint32_t desc_step = cfg.stage_bytes >> 4;
      uint32_t desc_lo = desc_lo0;
      const uint32_t a_tmem0 = tmem_base + a_col0;
      const bool issuer = ptx::elect_one();
      for (int t = 0; t < ntiles; ++t) {
        const int as = ACC == 2 ? (t & 1) : 0;
        const uint32_t aphase = (uint32_t)(ACC == 2 ? (t >> 1) : t) & 1u;
        long long c0 = stats_on ? clock64() : 0;
        ptx::mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        if (stats_on) w_tmem += clock64() - c0;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * TN);
        uint32_t a_tmem = a_tmem0;
        for (int kb = 0; kb < nkb; ++kb, a_tmem += BKE / 2) {
          c0 = stats_on ? clock64() : 0;
          if (!(debug & 4)) ptx::mbar_wait(&bars->full[stage], phase);
          if (stats_on) w_full += clock64() - c0;
          if (issuer) {
            if (!(debug & 1)) {
              const int reps = (debug & 16) ? 2 : 1;  // bit 4: issue every MMA twice (issue-rate experiment)
              for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
                for (int k = 0; k < BKE / UMMA_K; ++k) {
                  const uint64_t db = ptx::sw128_desc(desc_lo + (uint32_t)(k * UMMA_K * 2 / 16));
                  const uint32_t acc = (k != 0 || kb != 0) ? 1u : 0u;
                  if (CG == 2) ptx::mma_bf16_ts_2sm(tmem_d, a_tmem + k * (UMMA_K / 2), db, idesc, acc);
                  else ptx::mma_bf16_ts(tmem_d, a_tmem + k * (UMMA_K / 2), db, idesc, acc);
                }
              }
            }
            if (!(debug & 4)) {
              if (CG == 2) ptx::tc_commit_2sm(&bars->empty[stage], 3);  // frees this slot in BOTH CTAs
              else ptx::tc_commit(&bars->empty[stage]);
            }
          }
          desc_lo += desc_step;
          if (++stage == stages) { stage = 0; phase ^= 1; desc_lo = desc_lo0; }
        }
        if (issuer) {
          if (CG == 2) ptx::tc_commit_2sm(&bars->tmem_full[as], 3);
          else ptx::tc_commit(&bars->tmem_full[as]);
        }
        __syncwarp();
      }
      if (stats_on && issuer) {
        atomicAdd(cfg.stats + 0, (unsigned long long)(clock64() - m_begin));
        atomicAdd(cfg.stats + 1, (unsigned long long)w_tmem);
        atomicAdd(cfg.stats + 2, (unsigned long long)w_full);
        atomicAdd(cfg.stats + 8, 1ull);
      }
    }
  } else {
    // ===================================================================== selection (every CTA, 8 warps)
    constexpr int L = 32 * E;
    const int ew = warp - 2;            // 0..7
    const int grp = ew >> 2;            // column-chunk parity this thread owns
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int rloc = quarter * 32 + lane;
    const bool row_valid = row0 + rloc < p.nq;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;

    // ---- query tile -> TMEM: lane = row, column a_col0 + kb*32 + w holds elements (kb*64 + 2w, +1) ---------
    {
      const uint4* qrow = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.q) +
                                                         (row0 + rloc) * (int64_t)p.d);
      for (int kb = grp; kb < nkb; kb += 2) {
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int e0 = kb * BKE + c * 8;  // first element of this 16-byte piece (d % 8 == 0: all-or-nothing)
          uint4 x = make_uint4(0u, 0u, 0u, 0u);
          if (row_valid && e0 < p.d) x = __ldg(qrow + (e0 >> 3));
          v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        }
        ptx::tmem_st_32x32(tmem_base + lane_addr + a_col0 + (uint32_t)kb * (BKE / 2), v);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&bars->a_full);
        else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->a_full), 0));
      }
    }

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
    const int et = threadIdx.x - 64;
    const int nchunks = TN / 32;
    const int seed_stride = p.seed_stride;   // > 0: maxima mode of the threshold-seeding pass (select.cuh)
    SeedRun seed_run;
    seed_run.best = -INFINITY;
    seed_run.col = 0u;
    seed_run.since = 0;
    long long e_wait = 0, e_slow = 0;
    unsigned long long n_slow = 0;
    const long long e_begin = stats_on ? clock64() : 0;

    for (int t = 0; t < ntiles; ++t) {
      const int as = ACC == 2 ? (t & 1) : 0;
      const uint32_t aphase = (uint32_t)(ACC == 2 ? (t >> 1) : t) & 1u;
      const int64_t col0 = c_begin + (int64_t)t * TN;
      float* gst = gs + (t & 1) * kMaxTileN;
      if (kL2) {
        if (et < TN) {
          int64_t c = col0 + et;
          if (c >= p.ng) c = p.ng - 1;
          gst[et] = __ldg(p.gsq + c);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      const uint32_t tau_peek = seed_stride > 0 ? 0u : peek_tau(tau_row);  // L2 round trip hidden behind the barrier wait
      const long long cw = stats_on ? clock64() : 0;
      ptx::mbar_wait(&bars->tmem_full[as], aphase);
      if (stats_on) e_wait += clock64() - cw;
      ptx::tc_fence_after();
      apply_tau<kL2>(st, tau_peek);
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(as * TN);

      PendingHits pend;
      pend.n = 0;
      if (seed_stride > 0)
        seed_tile_tmem<kL2>(st, seed_run, seed_stride, taddr, grp, 2, nchunks, col0, c_end, gst, qn, self_row, row_valid);
      else if (!(debug & 8))
        select_tile_tmem<E, kL2>(st, pend, taddr, grp, 2, nchunks, col0, c_end, gst, qn, self_row, p.self_mode, p.k,
                                 lane, tau_row, row_valid && !(debug & 2), stats_on, e_slow, n_slow);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&bars->tmem_empty[as]);
        else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tmem_empty[as]), 0));
      }
      if (seed_stride == 0)
        flush_pending_hits<E, kL2>(st, pend, self_row, p.self_mode, p.k, lane, tau_row, stats_on, e_slow);
    }
    if (seed_stride == kSeedWholeUnit && p.maxima != nullptr) {
      // one maximum per (unit, selection thread), written straight into the table seed_select_kernel reads
      // ([row][splits * 2], order-preserving encoding, 0 = none): no list, no filter_lists launch
      if (row_valid)
        p.maxima[(row0 + rloc) * ((int64_t)p.splits * 2) + (sp * 2 + grp)] =
            seed_run.best > -INFINITY ? f2ord(exact_score<kL2>(seed_run.best) + 0.0f) : 0u;
    } else if (seed_stride > 0 && seed_run.since > 0) {
      seed_flush<kL2>(st, seed_run, row_valid);
    }
    if (stats_on && lane == 0) {
      atomicAdd(cfg.stats + 3, (unsigned long long)(clock64() - e_begin));
      atomicAdd(cfg.stats + 4, (unsigned long long)e_wait);
      atomicAdd(cfg.stats + 5, (unsigned long long)e_slow);
      atomicAdd(cfg.stats + 6, n_slow);
      atomicAdd(cfg.stats + 9, 1ull);
      atomicAdd(cfg.stats + 10, (unsigned long long)ntiles * ((nchunks + 1 - grp) / 2));
    }
    // end of unit: the list stays unordered; the unit merge reads `cnt` keys from it
    p.counts[vunit * TM + rloc] = row_valid ? st.cnt : 0;
  }

  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();  // nobody frees TMEM / exits while a peer may touch it
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

int env_int_ts(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

template <int CG, int E>
int launch_e(const SearchParams& p, cudaStream_t stream) {
  TsCfg cfg;
  cfg.nkb = (p.d + BKE - 1) / BKE;
  cfg.tn = ts_tile_cols(p.d);
  if (cfg.tn == 0) {
    set_error("internal: d=%d does not fit the TMEM-resident query tile", p.d);
    return KNN_E_INVALID;
  }
  cfg.acc_stages = (kTmemCols - cfg.nkb * (BKE / 2)) >= 2 * cfg.tn ? 2 : 1;
  const int rows_cta = cfg.tn / CG;
  cfg.stage_bytes = (uint32_t)rows_cta * BKE * 2;
  const size_t fixed = sizeof(float) * 2 * kMaxTileN + sizeof(TsBarriers);
  int stages = (int)((kSmemBudget - fixed) / cfg.stage_bytes);
  cfg.stages = stages > kMaxStages ? kMaxStages : stages;
  // one CTA per unit = HBM-bound streaming: ~128 KB in flight per SM is the sweet spot (tools/tma_probe.cu: 8 slots
  // of 16 KB stream at 7.3 TB/s, 13 slots at 6.5 TB/s)
  if (CG == 1 && cfg.stages > 8) cfg.stages = 8;
  if (const int want = env_int_ts("KNN_TS_STAGES")) {
    if (want >= 2 && want < cfg.stages) cfg.stages = want;
  }
  cfg.debug = env_int_ts("KNN_TS_DEBUG");
  cfg.stats = debug_stats_buffer();
  const size_t smem = (size_t)cfg.stages * cfg.stage_bytes + fixed;

  CUtensorMap tg;
  int rc = make_tmap_bf16_rows(&tg, p.g, p.ng, p.d, rows_cta);
  if (rc != KNN_OK) return rc;

  dim3 grid((unsigned)p.qblocks, (unsigned)p.splits);
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = grid;
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  const bool diag = cfg.stats != nullptr || cfg.debug != 0;
  if (p.metric == KNN_L2) {
    auto kern = search_bf16_ts_kernel<CG, E, true, false>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KNN_CHECK_CUDA(cudaLaunchKernelEx(&lc, kern, tg, p, cfg));
  } else if (diag && E == 8) {  // diagnostics build exists for the k <= 128, similarity instantiation only
    auto kern = search_bf16_ts_kernel<CG, 8, false, true>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KNN_CHECK_CUDA(cudaLaunchKernelEx(&lc, kern, tg, p, cfg));
  } else {
    auto kern = search_bf16_ts_kernel<CG, E, false, false>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KNN_CHECK_CUDA(cudaLaunchKernelEx(&lc, kern, tg, p, cfg));
  }
  KNN_LAUNCHED();
  return KNN_OK;
}

template <int CG>
int launch_cg(const SearchParams& p, cudaStream_t stream) {
  switch (p.kp) {
    case 32: return launch_e<CG, 2>(p, stream);
    case 64: return launch_e<CG, 4>(p, stream);
    case 128: return launch_e<CG, 8>(p, stream);
    case 256: return launch_e<CG, 16>(p, stream);
    default: set_error("unsupported padded k %d", p.kp); return KNN_E_UNSUPPORTED;
  }
}

}  // namespace

// Gallery rows per tile when the query tile lives in TMEM: 512 columns = D/2 (A, rounded up to whole k-blocks)
// + one or two accumulator stages of 128 columns.  0 = does not fit (use the shared-memory-A kernels).
int ts_tile_cols(int d) {
  const int a_cols = ((d + BKE - 1) / BKE) * (BKE / 2);
  return kTmemCols - a_cols >= 128 ? 128 : 0;
}

int launch_search_bf16_ts(const SearchParams& p, cudaStream_t stream) {
  if (p.groups != 2) {
    set_error("internal: the TMEM-resident kernel writes 2 candidate lists per row and split");
    return KNN_E_INVALID;
  }
  if (p.qblocks > 1) {
    if (p.qblocks % 2 != 0) {
      set_error("internal: CTA pairs need an even number of 128-row query blocks");
      return KNN_E_INVALID;
    }
    return launch_cg<2>(p, stream);
  }
  return launch_cg<1>(p, stream);
}

}  // namespace knn
