// bf16 distance + fused top-k on CTA PAIRS (tcgen05 cta_group::2) -- the tensor-bound kernel.
//
// A cluster of two CTAs (one TPC) owns 256 query rows and streams its gallery split in tiles of 256 rows:
//   * each CTA keeps ITS 128 query rows (UMMA A operand half) in shared memory -- resident for the whole unit when
//     D <= 512 (128 KB), otherwise streamed k-block by k-block next to B;
//   * each CTA TMA-loads HALF of every gallery tile (128 rows x 64 k, 16 KB per stage) into its own smem, so the
//     L2 -> SM traffic per SM is half of the single-CTA kernel's (32 B/cycle/SM at full tensor rate);
//   * one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16): the tensor cores of both
//     SMs read A/B halves from both shared memories; each CTA's TMEM receives its own 128 rows x 256 columns;
//   * smem slots and TMEM stages are recycled through mbarriers: TMA of both CTAs completes on the leader's `full`
//     barrier, tcgen05.commit multicasts to the `empty` / `tmem_full` barriers of both CTAs, the epilogue warps of
//     both CTAs arrive (remotely for the peer) on the leader's `tmem_empty`.
// Epilogue/selection is identical to search_tc.cu: thread i of warps 2..5 owns TMEM lane i = one query row.
#include <stdlib.h>
#include "select.cuh"
#include "ptx.cuh"
#include "kernels.h"

namespace knn {

namespace {

constexpr int TM = 128;        // query rows per CTA (256 per pair)
constexpr int TN = 256;        // gallery rows per pair tile
constexpr int TNH = 128;       // gallery rows loaded by each CTA
constexpr int BKE = 64;        // bf16 elements per k-block (128 B = one swizzle row)
constexpr int UMMA_K = 16;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 8;  // two per TMEM lane quarter: each row is owned by two threads (alternate chunks)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kMaxStages = 12;
constexpr uint32_t KB_BYTES = TM * BKE * 2;  // 16 KB: one k-block of 128 rows (A half or B half)
constexpr size_t kSmemBudget = 232448;       // 227 KB opt-in limit per CTA

struct alignas(8) PairBarriers {
  uint64_t full[kMaxStages];        // leader only: 2 arrivals (both producers) + tx bytes of both CTAs
  uint64_t empty[kMaxStages];       // every CTA: 1 arrival (multicast tcgen05.commit)
  uint64_t a_full;                  // leader only: resident query tile of both CTAs landed
  uint64_t tmem_full[kAccStages];   // every CTA: 1 arrival (multicast tcgen05.commit)
  uint64_t tmem_empty[kAccStages];  // leader only: 16 arrivals (8 selection warps x 2 CTAs)
  uint32_t tmem_base;
  uint32_t pad;
};

struct PairCfg {
  int stages;      // ring depth
  int resident;    // 1: query k-blocks stay in smem for the whole unit
  int nkb;         // k-blocks
  uint32_t a_bytes;      // resident A region
  uint32_t stage_bytes;  // bytes per ring stage in ONE CTA (B half [+ A k-block when streaming])
  int debug;             // measurement knob (KNN_PAIR_DEBUG=1): the epilogue skips the selection (results are
                         // garbage; isolates the TMA + MMA pipeline in timing experiments); 2 = fast path only
  int prefetch;          // L2 prefetch distance of the gallery stream in k-blocks (0 = off)
  int lockstep;          // soft lock-step window in tiles (0 = off), see SearchParams::progress
  int lock_ignore;       // peers further behind than this many tiles are late starters: not waited for
  int two;               // split mode with TWO products (q_hi.g_hi + q_lo.g_hi): the gallery lo part is neither loaded nor
                         // multiplied, a stage holds {G hi, Q hi, Q lo} (3 x 16 KB)
  unsigned long long* stats;  // diagnostics (KNN_PAIR_STATS=1): stall-cycle counters, see knn_debug_stats()
};

// kDiag = false is the production build: the stall counters and timing-experiment switches compile away.
// kSplit: the rows are bf16x3 split rows (knn_split_bf16x3; the filter of the tensor-core exact mode).  The score is
// qhi.ghi + qlo.ghi + qhi.glo: every ring stage holds the hi AND lo k-blocks of both operands (4 x 16 KB per CTA,
// tmap_q / tmap_g = hi parts, tmap_q2 / tmap_g2 = lo parts) and feeds 12 MMAs, so each part crosses L2 -> SM once
// per tile instead of once per product (as plain bf16 rows of 3x the width the kernel is L2 -> SM bound: 64 B/cycle/SM).
// kSplit = 0 (plain rows), 64 or 32 = elements per k-block in split mode (32: 64-byte swizzle, stages of 32 KB, so the
// ring is 7 deep instead of 3 -- 192 KB in flight instead of 128 KB).
template <int E, bool kL2, bool kDiag, int kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
search_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                        const __grid_constant__ CUtensorMap tmap_q2, const __grid_constant__ CUtensorMap tmap_g2,
                        SearchParams p, PairCfg cfg) {
  const bool stats_on = kDiag && cfg.stats != nullptr;
  const int debug = kDiag ? cfg.debug : 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;                                    // resident A: [nkb][128][64] bf16 (swizzled)
  uint8_t* ring = smem + cfg.a_bytes;                        // [stages][stage_bytes]
  float* gs = reinterpret_cast<float*>(ring + (size_t)cfg.stages * cfg.stage_bytes);    // [2][TN] (L2 metric)
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(gs + kAccStages * TN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int qb = blockIdx.x;  // 128-row query block of THIS CTA (pair = blockIdx.x >> 1)
  const int sp = blockIdx.y;
  const int64_t row0 = (int64_t)qb * TM;
  const int64_t c_begin = (int64_t)sp * p.split_len;
  const int64_t c_end = (c_begin + p.split_len < p.ng) ? c_begin + p.split_len : p.ng;
  const int ntiles = c_end > c_begin ? (int)((c_end - c_begin + TN - 1) / TN) : 0;
  const int nkb = cfg.nkb;
  const int stages = cfg.stages;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("b200knn: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_g);
    if (kSplit) {
      ptx::prefetch_tensormap(&tmap_q2);
      ptx::prefetch_tensormap(&tmap_g2);
    }
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&bars->full[s], 2);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 2);
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&bars->tmem_full[s], 1);
      ptx::mbar_init(&bars->tmem_empty[s], 2 * kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(&bars->tmem_base, kTmemCols);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barrier inits + TMEM allocation of BOTH CTAs visible before any remote traffic
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    // One thread; a single thread's dependent instruction stream costs ~5 cycles per instruction, so the loop
    // body is kept to a handful of instructions per slot (addresses advance by adds, nothing is recomputed).
    if (lane == 0) {
      if (cfg.resident) {
        const uint32_t a_full_leader = ptx::mapa(ptx::smem_u32(&bars->a_full), 0);
        for (int kb = 0; kb < nkb; ++kb)
          ptx::tma_load_2d_2sm(smem_a + (size_t)kb * KB_BYTES, &tmap_q, a_full_leader, kb * BKE, (int32_t)row0,
                               ptx::kEvictLast);
        if (leader) ptx::mbar_arrive_expect_tx(&bars->a_full, 2u * (uint32_t)nkb * KB_BYTES);
        else ptx::mbar_arrive_cluster(a_full_leader);
      }
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      const uint32_t ring_u32 = ptx::smem_u32(ring);
      const uint32_t full0 = ptx::mapa(ptx::smem_u32(&bars->full[0]), 0);
      const uint32_t tx_bytes = 2u * cfg.stage_bytes;
      // L2 prefetch cursor: runs cfg.prefetch k-blocks ahead of the ring's loads (experiment knob, default off)
      int pf_t = cfg.prefetch / nkb, pf_kb = cfg.prefetch % nkb;
      if (cfg.prefetch > 0) {
        for (int i = 0; i < cfg.prefetch && i < ntiles * nkb; ++i)
          ptx::tma_prefetch_2d(&tmap_g, (i % nkb) * BKE,
                               (int32_t)(c_begin + (int64_t)(i / nkb) * TN + (int64_t)rank * TNH));
      }
      // Soft lock-step: the pairs of one gallery split start together and stream the same rows; nothing keeps them
      // together, and once they drift further apart than L2 holds the gallery is re-read from HBM (2.6-3.8 x).  Every
      // 16 tiles the leader publishes its tile index and waits (bounded) while it is more than `lockstep` tiles ahead
      // of the slowest STARTED peer of its split that is not hopelessly behind (a late starter of a later wave).
      int32_t* prog = (leader && cfg.lockstep > 0 && p.progress != nullptr)
                          ? p.progress + (int64_t)sp * (p.qblocks / 2 + 1) : nullptr;
      const int npairs = p.qblocks / 2, my_pair = qb >> 1;
      int32_t col0 = (int32_t)(c_begin + (int64_t)rank * TNH);
      for (int t = 0; t < ntiles; ++t, col0 += TN) {
        if (prog != nullptr && (t & 15) == 0) {
          *reinterpret_cast<volatile int32_t*>(prog + my_pair) = t;
          for (int spin = 0; spin < 4000; ++spin) {
            int mn = t;
            for (int pp = 0; pp < npairs; ++pp) {
              const int v = *reinterpret_cast<volatile int32_t*>(prog + pp);
              if (v >= 0 && v >= t - cfg.lock_ignore && v < mn) mn = v;   // further behind: a late starter
            }
            if (t - mn <= cfg.lockstep) break;
            __nanosleep(256);
          }
        }
        for (int kb = 0; kb < nkb; ++kb) {
          if (cfg.prefetch > 0) {
            if (pf_t < ntiles) {
              ptx::tma_prefetch_2d(&tmap_g, pf_kb * (kSplit ? kSplit : BKE),
                                   (int32_t)(c_begin + (int64_t)pf_t * TN + (int64_t)rank * TNH));
              if (kSplit && !cfg.two)
                ptx::tma_prefetch_2d(&tmap_g2, pf_kb * kSplit, (int32_t)(c_begin + (int64_t)pf_t * TN + (int64_t)rank * TNH));
            }
            if (++pf_kb == nkb) { pf_kb = 0; ++pf_t; }
          }
          const long long c0 = stats_on ? clock64() : 0;
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          if (stats_on) w_empty += clock64() - c0;
          const uint32_t dst = ring_u32 + (uint32_t)stage * cfg.stage_bytes;
          const uint32_t full_bar = full0 + (uint32_t)stage * 8u;
          constexpr int kBk = kSplit ? kSplit : BKE;                 // elements per k-block
          constexpr uint32_t kKb = (uint32_t)(TM * kBk * 2);         // bytes of one 128-row k-block
          ptx::tma_load_2d_2sm_u32(dst, &tmap_g, full_bar, kb * kBk, col0, ptx::kEvictNormal);
          if (kSplit) {  // stage = {G hi, G lo, Q hi, Q lo}; two products: {G hi, Q hi, Q lo}
            const uint32_t qh = cfg.two ? kKb : 2 * kKb;
            if (!cfg.two) ptx::tma_load_2d_2sm_u32(dst + kKb, &tmap_g2, full_bar, kb * kBk, col0, ptx::kEvictNormal);
            ptx::tma_load_2d_2sm_u32(dst + qh, &tmap_q, full_bar, kb * kBk, (int32_t)row0, ptx::kEvictLast);
            ptx::tma_load_2d_2sm_u32(dst + qh + kKb, &tmap_q2, full_bar, kb * kBk, (int32_t)row0, ptx::kEvictLast);
          } else if (!cfg.resident)
            ptx::tma_load_2d_2sm_u32(dst + KB_BYTES, &tmap_q, full_bar, kb * BKE, (int32_t)row0, ptx::kEvictLast);
          if (leader) ptx::mbar_arrive_expect_tx_u32(full_bar, tx_bytes);
          else ptx::mbar_arrive_cluster(full_bar);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (prog != nullptr) *reinterpret_cast<volatile int32_t*>(prog + my_pair) = 0x7FFFFFFF;   // finished
      if (stats_on) atomicAdd(cfg.stats + 7, (unsigned long long)w_empty);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA)
    // The whole warp walks the loop (uniform control flow); one elected lane issues the tcgen05 instructions.
    // Per k-block: one try_wait, four MMAs whose descriptors differ by an immediate, one commit.
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * TM, TN);
      if (cfg.resident) {
        ptx::mbar_wait(&bars->a_full, 0);
        ptx::tc_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      long long w_tmem = 0, w_full = 0;
      const long long m_begin = stats_on ? clock64() : 0;
      const uint32_t b_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(ring));
      const uint32_t b_step = cfg.stage_bytes >> 4;
      const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
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
        uint32_t a_res = a_lo0;
        for (int kb = 0; kb < nkb; ++kb, a_res += KB_BYTES >> 4) {
          c0 = stats_on ? clock64() : 0;
          ptx::mbar_wait(&bars->full[stage], phase);
          if (stats_on) w_full += clock64() - c0;
          if (issuer) {
            if (kSplit) {
              constexpr int kBk = kSplit ? kSplit : BKE;
              constexpr uint32_t kPart = (uint32_t)(TM * kBk * 2) >> 4;  // descriptor units between the parts of a stage
              auto desc = [](uint32_t lo) { return kSplit == 32 ? ptx::sw64_desc(lo) : ptx::sw128_desc(lo); };
              const uint32_t qh = b_lo + (cfg.two ? kPart : 2 * kPart);   // query hi part; the lo part follows it
#pragma unroll
              for (int k = 0; k < kBk / UMMA_K; ++k) {   // q_hi . g_hi
                const uint32_t ko = (uint32_t)(k * UMMA_K * 2 / 16);
                ptx::mma_bf16_ss_2sm(tmem_d, desc(qh + ko), desc(b_lo + ko), idesc, (k != 0 || kb != 0) ? 1u : 0u);
              }
#pragma unroll
              for (int k = 0; k < kBk / UMMA_K; ++k) {   // q_lo . g_hi
                const uint32_t ko = (uint32_t)(k * UMMA_K * 2 / 16);
                ptx::mma_bf16_ss_2sm(tmem_d, desc(qh + kPart + ko), desc(b_lo + ko), idesc, 1u);
              }
              if (!cfg.two) {
#pragma unroll
                for (int k = 0; k < kBk / UMMA_K; ++k) {   // q_hi . g_lo
                  const uint32_t ko = (uint32_t)(k * UMMA_K * 2 / 16);
                  ptx::mma_bf16_ss_2sm(tmem_d, desc(qh + ko), desc(b_lo + kPart + ko), idesc, 1u);
                }
              }
            } else {
              const uint32_t a_lo = cfg.resident ? a_res : b_lo + (KB_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < BKE / UMMA_K; ++k) {
                const uint64_t da = ptx::sw128_desc(a_lo + (uint32_t)(k * UMMA_K * 2 / 16));
                const uint64_t db = ptx::sw128_desc(b_lo + (uint32_t)(k * UMMA_K * 2 / 16));
                ptx::mma_bf16_ss_2sm(tmem_d, da, db, idesc, (k != 0 || kb != 0) ? 1u : 0u);
              }
            }
            ptx::tc_commit_2sm(&bars->empty[stage], 3);  // frees this slot in BOTH CTAs when the MMAs retire
          }
          b_lo += b_step;
          if (++stage == stages) { stage = 0; phase ^= 1; b_lo = b_lo0; }
        }
        if (issuer) ptx::tc_commit_2sm(&bars->tmem_full[as], 3);  // accumulator tile complete (both epilogues)
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
    // ===================================================================== selection (both CTAs, 8 warps)
    // TMEM lane i = query row i is owned by TWO threads (warp groups 0 and 1) that take alternate 32-column
    // chunks of every tile and keep separate candidate lists; the unit merge sees splits * 2 lists per row.
    constexpr int L = 32 * E;
    const int grp = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int rloc = quarter * 32 + lane;
    const bool row_valid = row0 + rloc < p.nq;
    const int64_t vunit = ((int64_t)sp * 2 + grp) * p.qblocks + qb;
    const bool dense = p.dense_out != nullptr;   // dense mode: the scores are written out instead of selected
    RowState st;
    rowstate_init(st, dense ? nullptr : p.lists + ((vunit * TM + rloc) * (int64_t)L));
    uint32_t self_row = 0xFFFFFFFFu;
    float qn = 0.f;
    uint32_t* tau_row = nullptr;
    if (row_valid) {
      const int64_t sr = p.self_offset + row0 + rloc;
      if (p.self_mode != KNN_SELF_KEEP && sr >= 0 && sr < p.ng) self_row = (uint32_t)sr;
      if (kL2) qn = __ldg(p.qsq + row0 + rloc);
      if (!dense) tau_row = p.tau_global + row0 + rloc;
    }
    const int64_t dense_row = (dense && row_valid) ? row0 + rloc : -1;
    const int64_t dense_self = (p.self_mode != KNN_SELF_KEEP && dense_row >= 0) ? p.self_offset + dense_row : -1;
    const int et = threadIdx.x - 64;
    const int seed_stride = dense ? 0 : p.seed_stride;   // > 0: maxima mode of the threshold-seeding pass (select.cuh)
    SeedRun seed_run;
    seed_run.best = -INFINITY;
    seed_run.col = 0u;
    seed_run.since = 0;
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
      const uint32_t tau_peek = seed_stride > 0 ? 0u : peek_tau(tau_row);  // L2 round trip hidden behind the barrier wait
      const long long cw = stats_on ? clock64() : 0;
      ptx::mbar_wait(&bars->tmem_full[as], aphase);
      if (stats_on) e_wait += clock64() - cw;
      ptx::tc_fence_after();
      apply_tau<kL2>(st, tau_peek);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TN);
      PendingHits pend;
      pend.n = 0;
      if (dense)
        dense_store_tile_tmem<kL2>(taddr, grp, 2, TN / 32, col0, c_end, gst, qn, dense_row, dense_self, p.self_mode,
                                   p.dense_out, p.ng);
      else if (seed_stride > 0)
        seed_tile_tmem<kL2>(st, seed_run, seed_stride, taddr, grp, 2, TN / 32, col0, c_end, gst, qn, self_row, row_valid);
      else if (debug != 1)
        select_tile_tmem<E, kL2>(st, pend, taddr, grp, 2, TN / 32, col0, c_end, gst, qn, self_row, p.self_mode, p.k,
                                 lane, tau_row, row_valid && debug != 2, stats_on, e_slow, n_slow);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the accumulator stage is free again; the list work for this tile's hits comes after
        if (leader) ptx::mbar_arrive(&bars->tmem_empty[as]);
        else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tmem_empty[as]), 0));
      }
      if (!dense && seed_stride == 0)
        flush_pending_hits<E, kL2>(st, pend, self_row, p.self_mode, p.k, lane, tau_row, stats_on, e_slow);
    }
    if (seed_stride > 0 && seed_run.since > 0) seed_flush<kL2>(st, seed_run, row_valid);
    if (stats_on && lane == 0) {
      atomicAdd(cfg.stats + 3, (unsigned long long)(clock64() - e_begin));
      atomicAdd(cfg.stats + 4, (unsigned long long)e_wait);
      atomicAdd(cfg.stats + 5, (unsigned long long)e_slow);
      atomicAdd(cfg.stats + 6, n_slow);
      atomicAdd(cfg.stats + 9, 1ull);
      atomicAdd(cfg.stats + 10, (unsigned long long)ntiles * (TN / 64));
    }
    // end of unit: the list stays unordered; the unit merge reads `cnt` keys from it
    if (!dense) p.counts[vunit * TM + rloc] = row_valid ? st.cnt : 0;
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();  // nobody exits (or frees TMEM) while the peer may still touch its smem / barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

template <int E, int kSplit>
int launch_e(const SearchParams& p, cudaStream_t stream) {
  CUtensorMap tq, tg, tq2, tg2;
  const int dpart = kSplit ? p.d / 3 : p.d;  // columns of one operand part
  constexpr int kBk = kSplit ? kSplit : BKE;  // elements per k-block (32: 64-byte swizzle)
  constexpr uint32_t kKb = (uint32_t)(TM * kBk * 2);
  int rc = make_tmap_bf16_rows(&tq, p.q, p.nq, dpart, TM, p.d, kBk);
  if (rc != KNN_OK) return rc;
  rc = make_tmap_bf16_rows(&tg, p.g, p.ng, dpart, TNH, p.d, kBk);
  if (rc != KNN_OK) return rc;
  if (kSplit) {  // queries [hi | lo | hi], gallery [hi | hi | lo]: the lo parts
    const __nv_bfloat16* qb = reinterpret_cast<const __nv_bfloat16*>(p.q);
    const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(p.g);
    rc = make_tmap_bf16_rows(&tq2, qb + dpart, p.nq, dpart, TM, p.d, kBk);
    if (rc != KNN_OK) return rc;
    rc = make_tmap_bf16_rows(&tg2, gb + 2 * dpart, p.ng, dpart, TNH, p.d, kBk);
    if (rc != KNN_OK) return rc;
  } else {
    tq2 = tq;
    tg2 = tg;
  }

  PairCfg cfg;
  cfg.nkb = (dpart + kBk - 1) / kBk;
  const size_t fixed = sizeof(float) * kAccStages * TN + sizeof(PairBarriers);
  cfg.resident = (!kSplit && (size_t)cfg.nkb * KB_BYTES + 4 * (size_t)KB_BYTES + fixed <= kSmemBudget) ? 1 : 0;
  cfg.a_bytes = cfg.resident ? (uint32_t)cfg.nkb * KB_BYTES : 0u;
  cfg.two = (kSplit && p.split3 == 2) ? 1 : 0;
  cfg.stage_bytes = kSplit ? (cfg.two ? 3 : 4) * kKb : (cfg.resident ? KB_BYTES : 2 * KB_BYTES);
  int stages = (int)((kSmemBudget - fixed - cfg.a_bytes) / cfg.stage_bytes);
  cfg.stages = stages > kMaxStages ? kMaxStages : stages;
  cfg.debug = 0;
  if (const char* e = getenv("KNN_PAIR_DEBUG")) cfg.debug = atoi(e);
  cfg.stats = debug_stats_buffer();
  cfg.prefetch = 0;
  if (const char* e = getenv("KNN_PAIR_PREFETCH")) cfg.prefetch = atoi(e);
  // soft lock-step (see the producer loop): only when all pairs of a gallery split can be co-resident and the units are
  // long enough to drift (measured: 8192 x 50 M x 512 DRAM reads 131-195 GB -> 75 GB, +5 % throughput through the
  // higher clock under the power cap; 25 000 queries = 98 pairs per split on 74 pair slots: 25 % slower if forced)
  cfg.lockstep = 32;
  if (const char* e = getenv("KNN_PAIR_LOCKSTEP")) cfg.lockstep = atoi(e);
  {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t unit_tiles = (p.split_len + TN - 1) / TN;
    // units must be much longer than the "late starter" distance below: a pair that begins when an earlier CTA of its
    // split finishes must look hopelessly behind to the pairs still running, or they would wait for it
    int min_tiles = 2048;
    if (const char* e = getenv("KNN_PAIR_LOCKSTEP_MIN_TILES")) min_tiles = atoi(e);
    if (p.qblocks > sms || unit_tiles < min_tiles || p.progress == nullptr) cfg.lockstep = 0;
    cfg.lock_ignore = unit_tiles / 2 < 1024 ? (int)(unit_tiles / 2) : 1024;
  }
  if (const char* e = getenv("KNN_PAIR_LOCKSTEP_IGNORE")) cfg.lock_ignore = atoi(e);
  const size_t smem = (size_t)cfg.a_bytes + (size_t)cfg.stages * cfg.stage_bytes + fixed;

  dim3 grid((unsigned)p.qblocks, (unsigned)p.splits);  // qblocks is even: consecutive CTAs form the pair
  const bool diag = cfg.stats != nullptr || cfg.debug != 0;
  if (p.metric == KNN_L2) {
    auto kern = search_bf16_pair_kernel<E, true, false, kSplit>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, stream>>>(tq, tg, tq2, tg2, p, cfg);
  } else if (diag && E == 8) {  // diagnostics build exists for the k <= 128, similarity instantiation only
    auto kern = search_bf16_pair_kernel<8, false, true, kSplit>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, stream>>>(tq, tg, tq2, tg2, p, cfg);
  } else {
    auto kern = search_bf16_pair_kernel<E, false, false, kSplit>;
    KNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, stream>>>(tq, tg, tq2, tg2, p, cfg);
  }
  KNN_LAUNCHED();
  return KNN_OK;
}

template <int E>
int launch_e(const SearchParams& p, cudaStream_t stream) {
  static const int split_bk = [] {
    const char* e = getenv("KNN_SPLIT_BK");   // 64 | 32: k-block width of the split filter (experiment knob)
    const int v = e ? atoi(e) : 0;
    return v == 32 ? 32 : 64;
  }();
  if (!p.split3) return launch_e<E, 0>(p, stream);
  return split_bk == 32 ? launch_e<E, 32>(p, stream) : launch_e<E, 64>(p, stream);
}

}  // namespace

int launch_search_bf16_pair(const SearchParams& p, cudaStream_t stream) {
  if (p.qblocks % 2 != 0 || p.groups != 2) {
    set_error("internal: pair kernel needs an even number of 128-row query blocks and 2 lists per row and split");
    return KNN_E_INVALID;
  }
  switch (p.kp) {
    case 32: return launch_e<2>(p, stream);
    case 64: return launch_e<4>(p, stream);
    case 128: return launch_e<8>(p, stream);
    case 256: return launch_e<16>(p, stream);
    default: set_error("unsupported padded k %d", p.kp); return KNN_E_UNSUPPORTED;
  }
}

}  // namespace knn
