// Full-ranking metrics without the N x N ranking: rank of every RELEVANT gallery row in the full ranking of a query.
//
// The reference ranks whole score rows (torch.argsort / np.argsort / topk(N-1): test.py:1090, train.py:409,455,
// nih_multilabel_training.py:85-95, test.py:962) only to read off WHERE the relevant rows ended up: every full-ranking
// AP variant (compute_ap test.py:58-92, the rank-by-rank AP of test.py:970-981 / train.py:424-433, sklearn's
// average_precision_score train.py:473, nih_multilabel_training.py:95) is a function of the 0-based ranks of the
// positives (plus, for sklearn's tie-grouped curve, where the runs of equal scores begin and end).  This file computes
// exactly that from a dense block of score rows, one CTA per row, without sorting the row through global memory passes
// of a comparison sort and without an [N, N] index matrix:
//
//   1. bins    a monotone map score -> bin (linear over the range of a strided sample of the row, clamped) and a
//              shared-memory histogram of the row;
//   2. scatter an exclusive scan turns the histogram into bin offsets (best bin first) and every item is written as a
//              64-bit key (order-preserving score bits | inverted row | relevance bit) into its bin's range of a
//              per-CTA scratch row (L2 resident) -- equal scores always share a bin;
//   3. refine  bins holding more than `leaf_max` items (massive ties, outliers) are re-binned on the FULL key (score
//              then row, all keys distinct) until every leaf is small -- the rare path;
//   4. resolve every item counts the keys of its leaf that beat it: rank = leaf offset + that count.  Relevant items
//              write their rank at their index among the relevant ones, so the output row is the ascending list of
//              the positives' ranks.  In tie mode every item is also written to its rank position (a fully sorted
//              row) and a scan over it yields, per positive, the end of its run of equal scores and the number of
//              distinct score values above it -- what sklearn's threshold curve needs.
//
// Order: best score first (largest, or smallest with largest_first = 0), ties by ascending gallery row -- the order of
// knn_search / knn_rank_rows, so the ranks equal the positions a full knn_rank_rows ranking would give.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace knn {
namespace {

#ifndef RP_THREADS
#define RP_THREADS 512
#endif
constexpr int kThreads = RP_THREADS;
constexpr int kCtasPerSm = 1024 / kThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBins = 16384;
constexpr int kSubBins = 4096;
constexpr int kWorkCap = 1024;   // pending oversized segments (disjoint, each > leaf_max >= N / 512 items)
constexpr int kSample = 1024;

constexpr int kRelSingle = 0, kRelJaccardF32 = 1, kRelJaccardF64 = 2, kRelAny = 3;

struct RankParams {
  const float* scores;
  int64_t ld_scores;
  int64_t nq, ng;
  int largest;
  int rel_mode;
  const int64_t* qlab;   // labels (single) or 64-bit masks (multi-label)
  const int64_t* glab;
  double thr;
  int64_t self_offset;   // gallery row of query 0
  int drop_self;
  const int64_t* q_group;   // nullable pair: gallery rows whose group equals the query's are not ranked either
  const int64_t* g_group;   // (fusion_eval/metrics.py:67 drops every row that shares the query's image path)
  int ties;              // also produce pos_ge / pos_tgroup / ngroups (through the fully sorted row)
  int nbins;
  int leaf_max;
  int nslabs;            // the row is scattered / resolved in this many slabs of bins (live scratch stays L2 resident)
  uint64_t* scratch;     // [grid][ng]
  uint64_t* tmp;         // [grid][ng]
  unsigned int* next_row;   // rows are claimed dynamically: a CTA slowed down by a co-running kernel takes fewer
  int32_t* pos_ranks;
  int64_t ld_out;
  int32_t* pos_ge;
  int32_t* pos_tgroup;
  int32_t* npos;
  int32_t* nranked;
  int32_t* ngroups;
};

// key = [ord(score) : 32][0x3FFFFFFF - row : 30][leaf-start flag : 1][relevant : 1]; the order is (key >> 2) descending
__device__ __forceinline__ uint64_t rp_key(float v, uint32_t row, uint32_t rel) {
  return ((uint64_t)f2ord(v) << 32) | ((uint64_t)(0x3FFFFFFFu - row) << 2) | (uint64_t)rel;
}
__device__ __forceinline__ float rp_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }

struct BinMap {
  float lo, scale;
  int nb;
};
// Monotone non-increasing in v (bin 0 = best scores).  The same instruction sequence runs in every pass (explicit
// round-to-nearest intrinsics: no contraction), so an item always lands in the same bin.
__device__ __forceinline__ int bin_of(float v, const BinMap& m) {
  int b;
  if (v != v) {
    b = (__float_as_uint(v) & 0x80000000u) ? 0 : m.nb - 1;   // where f2ord puts the NaNs
  } else {
    float t = __fmul_rn(__fsub_rn(v, m.lo), m.scale);
    t = fminf(fmaxf(t, 0.0f), (float)(m.nb - 1));             // fmaxf(NaN, 0) = 0 (inf * 0)
    b = (int)t;
  }
  return m.nb - 1 - b;
}

constexpr int kItems = 4;                       // consecutive positions per thread and resolve step
constexpr int kChunk = kThreads * kItems;       // 2048 positions per resolve step
constexpr int kHalo = 256;                      // staged on either side of the chunk: a leaf of <= 256 items that holds
constexpr int kStage = kChunk + kThreads;        // a position of the chunk never leaves the staged window; kStage is a
                                                // multiple of kThreads, the right halo takes the rest

struct Shared {
  // refinement: sub-bin counters / offsets (2 x 16 KB); resolve: the staged window of keys -- never live together
  union {
    struct { uint32_t sub_cnt[kSubBins]; uint32_t sub_start[kSubBins]; } r;
    unsigned long long stage[kStage];
  } u;
  uint32_t work_start[kWorkCap];
  uint32_t work_cnt[kWorkCap];
  unsigned long long red[2 * kWarps];
  float fred[2 * kWarps];
  uint32_t work_head, work_tail;
  uint32_t row_claim;
  uint32_t slab_bin[34];    // first bin of every slab (+ nbins at the end)
  uint16_t plist[kChunk];   // chunk offsets of the relevant positions (resolve without ties)
  uint8_t rel_table[65 * 65];
};

// Block scans over one value per thread; `buf` holds kWarps entries.  Every thread of the CTA must call.
// add: returns the EXCLUSIVE prefix, writes the block total.
__device__ __forceinline__ unsigned long long block_excl_add(unsigned long long v, unsigned long long* buf,
                                                             unsigned long long& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl += o;
  }
  __syncthreads();   // buf may still be read by the previous call
  if (lane == 31) buf[warp] = incl;
  __syncthreads();
  unsigned long long pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const unsigned long long x = buf[w];
    if (w < warp) pre += x;
    tot += x;
  }
  total = tot;
  return pre + incl - v;
}
// max: returns the EXCLUSIVE prefix maximum (0 for the first thread), writes the block maximum.
__device__ __forceinline__ unsigned long long block_excl_max(unsigned long long v, unsigned long long* buf,
                                                             unsigned long long& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off && o > incl) incl = o;
  }
  unsigned long long excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
  if (lane == 0) excl = 0;
  __syncthreads();
  if (lane == 31) buf[warp] = incl;
  __syncthreads();
  unsigned long long pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const unsigned long long x = buf[w];
    if (w < warp && x > pre) pre = x;
    if (x > tot) tot = x;
  }
  total = tot;
  return excl > pre ? excl : pre;
}

// Exclusive scan of arr[0..n) in shared memory, n a multiple of kThreads: every warp owns n / kWarps consecutive
// entries and walks them 32 at a time (conflict-free), warp totals are combined through `buf`.  Entries whose count
// exceeds leaf_max are appended to the work list as (base + offset, count).  Every thread of the CTA must call.
struct WorkList {
  uint32_t* start;
  uint32_t* cnt;
  uint32_t* tail;
};
__device__ uint32_t smem_exclusive_scan(uint32_t* arr, int n, unsigned long long* buf, uint32_t base,
                                        uint32_t leaf_max, const WorkList& wl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = n / kWarps;
  uint32_t* mine = arr + warp * per_warp;
  uint32_t local = 0;
  for (int c = lane; c < per_warp; c += 32) local += mine[c];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, off);
  __syncthreads();   // buf may still be read by an earlier reduction
  if (lane == 0) buf[warp] = local;
  __syncthreads();
  uint32_t run = 0, total = 0;
  for (int w = 0; w < kWarps; ++w) {
    if (w < warp) run += (uint32_t)buf[w];
    total += (uint32_t)buf[w];
  }
  for (int c = 0; c < per_warp; c += 32) {
    const uint32_t v = mine[c + lane];
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off) incl += o;
    }
    const uint32_t excl = run + incl - v;
    mine[c + lane] = excl;
    if (v > leaf_max) {
      const uint32_t slot = atomicAdd(wl.tail, 1u);
      wl.start[slot % kWorkCap] = base + excl;
      wl.cnt[slot % kWorkCap] = v;
    }
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  __syncthreads();
  return total;
}

// Runs of equal scores over the fully sorted row `sorted` (nr keys): per positive its rank, the end of its run (ge = items
// with a score >= its own) and the number of distinct score values above it (tgroup); npos / ngroups of the row.
__device__ void ties_pass(const RankParams& p, const uint64_t* tmp, int64_t nr, int64_t row, unsigned long long* red) {
  const int tid = threadIdx.x;
  int32_t* out_rank = p.pos_ranks + row * p.ld_out;
  struct { unsigned long long* red; } sh{red};
  // ---- 5. runs of equal scores over the sorted row: per positive its rank, the end of its run (ge = items with a
  // score >= its own) and the number of distinct score values above it (tgroup)
  int32_t* out_ge = p.pos_ge + row * p.ld_out;
  int32_t* out_tg = p.pos_tgroup + row * p.ld_out;
  unsigned long long carry_add = 0, carry_max = 0;   // (positives << 32 | run starts) so far; last run start so far
  for (int64_t base = 0; base < nr; base += kChunk) {
    uint32_t rl[kItems], st[kItems];
    unsigned long long loc = 0;
    const int64_t i0 = base + tid * kItems;
    uint32_t prev_ord = (i0 > 0 && i0 - 1 < nr) ? (uint32_t)(tmp[i0 - 1] >> 32) : 0u;
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      const int64_t i = i0 + u;
      rl[u] = 0; st[u] = 0;
      if (i < nr) {
        const unsigned long long ki = tmp[i];
        rl[u] = (uint32_t)(ki & 1ull);
        st[u] = (i == 0) || (prev_ord != (uint32_t)(ki >> 32));
        prev_ord = (uint32_t)(ki >> 32);
      }
      loc += ((unsigned long long)rl[u] << 32) | st[u];
    }
    unsigned long long tot_add, tot_max;
    unsigned long long run = carry_add + block_excl_add(loc, sh.red, tot_add);   // before this thread's items
    // this thread's last run start as ((position + 1) << 32 | positives before it): the later start wins the max
    unsigned long long mine = 0, r2 = run;
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      if (st[u]) mine = ((unsigned long long)(i0 + u + 1) << 32) | (uint32_t)(r2 >> 32);
      r2 += ((unsigned long long)rl[u] << 32) | st[u];
    }
    unsigned long long last = block_excl_max(mine, sh.red, tot_max);
    if (carry_max > last) last = carry_max;
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      const int64_t i = i0 + u;
      if (i < nr) {
        const uint32_t pidx = (uint32_t)(run >> 32);   // positives before position i
        if (st[u]) {
          // a run starting at i > 0 closes the previous run: its positives P(start) .. pidx - 1 get ge = i
          if (i > 0)
            for (uint32_t t = (uint32_t)last; t < pidx; ++t) out_ge[t] = (int32_t)i;
          last = ((unsigned long long)(i + 1) << 32) | pidx;
        }
        if (rl[u]) {
          out_rank[pidx] = (int32_t)i;
          out_tg[pidx] = (int32_t)((uint32_t)run + st[u] - 1);   // run starts in [0, i] minus one
        }
      }
      run += ((unsigned long long)rl[u] << 32) | st[u];
    }
    carry_add += tot_add;
    if (tot_max > carry_max) carry_max = tot_max;
  }
  if (tid == 0) {
    const uint32_t total_pos = (uint32_t)(carry_add >> 32);
    for (uint32_t t = (uint32_t)carry_max; t < total_pos; ++t) out_ge[t] = (int32_t)nr;   // the last run
    p.npos[row] = (int32_t)total_pos;
    if (p.ngroups) p.ngroups[row] = (int32_t)(uint32_t)carry_add;
  }
}

#ifdef RP_TIMING
__device__ unsigned long long rp_timing[8];
#define RP_TICK(slot)                                                        \
  do {                                                                       \
    __syncthreads();                                                         \
    if (threadIdx.x == 0) {                                                  \
      const long long now_ = clock64();                                      \
      atomicAdd(&rp_timing[slot], (unsigned long long)(now_ - tick_));       \
      tick_ = now_;                                                          \
    }                                                                        \
  } while (0)
#else
#define RP_TICK(slot)
#endif

__global__ void __launch_bounds__(kThreads, kCtasPerSm) rank_positives_kernel(RankParams p) {
#ifdef RP_TIMING
  long long tick_ = clock64();
#endif
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* bins = reinterpret_cast<uint32_t*>(smem_raw);                       // [nbins]: counts -> offsets -> ends
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw + (size_t)p.nbins * sizeof(uint32_t));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = p.ng;
  uint64_t* scratch = p.scratch + (int64_t)blockIdx.x * n;
  uint64_t* tmp = p.tmp + (int64_t)blockIdx.x * n;

  // relevance of a multi-label pair depends on (|a & b|, |a | b|) only: evaluate the reference's arithmetic once per
  // pair of counts (fp32: train.py:462-466, test.py:956-965; fp64: evaluate_nih_zilliz.py:12-17; any: test.py:1045)
  if (p.rel_mode != kRelSingle) {
    for (int e = tid; e < 65 * 65; e += kThreads) {
      const int inter = e / 65, uni = e % 65;
      uint8_t r;
      if (p.rel_mode == kRelJaccardF32) r = __fdiv_rn((float)inter, __fadd_rn((float)uni, 1e-8f)) > (float)p.thr;
      else if (p.rel_mode == kRelJaccardF64) r = __ddiv_rn((double)inter, __dadd_rn((double)uni, 1e-8)) > p.thr;
      else r = inter > 0;
      sh.rel_table[e] = r;
    }
  }
  __syncthreads();

  while (true) {
    if (tid == 0) {
      const unsigned int c = atomicAdd(p.next_row, 1u);
      sh.row_claim = c < p.nq ? c : 0xFFFFFFFFu;
    }
    __syncthreads();
    if (sh.row_claim == 0xFFFFFFFFu) break;
    const int64_t row = sh.row_claim;
    const float* srow = p.scores + row * p.ld_scores;
    const int64_t self = p.drop_self ? p.self_offset + row : -1;
    const int64_t ql = p.qlab[row];
    const bool by_group = p.q_group != nullptr;
    const int64_t qgrp = by_group ? p.q_group[row] : 0;
    auto dropped = [&](int64_t g) -> bool { return g == self || (by_group && p.g_group[g] == qgrp); };
    auto rel_of = [&](int64_t lab) -> uint32_t {   // lab = the gallery row's label / label mask
      if (p.rel_mode == kRelSingle) return lab == ql ? 1u : 0u;
      const uint64_t a = (uint64_t)ql, b = (uint64_t)lab;
      return sh.rel_table[__popcll(a & b) * 65 + __popcll(a | b)];
    };
    auto canon = [&](float v) -> float {
      if (!p.largest) v = -v;
      return v + 0.0f;   // -0.0 -> +0.0, as make_key
    };

    // ---- 0. range of a strided sample -> bin map
    {
      const int64_t m = n < kSample ? n : kSample;
      float lo = INFINITY, hi = -INFINITY;
      for (int64_t j = tid; j < m; j += kThreads) {
        const int64_t g = j * n / m;
        if (dropped(g)) continue;
        const float v = canon(srow[g]);
        if (v - v == 0.0f) {   // finite
          lo = fminf(lo, v);
          hi = fmaxf(hi, v);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, off));
        hi = fmaxf(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, off));
      }
      if (lane == 0) { sh.fred[warp] = lo; sh.fred[kWarps + warp] = hi; }
      __syncthreads();
      lo = sh.fred[0]; hi = sh.fred[kWarps];
      for (int w = 1; w < kWarps; ++w) { lo = fminf(lo, sh.fred[w]); hi = fmaxf(hi, sh.fred[kWarps + w]); }
      __syncthreads();
      float mlo, mscale;
      if (!(hi > lo)) { mlo = (lo - lo == 0.0f) ? lo : 0.0f; mscale = 0.0f; }
      else {
        const float w = hi - lo;
        mlo = lo - 0.02f * w;
        mscale = (float)p.nbins / (1.04f * w);
        if (!(mscale - mscale == 0.0f)) mscale = 0.0f;   // overflowed range
      }
      if (tid == 0) { sh.fred[0] = mlo; sh.fred[1] = mscale; }
      __syncthreads();
    }
    BinMap bm;
    bm.lo = sh.fred[0]; bm.scale = sh.fred[1]; bm.nb = p.nbins;
    RP_TICK(0);

    // ---- 1. histogram.  Streaming passes read the row 4 consecutive scores per thread (one 16-byte load when the row
    // is aligned), two such groups in flight.
    for (int b = tid; b < p.nbins; b += kThreads) bins[b] = 0;
    if (tid == 0) { sh.work_head = 0; sh.work_tail = 0; }
    __syncthreads();
    const bool vec = (reinterpret_cast<uintptr_t>(srow) & 15) == 0;
    const bool any_drop = by_group || self >= 0;
    auto load4 = [&](int64_t g, float* v) {   // scores g .. g+3 (rows past the end: 0)
      if (vec && g + 3 < n) {
        const float4 x = __ldcs(reinterpret_cast<const float4*>(srow + g));   // streamed: keep L2 for the scratch rows
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (g + u < n) ? __ldcs(srow + g + u) : 0.0f;
      }
    };
    for (int64_t g0 = (int64_t)tid * 4; g0 < n; g0 += 8 * kThreads) {
      float v[8];
      const int64_t g1 = g0 + 4 * kThreads;
      load4(g0, v);
      if (g1 < n) load4(g1, v + 4);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t g = (u < 4 ? g0 : g1) + (u & 3);
        if (g < n && !(any_drop && dropped(g))) atomicAdd(&bins[bin_of(canon(v[u]), bm)], 1u);
      }
    }
    __syncthreads();
    RP_TICK(1);

    // ---- 2. offsets (bin 0 = best first); oversized bins go to the work list
    const WorkList wl{sh.work_start, sh.work_cnt, &sh.work_tail};
    const int64_t nr = smem_exclusive_scan(bins, p.nbins, sh.red, 0u, 0xFFFFFFFFu, wl);   // rows ranked
    // slab k = the bins whose first position is the first one >= k * nr / nslabs (whole bins: a leaf never straddles)
    if (tid <= p.nslabs) {
      uint32_t b = 0;
      if (tid == p.nslabs) b = (uint32_t)p.nbins;
      else if (tid > 0) {
        const uint32_t want = (uint32_t)(((int64_t)tid * nr) / p.nslabs);
        int lo_b = 0, hi_b = p.nbins;           // smallest b with bins[b] >= want
        while (lo_b < hi_b) {
          const int mid = (lo_b + hi_b) >> 1;
          if (bins[mid] >= want) hi_b = mid; else lo_b = mid + 1;
        }
        b = (uint32_t)lo_b;
      }
      sh.slab_bin[tid] = b;
    }
    __syncthreads();

    RP_TICK(2);
    int32_t* out_rank = p.pos_ranks + row * p.ld_out;
    unsigned long long carry = 0;   // positives in the slabs / chunks resolved so far
    for (int slab = 0; slab < p.nslabs; ++slab) {
    const int b0 = (int)sh.slab_bin[slab], b1 = (int)sh.slab_bin[slab + 1];
    if (b0 >= b1) continue;
    // positions [sb, se) of the ranking live in scratch[0 .. se - sb) while this slab is worked on
    const int64_t sb = bins[b0], se = b1 < p.nbins ? (int64_t)bins[b1] : nr;
    // bins of the slab holding more than leaf_max items go to the work list (bins[] still holds START offsets here)
    for (int b = b0 + tid; b < b1; b += kThreads) {
      const uint32_t c = (b + 1 < p.nbins ? bins[b + 1] : (uint32_t)nr) - bins[b];
      if (c > (uint32_t)p.leaf_max) {
        const uint32_t slot = atomicAdd(&sh.work_tail, 1u);
        sh.work_start[slot % kWorkCap] = bins[b];
        sh.work_cnt[slot % kWorkCap] = c;
      }
    }
    __syncthreads();
    // ---- scatter: bins[b] becomes the END of bin b (= start of bin b + 1)
    for (int64_t g0 = (int64_t)tid * 4; g0 < n; g0 += 8 * kThreads) {
      float v[8];
      const int64_t g1 = g0 + 4 * kThreads;
      load4(g0, v);
      if (g1 < n) load4(g1, v + 4);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t g = (u < 4 ? g0 : g1) + (u & 3);
        if (g < n) {
          const float c = canon(v[u]);
          const int b = bin_of(c, bm);
          if (b >= b0 && b < b1 && !(any_drop && dropped(g))) {
            const uint32_t slot = atomicAdd(&bins[b], 1u);
            scratch[slot - sb] = rp_key(c, (uint32_t)g, rel_of(p.glab[g]));
          }
        }
      }
    }
    __syncthreads();

    RP_TICK(3);
    // ---- 3. refine oversized segments on the full key (rare: massive ties / outliers)
    while (true) {
      const uint32_t head = sh.work_head, tail = sh.work_tail;
      if (head == tail) break;
      const uint32_t s0 = sh.work_start[head % kWorkCap], cnt = sh.work_cnt[head % kWorkCap];
      __syncthreads();
      if (tid == 0) sh.work_head = head + 1;
      unsigned long long kmin = ~0ull, kmax = 0ull;
      for (uint32_t i = tid; i < cnt; i += kThreads) {
        const unsigned long long k = scratch[s0 - sb + i] >> 2;
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xFFFFFFFFu, kmin, off), b = __shfl_xor_sync(0xFFFFFFFFu, kmax, off);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
      }
      if (lane == 0) { sh.red[warp] = kmin; sh.red[kWarps + warp] = kmax; }
      __syncthreads();
      kmin = sh.red[0]; kmax = sh.red[kWarps];
      for (int w = 1; w < kWarps; ++w) {
        kmin = sh.red[w] < kmin ? sh.red[w] : kmin;
        kmax = sh.red[kWarps + w] > kmax ? sh.red[kWarps + w] : kmax;
      }
      const unsigned long long span = kmax - kmin;   // > 0: keys are distinct and cnt > 1
      const int nbits = 64 - __clzll((long long)span);
      const int shift = nbits > 12 ? nbits - 12 : 0;
      const uint32_t top = (uint32_t)(span >> shift);          // sub-bin of the best key; sub' = top - sub
      uint32_t* sub_cnt = sh.u.r.sub_cnt;
      uint32_t* sub_start = sh.u.r.sub_start;
      for (int b = tid; b < kSubBins; b += kThreads) sub_cnt[b] = 0;
      __syncthreads();
      for (uint32_t i = tid; i < cnt; i += kThreads) {
        const unsigned long long k = scratch[s0 - sb + i] >> 2;
        atomicAdd(&sub_cnt[top - (uint32_t)((k - kmin) >> shift)], 1u);
      }
      __syncthreads();
      smem_exclusive_scan(sub_cnt, kSubBins, sh.red, s0, (uint32_t)p.leaf_max, wl);
      for (int b = tid; b < kSubBins; b += kThreads) sub_start[b] = sub_cnt[b];
      __syncthreads();
      for (uint32_t i = tid; i < cnt; i += kThreads) {
        const unsigned long long key = scratch[s0 - sb + i];
        const uint32_t sub = top - (uint32_t)(((key >> 2) - kmin) >> shift);
        const uint32_t slot = atomicAdd(&sub_cnt[sub], 1u);
        tmp[s0 + slot] = (key & ~2ull) | (slot == sub_start[sub] ? 2ull : 0ull);
      }
      __syncthreads();
      for (uint32_t i = tid; i < cnt; i += kThreads) scratch[s0 - sb + i] = tmp[s0 + i];
      __syncthreads();
    }

    RP_TICK(4);
    // ---- 4. resolve: exact position of an item inside its leaf = keys of the leaf that beat it.  A window of the
    // scratch row (the chunk plus a halo on either side) is staged in shared memory, so the leaf walks are shared-memory
    // reads; the leaf of a position is its bin's range [bins[b-1], bins[b]) -- every item of a leaf runs the same loop.
    // Without ties only the RELEVANT positions are resolved (compacted through the prefix scan that numbers them).
    unsigned long long pf[kStage / kThreads];   // the next window, loaded while the current one is being walked
    auto prefetch = [&](int64_t base) {
#pragma unroll
      for (int u = 0; u < kStage / kThreads; ++u) {
        const int64_t j = base - kHalo + u * kThreads + tid;
        pf[u] = (j >= sb && j < se) ? scratch[j - sb] : 0ull;
      }
    };
    prefetch(sb);
    for (int64_t base = sb; base < se; base += kChunk) {
      const int64_t w0 = base - kHalo;   // scratch position of stage[0]
      __syncthreads();                    // the previous chunk's walks are done
#pragma unroll
      for (int u = 0; u < kStage / kThreads; ++u) sh.u.stage[u * kThreads + tid] = pf[u];
      __syncthreads();
      if (base + kChunk < se) prefetch(base + kChunk);
      // rank of the item at chunk offset `off` (+ positives of its leaf: before it in the leaf / beating it)
      auto resolve = [&](int off, uint32_t& pos_before, uint32_t& pos_gt) -> uint32_t {
        const int64_t i = base + off;
        const unsigned long long ki = sh.u.stage[kHalo + off];
        const unsigned long long mine = ki >> 2;
        const int b = bin_of(rp_score(ki), bm);
        int64_t ls = b == 0 ? 0 : bins[b - 1], le = bins[b];
        uint32_t gt = 0;
        pos_before = 0; pos_gt = 0;
        if (le - ls > p.leaf_max) {
          // a refined bin: the leaf is delimited by the start flags (rare path, generic reads)
          auto key_at = [&](int64_t j) -> unsigned long long {
            const int64_t o = j - w0;
            return (o >= 0 && o < kStage) ? sh.u.stage[o] : (unsigned long long)scratch[j - sb];
          };
          int64_t a = i;
          while (a > ls && !(key_at(a) & 2ull)) --a;
          int64_t e = i + 1;
          while (e < le && !(key_at(e) & 2ull)) ++e;
          for (int64_t j = a; j < e; ++j) {
            const unsigned long long kj = key_at(j);
            const uint32_t g = (kj >> 2) > mine;
            gt += g;
            pos_gt += (uint32_t)(kj & 1ull) & g;
            pos_before += (uint32_t)(kj & 1ull) & (uint32_t)(j < i);
          }
          return (uint32_t)a + gt;
        }
        if (ls >= w0 && le <= w0 + kStage) {
          const int o0 = (int)(ls - w0), o1 = (int)(le - w0), oi = kHalo + off;
          int o = o0;
          for (; o + 4 <= o1; o += 4) {   // four independent shared-memory loads in flight (the walk is latency bound)
            const unsigned long long k0 = sh.u.stage[o], k1 = sh.u.stage[o + 1], k2 = sh.u.stage[o + 2], k3 = sh.u.stage[o + 3];
            const uint32_t g0 = (k0 >> 2) > mine, g1 = (k1 >> 2) > mine, g2 = (k2 >> 2) > mine, g3 = (k3 >> 2) > mine;
            gt += g0 + g1 + g2 + g3;
            pos_gt += ((uint32_t)k0 & g0) + ((uint32_t)k1 & g1) + ((uint32_t)k2 & g2) + ((uint32_t)k3 & g3);
            pos_before += ((uint32_t)k0 & (uint32_t)(o < oi)) + ((uint32_t)k1 & (uint32_t)(o + 1 < oi)) +
                          ((uint32_t)k2 & (uint32_t)(o + 2 < oi)) + ((uint32_t)k3 & (uint32_t)(o + 3 < oi));
          }
          for (; o < o1; ++o) {
            const unsigned long long kj = sh.u.stage[o];
            const uint32_t g = (kj >> 2) > mine;
            gt += g;
            pos_gt += (uint32_t)kj & g;                      // bit 1 is clear in an unrefined bin
            pos_before += (uint32_t)kj & (uint32_t)(o < oi);
          }
        } else {   // leaf_max > kHalo (more than 131 k rows): the leaf may leave the window
          for (int64_t j = ls; j < le; ++j) {
            const unsigned long long kj = scratch[j - sb];
            const uint32_t g = (kj >> 2) > mine;
            gt += g;
            pos_gt += (uint32_t)(kj & 1ull) & g;
            pos_before += (uint32_t)(kj & 1ull) & (uint32_t)(j < i);
          }
        }
        return (uint32_t)ls + gt;
      };
      if (p.ties) {
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          const int off = tid + u * kThreads;   // strided: the lanes of a warp walk neighbouring leaves
          if (base + off < se) {
            uint32_t pb, pg;
            const uint32_t rank = resolve(off, pb, pg);
            tmp[rank] = sh.u.stage[kHalo + off] & ~2ull;   // the fully sorted row
          }
        }
      } else {
        uint32_t rel[kItems], tsum = 0;
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          const int off = tid * kItems + u;
          rel[u] = (base + off < se) ? (uint32_t)(sh.u.stage[kHalo + off] & 1ull) : 0u;
          tsum += rel[u];
        }
        unsigned long long total;
        uint32_t pre = (uint32_t)block_excl_add(tsum, sh.red, total);   // relevant positions of the chunk before mine
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          if (rel[u]) sh.plist[pre] = (uint16_t)(tid * kItems + u);
          pre += rel[u];
        }
        __syncthreads();
        for (uint32_t e = tid; e < (uint32_t)total; e += kThreads) {
          uint32_t pb, pg;
          const uint32_t rank = resolve(sh.plist[e], pb, pg);
          __stcs(out_rank + ((uint32_t)carry + e - pb + pg), (int32_t)rank);
        }
        carry += total;
      }
    }
    __syncthreads();
    RP_TICK(5);
    }   // slabs

    if (p.ties) {
      ties_pass(p, tmp, nr, row, sh.red);
    } else if (tid == 0) {
      p.npos[row] = (int32_t)carry;
    }
    if (tid == 0 && p.nranked) p.nranked[row] = (int32_t)nr;
    __syncthreads();
    RP_TICK(6);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// AP variants from the ascending ranks of the positives (one warp per query).
//   ap_trapz / prs : compute_ap / compute_map, test.py:58-146 (self_last = 1: the query itself is one more positive at
//                    rank `nranked`, SURVEY Q2 -- it was left out of the ranking with its -inf score)
//   prec_sum       : sum over the positives of (positives so far) / (1-based rank) -- test.py:974-981,
//                    train.py:424-433, fusion_eval/metrics.py:78-83
//   first          : 1-based rank of the best positive (0 = none);  hits_at[t] : positives with rank < kappas[t]
// IEEE double, the reference's operation order, positives accumulated in rank order.
// ------------------------------------------------------------------------------------------------------------------
__global__ void ap_from_ranks_kernel(const int32_t* __restrict__ pos_ranks, int64_t ld, const int32_t* __restrict__ npos,
                                     const int32_t* __restrict__ nranked, int64_t nq, int self_last,
                                     const int32_t* __restrict__ kappas, int nkappa, double* __restrict__ ap_trapz,
                                     double* __restrict__ prs, int32_t* __restrict__ nres_out,
                                     double* __restrict__ prec_sum, int32_t* __restrict__ first,
                                     int32_t* __restrict__ hits_at) {
  __shared__ double terms[4][2][32];   // per warp: the 32 terms of a step, summed in rank order by every lane
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
  if (q >= nq) return;
  const int32_t* rk = pos_ranks + q * ld;
  const int np = npos[q];
  const int nres = np + (self_last ? 1 : 0);
  const bool want_ap = ap_trapz != nullptr, want_ps = prec_sum != nullptr;
  const double recall_step = nres > 0 ? __ddiv_rn(1.0, (double)nres) : 0.0;
  double ap = 0.0, ps = 0.0;
  int within[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int base = 0; base < nres; base += 32) {
    const int j = base + lane;
    double t_ap = 0.0, t_ps = 0.0;
    if (j < nres) {
      const int64_t rank = j < np ? rk[j] : nranked[q];
      const double p1 = __ddiv_rn((double)(j + 1), (double)(rank + 1));
      if (want_ap) {
        const double p0 = (rank == 0) ? 1.0 : __ddiv_rn((double)j, (double)rank);
        t_ap = __ddiv_rn(__dmul_rn(__dadd_rn(p0, p1), recall_step), 2.0);
      }
      t_ps = j < np ? p1 : 0.0;
      for (int t = 0; t < nkappa && t < 8; ++t) within[t] += (rank < kappas[t]);
    }
    terms[w][0][lane] = t_ap;
    terms[w][1][lane] = t_ps;
    __syncwarp();
    if (base + 32 <= np) {        // a full step of ranked positives: fixed trip count, the loads pipeline
      if (want_ap) {
#pragma unroll
        for (int t = 0; t < 32; ++t) ap = __dadd_rn(ap, terms[w][0][t]);
      }
      if (want_ps) {
#pragma unroll
        for (int t = 0; t < 32; ++t) ps = __dadd_rn(ps, terms[w][1][t]);
      }
    } else {
      const int valid = nres - base < 32 ? nres - base : 32;
      for (int t = 0; t < valid; ++t) {
        if (want_ap) ap = __dadd_rn(ap, terms[w][0][t]);
        if (want_ps && base + t < np) ps = __dadd_rn(ps, terms[w][1][t]);
      }
    }
    __syncwarp();
  }
  for (int t = 0; t < nkappa && t < 8; ++t)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) within[t] += __shfl_xor_sync(0xFFFFFFFFu, within[t], off);
  if (lane != 0) return;
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  if (ap_trapz) ap_trapz[q] = nres > 0 ? ap : nan;
  if (nres_out) nres_out[q] = nres;
  if (prec_sum) prec_sum[q] = ps;
  if (first) first[q] = np > 0 ? rk[0] + 1 : 0;
  const int64_t maxpos = nres == 0 ? 0 : (self_last ? (int64_t)nranked[q] + 1 : (int64_t)rk[np - 1] + 1);
  for (int t = 0; t < nkappa && t < 8; ++t) {
    // positives of the RANKED list below the cut-off (the appended self entry sits at rank nranked)
    int ranked_within = within[t];
    if (self_last && nranked[q] < kappas[t]) ranked_within -= 1;
    if (hits_at) hits_at[q * nkappa + t] = ranked_within;
    if (prs) {
      // kq = min(max(pos), kappa); prs = (pos <= kq).sum() / kq   (test.py:137-140)
      if (nres == 0) { prs[q * nkappa + t] = nan; continue; }
      const int64_t kq = maxpos < kappas[t] ? maxpos : kappas[t];
      const int c = kq < kappas[t] ? nres : within[t];
      prs[q * nkappa + t] = __ddiv_rn((double)c, (double)kq);
    }
  }
}

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src) of an array of length n that is zero except at
// m ascending positions idx[e] (values val[e]); `cur` walks the non-zeros.  Adding +0.0 never changes a partial sum
// (no term is -0.0), so empty ranges are skipped and only the shape of numpy's summation tree is reproduced.
struct SparseTerms {
  const int32_t* idx;   // stored DESCENDING (slot m-1-e holds ascending element e)
  const double* val;
  int m;
  __device__ int index(int e) const { return idx[m - 1 - e]; }
  __device__ double value(int e) const { return val[m - 1 - e]; }
};
// one block of numpy's tree (n <= 128): eight strided accumulators over the whole multiples of 8, then the tail
__device__ double np_block_sparse(const SparseTerms& a, int lo, int n, int& cur) {
  if (n < 8) {
    double res = 0.0;
    while (cur < a.m && a.index(cur) < lo + n) res = __dadd_rn(res, a.value(cur++));
    return res;
  }
  double r[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int body = lo + n - (n % 8);
  while (cur < a.m && a.index(cur) < body) {
    const int t = (a.index(cur) - lo) & 7;
    r[t] = __dadd_rn(r[t], a.value(cur++));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  while (cur < a.m && a.index(cur) < lo + n) res = __dadd_rn(res, a.value(cur++));
  return res;
}
// The recursion `sum(lo, n) = sum(lo, n2) + sum(lo + n2, n - n2)`, n2 = n/2 rounded down to a multiple of 8, walked with
// an explicit stack (depth <= 32 for n < 2^30; device recursion would need a run-time stack size) over the index range
// [lo0, lo0 + n0) of the sparse array; `cur` = first non-zero at or after lo0.
__device__ double np_pairwise_sparse(const SparseTerms& a, int lo0, int n0, int cur) {
  int lo_s[32], n_s[32];
  double left_s[32];
  unsigned char phase_s[32];
  int sp = 0;
  lo_s[0] = lo0; n_s[0] = n0; phase_s[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    const int lo = lo_s[sp], n = n_s[sp];
    if (phase_s[sp] == 0) {
      if (cur >= a.m || a.index(cur) >= lo + n) { ret = 0.0; --sp; continue; }   // nothing but zeros in here
      if (n <= 128) { ret = np_block_sparse(a, lo, n, cur); --sp; continue; }
      int n2 = n / 2;
      n2 -= n2 % 8;
      phase_s[sp] = 1;
      ++sp;
      lo_s[sp] = lo; n_s[sp] = n2; phase_s[sp] = 0;
    } else if (phase_s[sp] == 1) {
      left_s[sp] = ret;
      int n2 = n / 2;
      n2 -= n2 % 8;
      phase_s[sp] = 2;
      ++sp;
      lo_s[sp] = lo + n2; n_s[sp] = n - n2; phase_s[sp] = 0;
    } else {
      ret = __dadd_rn(left_s[sp], ret);
      --sp;
    }
  }
  return ret;
}

// sklearn.metrics.average_precision_score over the FULL ranking from the positives' tie structure
// (train.py:473, nih_multilabel_training.py:95): thresholds = runs of equal scores, AP = -sum(diff(recall) *
// precision[:-1]) on the reversed curves -- only runs holding a positive contribute a non-zero term.
// One warp per query.  The lanes build the non-zero terms (one per run of equal scores that holds a positive).  numpy's
// summation tree over all `ngroups` thresholds is then split at depth 5: lane L follows the bits of L down from the
// root (register arithmetic), sums ITS subtree exactly as numpy would, and the five top levels are combined with
// shuffles -- `left + right` at every node, the same additions in the same pairing.
__global__ void ap_sklearn_ranks_kernel(const int32_t* __restrict__ pos_ge, const int32_t* __restrict__ pos_tg,
                                        int64_t ld, const int32_t* __restrict__ npos,
                                        const int32_t* __restrict__ ngroups, int64_t nq, int32_t* __restrict__ ws_idx,
                                        double* __restrict__ ws_val, double* __restrict__ ap) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  const int np = npos[q];
  if (np == 0) {
    if (lane == 0) ap[q] = __longlong_as_double(0x7FF8000000000000ll);
    return;
  }
  const int32_t* ge = pos_ge + q * ld;
  const int32_t* tg = pos_tg + q * ld;
  int32_t* idx = ws_idx + q * ld;
  double* val = ws_val + q * ld;
  const int T = ngroups[q];
  int m = 0;
  for (int base = 0; base < np; base += 32) {
    const int i = base + lane;
    const bool last = i < np && (i + 1 == np || ge[i + 1] != ge[i]);   // last positive of its run
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, last);
    if (last) {
      int j = i;
      while (j > 0 && ge[j - 1] == ge[i]) --j;                          // positives before the run = recall before it
      const int tp = i + 1, tp_prev = j;
      const double prec = __ddiv_rn((double)tp, (double)ge[i]);        // tps / (tps + fps), tps + fps = ge
      const double rec = __ddiv_rn((double)tp, (double)np);
      const double rec_prev = tp_prev > 0 ? __ddiv_rn((double)tp_prev, (double)np) : 0.0;
      const int slot = m + __popc(bal & ((1u << lane) - 1u));
      idx[slot] = T - 1 - tg[i];
      val[slot] = __dmul_rn(__dadd_rn(rec_prev, -rec), prec);
    }
    m += __popc(bal);
  }
  __syncwarp();
  const SparseTerms a{idx, val, m};
  if (T <= 8192) {   // a tree this small may have leaves above depth 5: one lane walks all of it
    if (lane == 0) ap[q] = -np_pairwise_sparse(a, 0, T, 0);
    return;
  }
  int lo = 0, n = T;   // every node above depth 5 holds more than 128 thresholds: the tree is complete down to there
#pragma unroll
  for (int bit = 4; bit >= 0; --bit) {
    int n2 = n / 2;
    n2 -= n2 % 8;
    if ((lane >> bit) & 1) { lo += n2; n -= n2; }
    else n = n2;
  }
  int lo_e = 0, hi_e = m;   // first non-zero at or after lo (terms ascend in the reversed view)
  while (lo_e < hi_e) {
    const int mid = (lo_e + hi_e) >> 1;
    if (a.index(mid) >= lo) hi_e = mid; else lo_e = mid + 1;
  }
  double v = np_pairwise_sparse(a, lo, n, lo_e);
#pragma unroll
  for (int level = 0; level < 5; ++level) v = __dadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, 1 << level));
  if (lane == 0) ap[q] = -v;
}

int sm_count_rp() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    return 148;
  }
  return sms;
}

int bins_for(int64_t ng) {
  int nb = kThreads;   // a multiple of kThreads: every warp scans whole 32-entry groups
  while (nb < kMaxBins && nb < ng) nb <<= 1;
  return nb;
}

size_t rank_smem_bytes(int nbins) { return (size_t)nbins * sizeof(uint32_t) + sizeof(Shared); }

int64_t rank_grid(int64_t nq) {
  const int64_t cap = (int64_t)sm_count_rp() * kCtasPerSm;   // 1024 threads per SM (shared memory: ~110 KB per CTA)
  return nq < cap ? nq : cap;
}

}  // namespace
}  // namespace knn

using namespace knn;

#ifdef RP_TIMING
extern "C" __attribute__((visibility("default"))) int knn_rank_timing(unsigned long long* out8_host, int reset) {
  cudaDeviceSynchronize();
  if (out8_host) cudaMemcpyFromSymbol(out8_host, knn::rp_timing, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(knn::rp_timing, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" size_t knn_rank_of_positives_workspace(int64_t nq, int64_t ng) {
  if (nq <= 0 || ng <= 0) return 0;
  return (size_t)rank_grid(nq) * (size_t)ng * 2 * sizeof(uint64_t) + 256;   // key rows (two per resident CTA) + the row counter
}

extern "C" int knn_rank_of_positives(const float* scores, int64_t ld_scores, int64_t nq, int64_t ng, int largest_first,
                                     int rel_mode, const void* q_rel, const void* g_rel, double jaccard_thr,
                                     int64_t self_offset, int drop_self, const int64_t* q_group,
                                     const int64_t* g_group, int32_t* pos_ranks, int64_t ld_out,
                                     int32_t* pos_ge, int32_t* pos_tgroup, int32_t* npos, int32_t* nranked,
                                     int32_t* ngroups, void* workspace, size_t workspace_bytes, void* stream) {
  KNN_REQUIRE(nq >= 0 && ng >= 1 && ng < (1ll << 30), "knn_rank_of_positives: bad shape nq=%lld ng=%lld (ng < 2^30)",
              (long long)nq, (long long)ng);
  KNN_REQUIRE(ld_scores >= ng && ld_out >= 1, "knn_rank_of_positives: bad leading dimensions");
  KNN_REQUIRE(rel_mode >= kRelSingle && rel_mode <= kRelAny, "knn_rank_of_positives: bad rel_mode %d", rel_mode);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(scores && q_rel && g_rel && pos_ranks && npos, "knn_rank_of_positives: null pointer");
  KNN_REQUIRE((reinterpret_cast<uintptr_t>(g_rel) & 15) == 0, "knn_rank_of_positives: g_rel must be 16-byte aligned");
  const bool ties = pos_ge != nullptr || pos_tgroup != nullptr || ngroups != nullptr;
  KNN_REQUIRE(!ties || (pos_ge && pos_tgroup && ngroups), "knn_rank_of_positives: tie outputs come together");
  const size_t need = knn_rank_of_positives_workspace(nq, ng);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("knn_rank_of_positives: workspace too small (%zu < %zu)", workspace_bytes, need);
    return KNN_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  RankParams p;
  memset(&p, 0, sizeof(p));
  p.scores = scores; p.ld_scores = ld_scores; p.nq = nq; p.ng = ng; p.largest = largest_first ? 1 : 0;
  p.rel_mode = rel_mode;
  p.qlab = reinterpret_cast<const int64_t*>(q_rel);
  p.glab = reinterpret_cast<const int64_t*>(g_rel);
  p.thr = jaccard_thr;
  p.self_offset = self_offset; p.drop_self = drop_self ? 1 : 0;
  KNN_REQUIRE((q_group == nullptr) == (g_group == nullptr), "knn_rank_of_positives: q_group and g_group come together");
  p.q_group = q_group; p.g_group = g_group;
  p.ties = ties ? 1 : 0;
  p.nbins = bins_for(ng);
  p.leaf_max = 256;
  if ((ng + 511) / 512 > p.leaf_max) p.leaf_max = (int)((ng + 511) / 512);
  const int64_t grid = rank_grid(nq);
  // slabs: the keys being scattered / resolved at any time (grid rows x ng / nslabs x 8 bytes) should stay L2 resident
  {
    static const long long l2_budget = [] {
      const char* e = getenv("KNN_RANK_L2_BYTES");   // experiment knob
      return e ? atoll(e) : (96ll << 20);   // measured best on B200 (126 MB L2): 3 slabs at 296 x 112 k
    }();
    long long ns = ((long long)grid * ng * 8 + l2_budget - 1) / l2_budget;
    p.nslabs = (int)(ns < 1 ? 1 : (ns > 32 ? 32 : ns));
  }
  p.scratch = reinterpret_cast<uint64_t*>(workspace);
  p.tmp = p.scratch + (size_t)grid * (size_t)ng;
  p.next_row = reinterpret_cast<unsigned int*>(p.tmp + (size_t)grid * (size_t)ng);
  KNN_CHECK_CUDA(cudaMemsetAsync(p.next_row, 0, sizeof(unsigned int), st));
  p.pos_ranks = pos_ranks; p.ld_out = ld_out; p.pos_ge = pos_ge; p.pos_tgroup = pos_tgroup;
  p.npos = npos; p.nranked = nranked; p.ngroups = ngroups;
  const size_t smem = rank_smem_bytes(p.nbins);
  KNN_CHECK_CUDA(cudaFuncSetAttribute(rank_positives_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rank_positives_kernel<<<(unsigned)grid, kThreads, smem, st>>>(p);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_ap_from_ranks(const int32_t* pos_ranks, int64_t ld, const int32_t* npos, const int32_t* nranked,
                                 int64_t nq, int self_last_positive, const int32_t* kappas, int nkappa,
                                 double* ap_trapz, double* prs, int32_t* nres, double* prec_sum, int32_t* first,
                                 int32_t* hits_at, void* stream) {
  KNN_REQUIRE(nq >= 0 && ld >= 1 && nkappa >= 0 && nkappa <= 8, "knn_ap_from_ranks: bad sizes (nkappa <= 8)");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(pos_ranks && npos && nranked && (kappas || nkappa == 0), "knn_ap_from_ranks: null pointer");
  ap_from_ranks_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
      pos_ranks, ld, npos, nranked, nq, self_last_positive ? 1 : 0, kappas, nkappa, ap_trapz, prs, nres, prec_sum,
      first, hits_at);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" size_t knn_ap_sklearn_from_ranks_workspace(int64_t nq, int64_t ld) {
  return nq > 0 && ld > 0 ? (size_t)nq * (size_t)ld * (sizeof(double) + sizeof(int32_t)) : 0;
}

extern "C" int knn_ap_sklearn_from_ranks(const int32_t* pos_ge, const int32_t* pos_tgroup, int64_t ld,
                                         const int32_t* npos, const int32_t* ngroups, int64_t nq, double* ap,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  KNN_REQUIRE(nq >= 0 && ld >= 1, "knn_ap_sklearn_from_ranks: bad sizes");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(pos_ge && pos_tgroup && npos && ngroups && ap, "knn_ap_sklearn_from_ranks: null pointer");
  if (!workspace || workspace_bytes < knn_ap_sklearn_from_ranks_workspace(nq, ld)) {
    set_error("knn_ap_sklearn_from_ranks: workspace too small");
    return KNN_E_WORKSPACE;
  }
  double* ws_val = reinterpret_cast<double*>(workspace);
  int32_t* ws_idx = reinterpret_cast<int32_t*>(ws_val + (size_t)nq * (size_t)ld);
  ap_sklearn_ranks_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, (cudaStream_t)stream>>>(pos_ge, pos_tgroup, ld, npos,
                                                                                      ngroups, nq, ws_idx, ws_val, ap);
  KNN_LAUNCHED();
  return KNN_OK;
}
