// Final selection over per-unit candidate lists, k-way merge of per-shard results, and full row ranking.
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

namespace knn {

namespace {

constexpr int kSortCap = 4096;      // keys sorted per shared-memory block
constexpr int kSortThreads = 512;

// In-shared-memory bitonic sort, descending overall; `gbase` is the global position of s[0] so the
// same routine serves as the local stage of a larger network (direction depends on global position).
__device__ __forceinline__ void smem_bitonic(uint64_t* s, int n, int size_from, int size_to, int stride_cap,
                                             int64_t gbase) {
  for (int size = size_from; size <= size_to; size <<= 1) {
    int stride = size >> 1;
    if (stride > stride_cap) stride = stride_cap;
    for (; stride > 0; stride >>= 1) {
      const int lowmask = stride - 1;  // strides are powers of two: no integer division in the inner loop
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = ((i & ~lowmask) << 1) | (i & lowmask);
        const bool desc = (((gbase + pos) & (int64_t)size) == 0);
        uint64_t a = s[pos], b = s[pos + stride];
        const bool swap = desc ? (a < b) : (a > b);
        if (swap) { s[pos] = b; s[pos + stride] = a; }
      }
      __syncthreads();
    }
  }
}

__host__ __device__ __forceinline__ int pow2_ge(int x) {
  int p = 2;
  while (p < x) p <<= 1;
  return p;
}

// One CTA per query row: exact top-k of the union of the row's candidate lists (one UNORDERED list of counts[..]
// keys per gallery split and selection group), written best-first.
//   1. keys below the shared threshold tau_global[row] (a lower bound of the final k-th best) cannot be in the answer;
//   2. if more than kSortCap keys remain (many splits), an MSB radix select over the 64-bit keys (8 bits per pass,
//      256-bin shared-memory histogram) raises the bound until at most kSortCap keys are >= it -- the k-th best
//      key always stays above the bound, whatever the score distribution or number of ties;
//   3. the survivors are gathered into shared memory, sorted by a bitonic network and the first k are emitted.
// Seeding mode (tau_out != nullptr): nothing is emitted; the k-th best SCORE of the sample is published as the
// row's starting threshold for the main pass.
__global__ void __launch_bounds__(kSortThreads) merge_units_kernel(SearchParams p, int64_t index_base,
                                                                 float* __restrict__ out_val,
                                                                 int64_t* __restrict__ out_idx,
                                                                 uint32_t* __restrict__ tau_out, int cap) {
  // `cap` (power of two <= kSortCap) = shared-memory sort capacity of this launch: min(kSortCap, keys per row)
  extern __shared__ __align__(16) uint8_t merge_smem[];
  uint64_t* s = reinterpret_cast<uint64_t*>(merge_smem);
  __shared__ int hist[256];
  __shared__ int n_surv;
  __shared__ unsigned long long sh_lo;
  __shared__ int sh_done, sh_kk;
  const int64_t r = blockIdx.x;
  const int qb = (int)(r / kRowsPerUnit), lr = (int)(r % kRowsPerUnit);
  const int L = 2 * p.kp;
  const int V = p.splits * p.groups;  // (split, group) pairs are laid out as virtual splits
  const bool l2 = p.metric == KNN_L2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // visit every key of this row: warp w walks lists w, w + nwarps, ...; lanes stride through a list (coalesced).
  // fn(key, valid) is called by ALL lanes of the warp in every iteration (warp-aggregated atomics inside).
  // two-phase form (small query batches): a list-parallel pre-filter already gathered this row's keys >= tau_global
  const int pre_n = (p.pre != nullptr && tau_out == nullptr) ? __ldcg(p.precount + r) : -1;
  const bool use_pre = pre_n >= 0 && pre_n <= p.pre_cap;   // overflow (mass ties): read the lists after all
  auto for_each_key = [&](auto&& fn) {
    if (use_pre) {
      const uint64_t* base = p.pre + r * (int64_t)p.pre_cap;
      for (int j0 = warp * 32; j0 < pre_n; j0 += nwarps * 32) {
        const bool valid = j0 + lane < pre_n;
        fn(valid ? __ldcg(base + j0 + lane) : 0ull, valid);
      }
      return;
    }
    for (int v = warp; v < V; v += nwarps) {
      const int64_t unit_row = ((int64_t)v * p.qblocks + qb) * kRowsPerUnit + lr;
      int cnt = __ldcg(p.counts + unit_row);
      cnt = cnt < 0 ? 0 : (cnt > L ? L : cnt);
      const uint64_t* base = p.lists + unit_row * (int64_t)L;
      for (int j0 = 0; j0 < cnt; j0 += 32) {
        const bool valid = j0 + lane < cnt;
        fn(valid ? __ldcg(base + j0 + lane) : 0ull, valid);
      }
    }
  };

  unsigned long long lo = (unsigned long long)__ldcg(p.tau_global + r) << 32;
  if (threadIdx.x == 0) n_surv = 0;
  __syncthreads();
  // gather the keys >= lo into shared memory, counting them all (slots beyond `cap` are not written)
  auto gather = [&]() {
    for_each_key([&](uint64_t key, bool valid) {
      const bool keep = valid && key >= lo;
      const unsigned b = __ballot_sync(0xFFFFFFFFu, keep);
      int base = 0;
      if (lane == 0 && b) base = atomicAdd(&n_surv, __popc(b));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      const int slot = base + __popc(b & ((1u << lane) - 1u));
      if (keep && slot < cap) s[slot] = key;
    });
  };
  // Few keys in total (they all fit the sort buffer, or the pre-filter already gathered them): ONE pass.  Otherwise
  // count first -- with many long lists (8192 x 50 M: 74 lists of ~150 keys) the count usually exceeds `cap`, and a
  // wasted gather costs more than a counting pass.
  const bool count_first = !use_pre && (int64_t)V * L > cap;
  if (count_first) {
    int c = 0;
    for_each_key([&](uint64_t key, bool valid) { c += (valid && key >= lo) ? 1 : 0; });
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane == 0 && c) atomicAdd(&n_surv, c);
  } else {
    gather();
  }
  __syncthreads();
  const int total = n_surv;
  __syncthreads();
  if (count_first && total <= cap) {
    if (threadIdx.x == 0) n_surv = 0;
    __syncthreads();
    gather();
    __syncthreads();
  }
  if (total > cap) {
    // MSB radix select of the k-th largest key among the keys >= lo.  sh_kk = rank still to find inside the
    // current bin; k - sh_kk = keys known to lie strictly above it (all of them are in the answer).
    if (threadIdx.x == 0) {
      sh_lo = 0ull;
      sh_kk = p.k;
      sh_done = 0;
    }
    __syncthreads();
    for (int d = 56; d >= 0; d -= 8) {
      const unsigned long long prefix = sh_lo;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for_each_key([&](uint64_t key, bool valid) {
        if (valid && key >= lo && (d == 56 || (key >> (d + 8)) == (prefix >> (d + 8))))
          atomicAdd(&hist[(int)((key >> d) & 255ull)], 1);
      });
      __syncthreads();
      if (threadIdx.x == 0) {
        const int kk = sh_kk;
        int acc = 0, b = 255;
        for (; b > 0; --b) {
          if (acc + hist[b] >= kk) break;
          acc += hist[b];
        }
        sh_kk = kk - acc;
        sh_lo = prefix | ((unsigned long long)b << d);
        const int cnt_ge = (p.k - sh_kk) + hist[b];  // keys >= the new bound
        sh_done = (cnt_ge <= cap || d == 0) ? 1 : 0;
      }
      __syncthreads();
      if (sh_done) break;
    }
    if (sh_lo > lo) lo = sh_lo;
    if (threadIdx.x == 0) n_surv = 0;
    __syncthreads();
    gather();   // again, with the raised bound
    __syncthreads();
  }
  int ns = n_surv < cap ? n_surv : cap;
  __syncthreads();
  if (ns > 4 * p.kp) {
    // Still far more survivors than k: the bitonic network below costs O(n log^2 n) shared-memory passes, so first
    // cut the set down with the same MSB radix select, now over the keys held in shared memory (a pass is ns/threads
    // shared-memory atomics per thread), stopping as soon as at most 2*KP keys are >= the bound.
    if (threadIdx.x == 0) {
      sh_lo = 0ull;
      sh_kk = p.k;
      sh_done = 0;
    }
    __syncthreads();
    for (int d = 56; d >= 0; d -= 8) {
      const unsigned long long prefix = sh_lo;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        const uint64_t key = s[i];
        if (d == 56 || (key >> (d + 8)) == (prefix >> (d + 8))) atomicAdd(&hist[(int)((key >> d) & 255ull)], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        const int kk = sh_kk;
        int acc = 0, b = 255;
        for (; b > 0; --b) {
          if (acc + hist[b] >= kk) break;
          acc += hist[b];
        }
        sh_kk = kk - acc;
        sh_lo = prefix | ((unsigned long long)b << d);
        const int cnt_ge = (p.k - sh_kk) + hist[b];
        sh_done = (cnt_ge <= 2 * p.kp || d == 0) ? 1 : 0;
      }
      __syncthreads();
      if (sh_done) break;
    }
    // compact the keys >= bound to the front: every thread first pulls its keys into registers (cap / threads = 8)
    const unsigned long long bound = sh_lo;
    uint64_t mine[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = threadIdx.x + e * blockDim.x;
      mine[e] = i < ns ? s[i] : 0ull;
    }
    if (threadIdx.x == 0) n_surv = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const bool keep = mine[e] != 0ull && mine[e] >= bound;
      const unsigned b = __ballot_sync(0xFFFFFFFFu, keep);
      int base = 0;
      if (lane == 0 && b) base = atomicAdd(&n_surv, __popc(b));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (keep) s[base + __popc(b & ((1u << lane) - 1u))] = mine[e];
    }
    __syncthreads();
    ns = n_surv;
    __syncthreads();
  }
  const int n = pow2_ge(ns > p.k ? ns : p.k);
  for (int i = ns + threadIdx.x; i < n; i += blockDim.x) s[i] = 0ull;
  __syncthreads();
  smem_bitonic(s, n, 2, n, n, 0);

  if (tau_out != nullptr) {
    if (threadIdx.x == 0 && ns >= p.k) atomicMax(tau_out + r, (uint32_t)(s[p.k - 1] >> 32));
    return;
  }
  for (int j = threadIdx.x; j < p.k; j += blockDim.x) {
    const uint64_t key = s[j];
    float v; int64_t id;
    if (key == 0ull) {
      v = l2 ? INFINITY : -INFINITY;
      id = -1;
    } else {
      const float sc = key_score(key);
      v = l2 ? (0.0f - sc) : sc;
      id = (int64_t)key_row(key) + index_base;
    }
    out_val[r * p.k + j] = v;
    out_idx[r * p.k + j] = id;
  }
}

// ------------------------------------------------------------------------------------------------
// Two-phase unit merge for SMALL query batches.  One CTA per row leaves a 148-SM GPU idle when there are 64 rows, and
// each CTA then walks hundreds of lists through dependent loads (count -> keys): 134-180 us per merge at 64 queries x
// 592 lists.  Phase 1 is parallel over (list, row): one warp per list.
//   * final merge: keys >= tau_global[row] are appended to the row's compact buffer (warp-aggregated atomics); the
//     CTA-per-row kernel then sorts that buffer alone.
//   * threshold seeding: only each list's MAXIMUM is kept.  With V >= 2k lists per row the k-th largest of the V
//     maxima (scores of k distinct gallery rows) is a valid threshold, and nearly as tight as the exact k-th best of
//     the sample (592 lists, k = 100: ~110 sample rows lie above it instead of 100).
// ------------------------------------------------------------------------------------------------
template <bool kSeed>
__global__ void __launch_bounds__(256) filter_lists_kernel(SearchParams p) {
  const int lane = threadIdx.x & 31;
  const int V = p.splits * p.groups;
  const int64_t rows = (int64_t)p.qblocks * kRowsPerUnit;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // = v * rows + row
  if (w >= (int64_t)V * rows) return;
  const int v = (int)(w / rows);
  const int64_t r = w % rows;
  if (r >= p.nq) return;
  const int L = 2 * p.kp;
  const int64_t unit_row = (int64_t)v * rows + r;
  int cnt = __ldcg(p.counts + unit_row);
  cnt = cnt < 0 ? 0 : (cnt > L ? L : cnt);
  const uint64_t* base = p.lists + unit_row * (int64_t)L;
  if (kSeed) {
    uint32_t m = 0u;
    for (int j = lane; j < cnt; j += 32) m = max(m, (uint32_t)(__ldcg(base + j) >> 32));
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if (lane == 0) p.maxima[r * V + v] = m;
  } else {
    const unsigned long long lo = (unsigned long long)__ldcg(p.tau_global + r) << 32;
    uint64_t* dst = p.pre + r * (int64_t)p.pre_cap;
    for (int j0 = 0; j0 < cnt; j0 += 32) {
      const uint64_t key = j0 + lane < cnt ? __ldcg(base + j0 + lane) : 0ull;
      const bool keep = key != 0ull && key >= lo;
      const unsigned b = __ballot_sync(0xFFFFFFFFu, keep);
      if (b == 0u) continue;
      int at = 0;
      if (lane == 0) at = atomicAdd(p.precount + r, __popc(b));
      at = __shfl_sync(0xFFFFFFFFu, at, 0) + __popc(b & ((1u << lane) - 1u));
      if (keep && at < p.pre_cap) __stcg(dst + at, key);
    }
  }
}

// One CTA per row: k-th largest of the row's V list maxima (V <= 4096) -> starting threshold.
__global__ void __launch_bounds__(256) seed_select_kernel(const uint32_t* __restrict__ maxima, int V, int64_t nq, int k,
                                                          uint32_t* __restrict__ tau_out) {
  extern __shared__ __align__(16) uint8_t seed_smem[];
  uint32_t* s = reinterpret_cast<uint32_t*>(seed_smem);
  const int64_t r = blockIdx.x;
  const int n = pow2_ge(V);
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = i < V ? __ldcg(maxima + r * V + i) : 0u;
  __syncthreads();
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int lowmask = stride - 1;
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = ((i & ~lowmask) << 1) | (i & lowmask);
        const bool desc = (pos & size) == 0;
        const uint32_t a = s[pos], b = s[pos + stride];
        if (desc ? (a < b) : (a > b)) { s[pos] = b; s[pos + stride] = a; }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0 && s[k - 1] != 0u) atomicMax(tau_out + r, s[k - 1]);
}

// ------------------------------------------------------------------------------------------------
// k-way merge of `parts` sorted candidate lists per query by rank counting (no sort): the output slot
// of candidate (part, j) is j + sum over the other parts of the number of their entries that precede
// it in the total order (score best-first, then ascending gallery index, then part).
// ------------------------------------------------------------------------------------------------
struct Cand { uint32_t o; int64_t id; };

__device__ __forceinline__ bool precedes(const Cand& x, int px, const Cand& a, int pa) {
  if (x.o != a.o) return x.o > a.o;
  if (x.id != a.id) return x.id < a.id;
  return px < pa;
}

// Part p's candidates live at idx[p] / val[p] ([nq, k] each): slices of one dense or all-gathered buffer, or -- the
// fused exchange of the row-sharded search -- the symmetric-memory buffers of the PEER GPUs, read straight over
// NVLink by this kernel (no all-gather, no staging copy: the transfer overlaps the merge of other queries).
constexpr int kMaxMergeParts = 16;
struct MergeParts {
  const int64_t* idx[kMaxMergeParts];
  const float* val[kMaxMergeParts];
};
struct PeerFlags {
  int32_t* flags[kMaxMergeParts];   // flags[p] = rank p's flag array (int32 [parts]) as mapped into this process
};

// Peer exchange without a separate barrier kernel: every rank publishes the number of the current search (`epoch`) into
// slot [rank] of EVERY rank's flag array once its candidates are written (release, system scope); the merge kernel's CTAs
// wait until all slots of THEIR rank's array have reached the epoch (acquire) before they read the peers' lists.
__global__ void peer_signal_kernel(PeerFlags pf, int parts, int rank, int epoch) {
  if ((int)threadIdx.x < parts) {
    __threadfence_system();   // the candidate lists (written by the kernels before this one) before the flag
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(pf.flags[threadIdx.x] + rank), "r"(epoch) : "memory");
  }
}

// Pipelined exchange: ONE warp waits for the peers' flags on the side stream, the merge kernel is launched behind it.
// (8192 merge CTAs spinning would hold every SM while a peer is late and keep the next local search out; one 32-thread
// CTA costs nothing.)  Bounded: a peer that never publishes surfaces as a launch failure, not as a hung GPU.
__global__ void peer_wait_kernel(const int32_t* __restrict__ my_flags, int parts, int epoch) {
  if ((int)threadIdx.x < parts) {
    const long long t0 = clock64();
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(my_flags + threadIdx.x) : "memory");
      if (v < epoch) {
        __nanosleep(256);
        if (clock64() - t0 > 60000000000ll) {   // ~30 s
          printf("b200knn: peer %d never published search %d\n", (int)threadIdx.x, epoch);
          __trap();
        }
      }
    } while (v < epoch);
  }
}

__global__ void __launch_bounds__(256) merge_topk_kernel(MergeParts mp, int parts, int64_t nq, int k, int l2,
                                                        float* __restrict__ out_val,
                                                        int64_t* __restrict__ out_idx,
                                                        const int32_t* __restrict__ my_flags, int epoch) {
  extern __shared__ __align__(16) uint8_t sm[];
  if (my_flags != nullptr) {
    if ((int)threadIdx.x < parts) {
      int v;
      do {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(my_flags + threadIdx.x) : "memory");
        if (v < epoch) __nanosleep(64);
      } while (v < epoch);
    }
    __syncthreads();
  }
  int64_t* sid = reinterpret_cast<int64_t*>(sm);                 // [parts*k]
  uint32_t* so = reinterpret_cast<uint32_t*>(sid + parts * k);   // [parts*k]
  float* sv = reinterpret_cast<float*>(so + parts * k);          // [parts*k]
  const int64_t r = blockIdx.x;
  const int n = parts * k;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const int pp = c / k, j = c % k;
    const int64_t id = __ldcg(mp.idx[pp] + r * k + j);           // .cg: peer / freshly gathered data, never via L1
    const float v = __ldcg(mp.val[pp] + r * k + j);
    sv[c] = v;
    if (id < 0) { so[c] = 0u; sid[c] = INT64_MAX; }
    else { so[c] = f2ord((l2 ? -v : v) + 0.0f); sid[c] = id; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const int pa = c / k, j = c % k;
    const Cand a{so[c], sid[c]};
    int rank = j;
    for (int px = 0; px < parts; ++px) {
      if (px == pa) continue;
      int lo = 0, hi = k;  // number of entries of list px preceding a
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const Cand x{so[px * k + mid], sid[px * k + mid]};
        if (precedes(x, px, a, pa)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      out_val[r * k + rank] = sv[c];
      out_idx[r * k + rank] = sid[c] == INT64_MAX ? -1 : sid[c];
    }
  }
}

static int launch_merge_parts(const MergeParts& mp, int parts, int64_t nq, int k, int metric, float* out_val,
                              int64_t* out_idx, cudaStream_t stream, const PeerFlags* pf = nullptr, int rank = 0,
                              int epoch = 0, bool publish = true) {
  const size_t smem = (size_t)parts * k * 16;
  KNN_REQUIRE(smem <= 200 * 1024, "knn_merge_topk: parts*k=%d too large (max %d)", parts * k, 200 * 1024 / 16);
  if (pf != nullptr && publish) {   // published even for an empty batch: the peers wait for it
    peer_signal_kernel<<<1, 32, 0, stream>>>(*pf, parts, rank, epoch);
    KNN_LAUNCHED();
  }
  if (pf != nullptr && !publish) {   // pipelined exchange: one warp waits, the merge CTAs find the flags set
    peer_wait_kernel<<<1, 32, 0, stream>>>(pf->flags[rank], parts, epoch);
    KNN_LAUNCHED();
  }
  if (nq == 0) return KNN_OK;
  KNN_CHECK_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_topk_kernel<<<(unsigned)nq, 256, smem, stream>>>(mp, parts, nq, k, metric == KNN_L2 ? 1 : 0, out_val, out_idx,
                                                         pf ? pf->flags[rank] : nullptr, epoch);
  KNN_LAUNCHED();
  return KNN_OK;
}

// ------------------------------------------------------------------------------------------------
// Full ranking of every row: one CTA per row, bitonic network over Npad = pow2 >= ng keys, the
// sub-networks of span <= 4096 run in shared memory, the wider strides in the global workspace.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) rank_rows_kernel(const float* __restrict__ scores, int64_t nq,
                                                               int64_t ng, int largest_first, int64_t npad,
                                                               uint64_t* __restrict__ ws,
                                                               int64_t* __restrict__ ranks) {
  __shared__ uint64_t s[kSortCap];
  const int64_t r = blockIdx.x;
  const float* srow = scores + r * ng;
  uint64_t* wrow = ws + r * npad;
  const int64_t nblocks = (npad + kSortCap - 1) / kSortCap;
  const int bn = (int)(npad < kSortCap ? npad : kSortCap);

  // phase A: sort every block of <= 4096 keys in shared memory (directions follow the global network)
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t g0 = b * kSortCap;
    for (int i = threadIdx.x; i < bn; i += blockDim.x) {
      const int64_t c = g0 + i;
      uint64_t key = 0ull;
      if (c < ng) {
        float v = srow[c];
        v = largest_first ? v : -v;
        key = ((uint64_t)f2ord(v + 0.0f) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)c);
      }
      s[i] = key;
    }
    __syncthreads();
    smem_bitonic(s, bn, 2, bn, bn, g0);
    if (nblocks == 1) {
      for (int i = threadIdx.x; i < ng; i += blockDim.x) ranks[r * ng + i] = (int64_t)key_row(s[i]);
      return;
    }
    for (int i = threadIdx.x; i < bn; i += blockDim.x) wrow[g0 + i] = s[i];
    __syncthreads();
  }
  // phase B: wider sub-networks
  for (int64_t size = 2 * (int64_t)kSortCap; size <= npad; size <<= 1) {
    for (int64_t stride = size >> 1; stride >= kSortCap; stride >>= 1) {
      for (int64_t i = threadIdx.x; i < (npad >> 1); i += blockDim.x) {
        const int64_t pos = ((i / stride) * (stride << 1)) + (i % stride);
        const bool desc = ((pos & size) == 0);
        uint64_t a = wrow[pos], b = wrow[pos + stride];
        const bool swap = desc ? (a < b) : (a > b);
        if (swap) { wrow[pos] = b; wrow[pos + stride] = a; }
      }
      __syncthreads();
    }
    for (int64_t b = 0; b < nblocks; ++b) {
      const int64_t g0 = b * kSortCap;
      for (int i = threadIdx.x; i < kSortCap; i += blockDim.x) s[i] = wrow[g0 + i];
      __syncthreads();
      // remaining strides (< 4096) of this `size`: direction bit comes from the global position
      int stride = kSortCap >> 1;
      for (; stride > 0; stride >>= 1) {
        for (int i = threadIdx.x; i < (kSortCap >> 1); i += blockDim.x) {
          const int pos = ((i / stride) * (stride << 1)) + (i % stride);
          const bool desc = (((g0 + pos) & size) == 0);
          uint64_t a = s[pos], bb = s[pos + stride];
          const bool swap = desc ? (a < bb) : (a > bb);
          if (swap) { s[pos] = bb; s[pos + stride] = a; }
        }
        __syncthreads();
      }
      for (int i = threadIdx.x; i < kSortCap; i += blockDim.x) wrow[g0 + i] = s[i];
      __syncthreads();
    }
  }
  for (int64_t i = threadIdx.x; i < ng; i += blockDim.x) ranks[r * ng + i] = (int64_t)key_row(wrow[i]);
}

}  // namespace

__global__ void stats_reduce_kernel(const double* __restrict__ partials, int splits, int qblocks, int64_t nq,
                                    double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nq) return;
  double sum = 0.0, sq = 0.0, mn = INFINITY, mx = -INFINITY;
  for (int sp = 0; sp < splits; ++sp) {
    const double* pp = partials + ((int64_t)sp * qblocks * kRowsPerUnit + r) * 4;
    sum += pp[0];
    sq += pp[1];
    mn = fmin(mn, pp[2]);
    mx = fmax(mx, pp[3]);
  }
  out[r * 4 + 0] = sum; out[r * 4 + 1] = sq; out[r * 4 + 2] = mn; out[r * 4 + 3] = mx;
}

__global__ void rescore_topk_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx, int64_t nq, int k,
                                    const float* __restrict__ table, int64_t table_rows, int table_cols,
                                    const int64_t* __restrict__ qcol, float alpha, float beta, int first_m,
                                    int64_t self_offset, int mask_self, float* __restrict__ out_vals) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * k) return;
  const int64_t q = t / k;
  const int j = (int)(t % k);
  float v = vals[t];
  const int64_t g = idx[t];
  const int64_t c = qcol[q];
  const bool is_self = self_offset >= 0 && g == self_offset + q;
  if (j < first_m && g >= 0 && g < table_rows && !is_self && c >= 0 && c < table_cols)
    v = __fadd_rn(__fmul_rn(alpha, v), __fmul_rn(beta, __ldg(table + g * table_cols + c)));
  if (is_self && mask_self) v = -INFINITY;  // dists.fill_diagonal_(-inf) after the re-scoring (test.py:623)
  out_vals[t] = v;
}

// One CTA per row: order k (value, index) candidates best-first, ties by ascending index (64-bit keys as everywhere).
__global__ void sort_topk_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx, int64_t nq, int k,
                                 int largest, float* __restrict__ out_vals, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) uint8_t sort_smem[];
  uint64_t* s = reinterpret_cast<uint64_t*>(sort_smem);
  const int64_t r = blockIdx.x;
  const int n = pow2_ge(k);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    uint64_t key = 0ull;
    if (j < k) {
      const int64_t g = idx[r * k + j];
      const float v = vals[r * k + j];
      if (g >= 0) key = make_key(largest ? v : (0.0f - v), (uint32_t)g);
    }
    s[j] = key;
  }
  __syncthreads();
  smem_bitonic(s, n, 2, n, n, 0);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = s[j];
    float v; int64_t id;
    if (key == 0ull) {
      v = largest ? -INFINITY : INFINITY;
      id = -1;
    } else {
      const float sc = key_score(key);
      v = largest ? sc : (0.0f - sc);
      id = (int64_t)key_row(key);
    }
    out_vals[r * k + j] = v;
    out_idx[r * k + j] = id;
  }
}

int launch_sort_topk(const float* vals, const int64_t* idx, int64_t nq, int k, int largest, float* out_vals,
                     int64_t* out_idx, cudaStream_t stream) {
  if (nq == 0) return KNN_OK;
  const int n = pow2_ge(k);
  const int threads = n >= 512 ? 256 : (n >= 128 ? 64 : 32);
  sort_topk_kernel<<<(unsigned)nq, threads, (size_t)n * sizeof(uint64_t), stream>>>(vals, idx, nq, k, largest, out_vals,
                                                                                  out_idx);
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_stats_reduce(const double* partials, int splits, int qblocks, int64_t nq, double* out, cudaStream_t stream) {
  if (nq == 0) return KNN_OK;
  stats_reduce_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(partials, splits, qblocks, nq, out);
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_rescore_topk(const float* vals, const int64_t* idx, int64_t nq, int k, const float* table, int64_t table_rows,
                        int table_cols, const int64_t* qcol, float alpha, float beta, int first_m, int64_t self_offset,
                        int mask_self, float* out_vals, cudaStream_t stream) {
  const int64_t n = nq * k;
  if (n == 0) return KNN_OK;
  rescore_topk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(vals, idx, nq, k, table, table_rows, table_cols,
                                                                      qcol, alpha, beta, first_m, self_offset, mask_self,
                                                                      out_vals);
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_seed_from_maxima(const SearchParams& p, uint32_t* tau_out, cudaStream_t stream, bool reduced) {
  if (p.nq == 0) return KNN_OK;
  const int V = p.splits * p.groups;
  if (p.maxima == nullptr || V < p.k || V > 4096) {
    set_error("internal: threshold seeding from list maxima needs k <= lists <= 4096 and scratch (lists=%d k=%d)", V, p.k);
    return KNN_E_INVALID;
  }
  if (!reduced) {
    const int64_t warps = (int64_t)V * p.qblocks * kRowsPerUnit;
    filter_lists_kernel<true><<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(p);
    KNN_LAUNCHED();
  }
  seed_select_kernel<<<(unsigned)p.nq, 256, (size_t)pow2_ge(V) * sizeof(uint32_t), stream>>>(p.maxima, V, p.nq, p.k, tau_out);
  KNN_LAUNCHED();
  return KNN_OK;
}

int launch_merge_units(const SearchParams& p, int64_t index_base, float* out_val, int64_t* out_idx,
                       uint32_t* tau_out, cudaStream_t stream) {
  if (p.nq == 0) return KNN_OK;
  if (p.pre != nullptr && tau_out == nullptr && p.splits > 0) {  // two-phase form: list-parallel pre-filter (precount zeroed by the caller)
    const int64_t warps = (int64_t)p.splits * p.groups * p.qblocks * kRowsPerUnit;
    // With >= 2k lists per row, first tighten the shared threshold to the k-th largest LIST MAXIMUM: the running
    // threshold only ever reflects one list's (or the seeding sample's) k-th best, which leaves ~k * rows / sample keys
    // per list above it (15 k keys per row at 64 x 10 M); above the k-th largest maximum there are ~1.1 k in total.
    const int V = p.splits * p.groups;
    if (p.maxima != nullptr && V >= 2 * p.k && V <= 4096) {
      int rc = launch_seed_from_maxima(p, p.tau_global, stream);
      if (rc != KNN_OK) return rc;
    }
    filter_lists_kernel<false><<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(p);
    KNN_LAUNCHED();
  }
  // size the CTA and its sort buffer to the input: few short lists per row (small galleries, many queries) get
  // small CTAs so that many rows are merged per SM at once
  const int64_t max_keys = (int64_t)p.splits * p.groups * 2 * p.kp;
  int cap = 2 * p.kp;  // >= k, power of two
  while (cap < max_keys && cap < kSortCap) cap <<= 1;
  // Many rows, final merge: the shared threshold every unit published leaves k .. ~2k keys per row above it, whatever
  // the number of lists, so a row rarely needs more than 4 * kp sort slots -- and a 128-thread CTA with an 8 KB buffer
  // (12 per SM) spends far less time in block barriers than a 512-thread one with 32 KB (3 per SM; ncu: 16.5 barrier
  // stalls per issue).  Rows with more survivors (mass ties at the cut-off) take the radix-select passes of the kernel.
  static const int optimistic = [] {
    const char* e = getenv("KNN_MERGE_SMALL_CTA");   // 0: size the CTA for the worst case (round-1 behaviour)
    return e ? atoi(e) : 1;
  }();
  // (up to 32 lists per row -- 8192 x 6.25 M, 18 lists: 0.57 -> 0.31 ms; with the 74 lists of 8192 x 50 M more rows
  // overflow the small buffer than the shorter barriers save: 1.19 -> 1.30 ms, so those keep the large CTA)
  if (optimistic && tau_out == nullptr && p.nq >= 2048 && p.splits * p.groups <= 32) {
    const int small = 4 * p.kp > 1024 ? 4 * p.kp : 1024;
    if (cap > small) cap = small;
  }
  const int threads = cap <= 1024 ? 128 : (cap <= 2048 ? 256 : kSortThreads);
  merge_units_kernel<<<(unsigned)p.nq, threads, (size_t)cap * sizeof(uint64_t), stream>>>(p, index_base, out_val,
                                                                                        out_idx, tau_out, cap);
  KNN_LAUNCHED();
  return KNN_OK;
}

}  // namespace knn

using namespace knn;

extern "C" int knn_merge_topk(const float* vals, const int64_t* idx, int parts, int64_t nq, int k, int metric,
                              float* out_val, int64_t* out_idx, void* stream) {
  KNN_REQUIRE(vals && idx && out_val && out_idx, "knn_merge_topk: null pointer");
  KNN_REQUIRE(parts >= 1 && parts <= kMaxMergeParts && k >= 1 && nq >= 0,
              "knn_merge_topk: bad sizes parts=%d (max %d) k=%d nq=%lld", parts, kMaxMergeParts, k, (long long)nq);
  MergeParts mp;
  for (int p = 0; p < parts; ++p) {
    mp.idx[p] = idx + (int64_t)p * nq * k;
    mp.val[p] = vals + (int64_t)p * nq * k;
  }
  return launch_merge_parts(mp, parts, nq, k, metric, out_val, out_idx, (cudaStream_t)stream);
}

extern "C" int knn_merge_topk_parts(const float* const* val_parts_host, const int64_t* const* idx_parts_host, int parts,
                                    int64_t nq, int k, int metric, float* out_val, int64_t* out_idx, void* stream) {
  KNN_REQUIRE(val_parts_host && idx_parts_host && out_val && out_idx, "knn_merge_topk_parts: null pointer");
  KNN_REQUIRE(parts >= 1 && parts <= kMaxMergeParts && k >= 1 && nq >= 0,
              "knn_merge_topk_parts: bad sizes parts=%d (max %d) k=%d nq=%lld", parts, kMaxMergeParts, k, (long long)nq);
  MergeParts mp;
  for (int p = 0; p < parts; ++p) {
    KNN_REQUIRE(val_parts_host[p] && idx_parts_host[p], "knn_merge_topk_parts: null part %d", p);
    mp.idx[p] = idx_parts_host[p];
    mp.val[p] = val_parts_host[p];
  }
  return launch_merge_parts(mp, parts, nq, k, metric, out_val, out_idx, (cudaStream_t)stream);
}

extern "C" int knn_merge_topk_parts_sync(const float* const* val_parts_host, const int64_t* const* idx_parts_host,
                                         int parts, int64_t nq, int k, int metric, int32_t* const* flags_host, int rank,
                                         int epoch, float* out_val, int64_t* out_idx, void* stream) {
  KNN_REQUIRE(val_parts_host && idx_parts_host && flags_host && (nq == 0 || (out_val && out_idx)),
              "knn_merge_topk_parts_sync: null pointer");
  KNN_REQUIRE(parts >= 1 && parts <= kMaxMergeParts && parts <= 32 && k >= 1 && nq >= 0 && rank >= 0 && rank < parts,
              "knn_merge_topk_parts_sync: bad sizes parts=%d (max %d) rank=%d k=%d nq=%lld", parts, kMaxMergeParts, rank, k,
              (long long)nq);
  MergeParts mp;
  PeerFlags pf;
  for (int p = 0; p < parts; ++p) {
    KNN_REQUIRE(val_parts_host[p] && idx_parts_host[p] && flags_host[p], "knn_merge_topk_parts_sync: null part %d", p);
    mp.idx[p] = idx_parts_host[p];
    mp.val[p] = val_parts_host[p];
    pf.flags[p] = flags_host[p];
  }
  return launch_merge_parts(mp, parts, nq, k, metric, out_val, out_idx, (cudaStream_t)stream, &pf, rank, epoch);
}

// The two halves of knn_merge_topk_parts_sync as separate calls, for a PIPELINED exchange: the publish runs on the search
// stream right behind the local search, the wait + merge on a side stream, so the next local search does not wait for the
// slowest shard (sharded.py).
extern "C" int knn_peer_publish(int32_t* const* flags_host, int parts, int rank, int epoch, void* stream) {
  KNN_REQUIRE(flags_host && parts >= 1 && parts <= kMaxMergeParts && parts <= 32 && rank >= 0 && rank < parts,
              "knn_peer_publish: bad arguments parts=%d rank=%d", parts, rank);
  PeerFlags pf;
  for (int p = 0; p < parts; ++p) {
    KNN_REQUIRE(flags_host[p], "knn_peer_publish: null flag array %d", p);
    pf.flags[p] = flags_host[p];
  }
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, parts, rank, epoch);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_merge_topk_parts_wait(const float* const* val_parts_host, const int64_t* const* idx_parts_host,
                                         int parts, int64_t nq, int k, int metric, int32_t* const* flags_host, int rank,
                                         int epoch, float* out_val, int64_t* out_idx, void* stream) {
  KNN_REQUIRE(val_parts_host && idx_parts_host && flags_host && (nq == 0 || (out_val && out_idx)),
              "knn_merge_topk_parts_wait: null pointer");
  KNN_REQUIRE(parts >= 1 && parts <= kMaxMergeParts && parts <= 32 && k >= 1 && nq >= 0 && rank >= 0 && rank < parts,
              "knn_merge_topk_parts_wait: bad sizes parts=%d (max %d) rank=%d k=%d nq=%lld", parts, kMaxMergeParts, rank, k,
              (long long)nq);
  MergeParts mp;
  PeerFlags pf;
  for (int p = 0; p < parts; ++p) {
    KNN_REQUIRE(val_parts_host[p] && idx_parts_host[p] && flags_host[p], "knn_merge_topk_parts_wait: null part %d", p);
    mp.idx[p] = idx_parts_host[p];
    mp.val[p] = val_parts_host[p];
    pf.flags[p] = flags_host[p];
  }
  return launch_merge_parts(mp, parts, nq, k, metric, out_val, out_idx, (cudaStream_t)stream, &pf, rank, epoch, false);
}

static int64_t rank_npad(int64_t ng) {
  int64_t p = 2;
  while (p < ng) p <<= 1;
  return p;
}

extern "C" size_t knn_rank_rows_workspace(int64_t nq, int64_t ng) {
  if (nq <= 0 || ng <= 0) return 0;
  const int64_t npad = rank_npad(ng);
  if (npad <= kSortCap) return 0;
  return (size_t)nq * (size_t)npad * sizeof(uint64_t);
}

extern "C" int knn_rank_rows(const float* scores, int64_t nq, int64_t ng, int largest_first, int64_t* ranks,
                             void* workspace, size_t workspace_bytes, void* stream) {
  KNN_REQUIRE(nq >= 0 && ng >= 0 && ng < 0xFFFFFFFFll, "knn_rank_rows: bad sizes");
  if (nq == 0 || ng == 0) return KNN_OK;
  KNN_REQUIRE(scores && ranks, "knn_rank_rows: null pointer");
  const size_t need = knn_rank_rows_workspace(nq, ng);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
    set_error("knn_rank_rows: workspace too small (%zu < %zu)", workspace_bytes, need);
    return KNN_E_WORKSPACE;
  }
  rank_rows_kernel<<<(unsigned)nq, kSortThreads, 0, (cudaStream_t)stream>>>(
      scores, nq, ng, largest_first, rank_npad(ng), reinterpret_cast<uint64_t*>(workspace), ranks);
  KNN_LAUNCHED();
  return KNN_OK;
}
