// Fused top-k selection shared by the fp32 (FFMA) and bf16 (tcgen05) distance kernels.
//
// Ownership model: inside a CTA every query row is owned by exactly ONE thread (the thread that reads
// that row's scores: TMEM lane i <-> thread i of the epilogue warps).  The owner keeps, in registers,
//   cnt  - number of candidates currently in the row's list
//   tau  - score of the current k-th best candidate (-inf until k candidates were seen)
// and appends every score that can still enter the top-k to a per-(unit,row) list in the global
// workspace (L2 resident, capacity L = 2*KP >= 2k).  99.9 % of the scores are rejected by one max +
// one compare per 32-score chunk.  When a list is about to overflow the WARP compacts it
// cooperatively (bitonic sort of L 64-bit keys held L/32 per lane), keeps the best k and tightens tau.
// Thresholds are shared between all CTAs working on the same query row through tau_global
// (atomicMax of the order-preserving encoding), so late gallery splits start with a tight filter.
//
// Exactness: a candidate is dropped only if its score is < tau where tau is the k-th best key of some
// subset of the gallery, hence <= the final k-th best; every member of the true top-k (ties resolved
// by ascending gallery row through the key encoding) therefore survives to the final merge.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace knn {

constexpr unsigned kFullMask = 0xFFFFFFFFu;

template <int E>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&v)[E], int lane) {
  // Bitonic network over n = 32*E keys, element index i = e*32 + lane, final order descending in i.
  constexpr int N = 32 * E;
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int se = stride >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & se) == 0) {
            const int i = e * 32 + lane;
            const bool desc = ((i & size) == 0);
            uint64_t a = v[e], b = v[e | se];
            uint64_t hi = a > b ? a : b, lo = a > b ? b : a;
            v[e] = desc ? hi : lo;
            v[e | se] = desc ? lo : hi;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = e * 32 + lane;
          const bool desc = ((i & size) == 0);
          const bool lower = ((lane & stride) == 0);
          uint64_t a = v[e];
          uint64_t b = __shfl_xor_sync(kFullMask, a, stride);
          const bool keep_max = (lower == desc);
          v[e] = keep_max ? (a > b ? a : b) : (a > b ? b : a);
        }
      }
    }
  }
}

// Warp-cooperative compaction of one row's list: sort, keep the best `keep` (k, or KP at unit end),
// return the new count and (if >= k candidates exist) the new threshold key.  All 32 lanes must call.
// Deliberately NOT inlined: the sorting network is ~2.5k instructions and runs rarely; one copy per kernel keeps
// the hot epilogue loop inside the instruction cache.
struct CompactOut {
  int cnt;
  uint64_t taukey;
};
template <int E>
__device__ __noinline__ CompactOut compact_row(uint64_t* __restrict__ list, int cnt, int k, int keep, int lane) {
  uint64_t v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    v[e] = (i < cnt) ? __ldcg(list + i) : 0ull;
  }
  warp_sort_desc<E>(v, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    if (i < keep) __stcg(list + i, v[e]);
  }
  uint64_t kth = 0;
  const int ke = (k - 1) >> 5, kl = (k - 1) & 31;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    uint64_t t = __shfl_sync(kFullMask, v[e], kl);
    if (e == ke) kth = t;
  }
  CompactOut o;
  o.cnt = cnt < k ? cnt : k;
  o.taukey = (cnt >= k) ? kth : 0ull;  // 0 = "no threshold yet"
  return o;
}

// Cheap compaction by SELECTION instead of sorting (the list does not need to be ordered, only trimmed): a pivot
// key p with at least k list entries >= p is a valid threshold, and every entry below it can be dropped.  The warp
// sorts a 32-key sample of the list (one key per lane), binary-searches the sample for the LARGEST sample with
// count(list >= sample) >= k (5 counting rounds: E compares per lane + one warp reduction each), and writes the
// survivors back contiguously (unordered).  ~350 instructions instead of the ~2.5k of the full sorting network, so
// a compaction no longer holds an accumulator stage long enough to stall the tensor pipe.  Expected survivors:
// k .. k + ~L/32.  Falls back to the sorting network when the sample cannot shrink the list (returns cnt unchanged).
template <int E>
__device__ __noinline__ CompactOut compact_row_select(uint64_t* __restrict__ list, int cnt, int k, int lane) {
  uint64_t v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    v[e] = (i < cnt) ? __ldcg(list + i) : 0ull;
  }
  // sample: lane l offers its key in slot (l mod E) -> positions spread over the whole list
  uint64_t smp = 0ull;
#pragma unroll
  for (int e = 0; e < E; ++e)
    if ((lane % E) == e) smp = v[e];
  // 32-key bitonic sort across lanes, descending in lane index
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const uint64_t o = __shfl_xor_sync(kFullMask, smp, stride);
      const bool desc = (lane & size) == 0;
      const bool lower = (lane & stride) == 0;
      const bool keep_max = lower == desc;
      smp = keep_max ? (smp > o ? smp : o) : (smp > o ? o : smp);
    }
  }
  // smallest sample index j whose count(list >= sample[j]) >= k  (counts grow with j)
  int lo = 0, hi = 32;  // answer in [lo, hi]; hi == 32 means "no sample qualifies"
#pragma unroll 1
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const uint64_t pv = __shfl_sync(kFullMask, smp, mid);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (v[e] >= pv) ? 1 : 0;
    c = __reduce_add_sync(kFullMask, c);
    if (pv != 0ull && c >= k) hi = mid; else lo = mid + 1;
  }
  CompactOut o;
  o.cnt = cnt;
  o.taukey = 0ull;
  if (lo >= 32) return o;  // (near-)degenerate sample: caller falls back to the sorting network
  const uint64_t pv = __shfl_sync(kFullMask, smp, lo);
  int base = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool keep = v[e] >= pv;
    const unsigned b = __ballot_sync(kFullMask, keep);
    if (keep) __stcg(list + base + __popc(b & ((1u << lane) - 1u)), v[e]);
    base += __popc(b);
  }
  o.cnt = base;
  o.taukey = pv;
  return o;
}

struct RowState {
  uint64_t* list;   // this row's candidate list (capacity L)
  int cnt;
  float tau;        // exact score threshold (larger = better)
  float ftau;       // threshold in "filter space" (== tau for similarities, -d^2 bound for L2)
  uint64_t taukey;  // a candidate must have key > taukey (exact, tie-aware form of tau)
};

__device__ __forceinline__ void rowstate_init(RowState& st, uint64_t* list) {
  st.list = list;
  st.cnt = 0;
  st.tau = -INFINITY;
  st.ftau = -INFINITY;
  st.taukey = 0ull;
}

template <bool kL2>
__device__ __forceinline__ float filter_tau(float tau) {
  if (!kL2) return tau;
  // tau = -d_k.  A candidate qualifies iff d <= d_k with d = rn(sqrt(d2)); conservative bound on d2:
  // d2 <= d_k^2 * (1 + 4e-6) never rejects a qualifying candidate.
  if (tau == -INFINITY) return -INFINITY;
  float dk = -tau;
  return -(dk * dk * 1.000004f + 1e-37f);
}

template <bool kL2>
__device__ __forceinline__ float exact_score(float f) {
  // f is the filter value: dot for similarities, -(|q|^2+|g|^2-2 q.g) for L2.
  if (!kL2) return f;
  return -__fsqrt_rn(fmaxf(-f, 0.0f));
}

// Slow path of the selection, shared by all kernels and NOT inlined (one compact loop instead of 32 unrolled,
// predicated copies per call site).  The chunk's raw dot products were staged in shared memory by the caller:
// value j of this row lives at sv[(j >> 2) * quad_stride + (j & 3)].  `mask` has bit j set when the filter value
// of column col0 + j passed the cheap threshold test.  Returns the new list length.
template <bool kL2>
__device__ __noinline__ int append_hits(uint64_t* __restrict__ list, int cnt, uint64_t taukey, const float* sv,
                                        int quad_stride, float qn, const float* __restrict__ gn, uint32_t mask,
                                        uint32_t col0, uint32_t self_row, int self_mode) {
  while (mask) {
    const int j = __ffs(mask) - 1;
    mask &= mask - 1;
    float f = sv[(j >> 2) * quad_stride + (j & 3)];
    if (kL2) f = fmaf(2.0f, f, -(qn + gn[j]));
    const uint32_t row = col0 + (uint32_t)j;
    float s = exact_score<kL2>(f);
    bool take = true;
    if (row == self_row) {
      if (self_mode == KNN_SELF_EXCLUDE) take = false;
      else if (self_mode == KNN_SELF_MINUS1) s = -1.0f;
    }
    const uint64_t key = make_key(s, row);
    if (take && key > taukey) {
      __stcg(list + cnt, key);
      ++cnt;
    }
  }
  return cnt;
}

// Max of 32 values as a shallow tree of 3-input maxima (FMNMX3): 16 instructions, depth 4.
template <class FV>
__device__ __forceinline__ float chunk_max32(FV fv) {
  float t[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) t[i] = fmaxf(fmaxf(fv(3 * i), fv(3 * i + 1)), fv(3 * i + 2));
  t[10] = fmaxf(fv(30), fv(31));
  const float a = fmaxf(fmaxf(t[0], t[1]), t[2]), b = fmaxf(fmaxf(t[3], t[4]), t[5]);
  const float c = fmaxf(fmaxf(t[6], t[7]), t[8]), d = fmaxf(t[9], t[10]);
  return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

// Bit j set iff fv(j) >= ftau and column j is valid.
template <class FV>
__device__ __forceinline__ uint32_t chunk_mask32(FV fv, float ftau, uint32_t ncols_valid) {
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) mask |= (fv(j) >= ftau) ? (1u << j) : 0u;
  return ncols_valid >= 32 ? mask : (mask & ((1u << ncols_valid) - 1u));
}

// One chunk of 32 filter values owned by this thread (fv(j): value of column col0+j, larger = better) whose raw
// dot products are ALSO available in memory (sv/quad_stride, see append_hits).  Fast path: max tree + compare.
template <bool kL2, class FV>
__device__ __forceinline__ void select_chunk_mem(RowState& st, FV fv, const float* sv, int quad_stride, float qn,
                                                 const float* gn, uint32_t col0, uint32_t ncols_valid,
                                                 uint32_t self_row, int self_mode, bool row_valid) {
  const float m = chunk_max32(fv);
  if (row_valid && m >= st.ftau) {
    const uint32_t mask = chunk_mask32(fv, st.ftau, ncols_valid);
    if (mask) st.cnt = append_hits<kL2>(st.list, st.cnt, st.taukey, sv, quad_stride, qn, gn, mask, col0, self_row, self_mode);
  }
}

// Same for a chunk held in REGISTERS (tcgen05.ld output): on the slow path the 32 raw dot products are first
// staged in this thread's private slice of a shared-memory scratch area laid out [8 quads][128 threads] float4
// (conflict-free STS.128), so append_hits can index them dynamically.  `dump` points at this thread's float4 slot 0.
constexpr int kDumpQuadStride = 128 * 4;                 // floats between consecutive quads of one thread
constexpr int kDumpBytes = 8 * 128 * 16;                 // 16 KB per 128 row-owner threads
template <bool kL2>
__device__ __forceinline__ void select_chunk_regs(RowState& st, const uint32_t (&v)[32], float* dump, float qn,
                                                  const float* gn, uint32_t col0, uint32_t ncols_valid,
                                                  uint32_t self_row, int self_mode, bool row_valid) {
  auto fv = [&](int j) -> float {
    const float dot = __uint_as_float(v[j]);
    if (kL2) return fmaf(2.0f, dot, -(qn + gn[j]));
    return dot;
  };
  const float m = chunk_max32(fv);
  if (row_valid && m >= st.ftau) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<uint4*>(dump + q * kDumpQuadStride) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    const uint32_t mask = chunk_mask32(fv, st.ftau, ncols_valid);
    if (mask) st.cnt = append_hits<kL2>(st.list, st.cnt, st.taukey, dump, kDumpQuadStride, qn, gn, mask, col0, self_row, self_mode);
  }
}

// After a chunk was appended: compact every row of this warp whose list could overflow on the next
// chunk (cnt > L - CH).  Warp-uniform control flow; `valid` rows only.
template <int E, int CH, bool kL2>
__device__ __forceinline__ void warp_compact_if_needed(RowState& st, int k, int lane,
                                                       uint32_t* __restrict__ tau_global_row) {
  constexpr int L = 32 * E;
#ifdef KNN_BOUNDS_CHECK  // checked build (build.py --check): the list invariants every append relies on
  if (st.cnt < 0 || st.cnt > L) {
    printf("b200knn check: candidate list overflow, cnt=%d capacity=%d (block %d,%d thread %d)\n", st.cnt, L,
           blockIdx.x, blockIdx.y, threadIdx.x);
    __trap();
  }
#endif
  unsigned need = __ballot_sync(kFullMask, st.cnt > L - CH);
  while (need) {
    const int r = __ffs(need) - 1;
    need &= need - 1;
    uint64_t* rl = (uint64_t*)__shfl_sync(kFullMask, (unsigned long long)st.list, r);
    const int rc = __shfl_sync(kFullMask, st.cnt, r);
    __syncwarp();
    // The list must end up with room for a whole chunk (cnt <= L - CH is the invariant every append relies on) and
    // should end up well below capacity.  The selection leaves k .. k + ~L/32 keys; when that cannot meet the target
    // (k close to L - CH: the shortest lists, L = 64), or did not, sort and keep exactly k.  The selection may have
    // rewritten the list (survivors in front, stale copies behind): the sort reads co.cnt entries, not the old count.
    constexpr int kRoom = L - CH;
    const int keep_max = (L + k) / 2 < kRoom ? (L + k) / 2 : kRoom;
    CompactOut co;
    if (k + L / 32 > keep_max) {
      co = compact_row<E>(rl, rc, k, k, lane);
    } else {
      co = compact_row_select<E>(rl, rc, k, lane);
      if (co.cnt > keep_max) co = compact_row<E>(rl, co.cnt, k, k, lane);
    }
    const int nc = co.cnt;
    const uint64_t nk = co.taukey;
    if (lane == r) {
      st.cnt = nc;
      if (nk > st.taukey) {
        st.taukey = nk;
        st.tau = key_score(nk);
        st.ftau = filter_tau<kL2>(st.tau);
        if (tau_global_row) atomicMax(tau_global_row, (uint32_t)(nk >> 32));
      }
    }
    __syncwarp();
  }
#ifdef KNN_BOUNDS_CHECK
  if (st.cnt > L - CH) {
    printf("b200knn check: list left with %d keys, more than capacity - chunk = %d (block %d,%d thread %d)\n", st.cnt,
           L - CH, blockIdx.x, blockIdx.y, threadIdx.x);
    __trap();
  }
#endif
}

// Pull the shared threshold (other CTAs working on the same query row may have tightened it).
// Split in two so the L2 round trip of the load can be hidden behind a barrier wait: o = peek_tau(); ...; apply_tau(o).
__device__ __forceinline__ uint32_t peek_tau(const uint32_t* __restrict__ tau_global_row) {
  return tau_global_row ? __ldcg(tau_global_row) : 0u;
}
template <bool kL2>
__device__ __forceinline__ void apply_tau(RowState& st, uint32_t o) {
  if (o != 0u && ord2f(o) > st.tau) {
    st.tau = ord2f(o);
    st.ftau = filter_tau<kL2>(st.tau);
    st.taukey = (uint64_t)o << 32;
  }
}
template <bool kL2>
__device__ __forceinline__ void refresh_tau(RowState& st, const uint32_t* __restrict__ tau_global_row) {
  if (!tau_global_row) return;
  uint32_t o = __ldcg(tau_global_row);
  // The shared value is a SCORE bound only: any key with that score may still qualify, so the key
  // form is (ord << 32) | 0 (the worst key of that score); only adopt it when strictly tighter.
  if (o != 0u && ord2f(o) > st.tau) {
    st.tau = ord2f(o);
    st.ftau = filter_tau<kL2>(st.tau);
    st.taukey = (uint64_t)o << 32;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Selection over one accumulator tile that lives in TENSOR MEMORY (tcgen05 kernels).  The calling thread owns TMEM
// lane `taddr.lane + lane` = one query row and examines the 32-column chunks first, first + step, ... < nchunks.
// Two phases, so that the accumulator stage is held only as long as the scores must be READ:
//   select_tile_tmem   (stage held)     per chunk one tcgen05.ld.x32, a max tree, one compare.  A lane whose chunk
//                      has exactly ONE passing score -- the chunk maximum, already in a register -- just records
//                      (column, value) in a two-entry register queue.  Several passing scores in one lane's chunk,
//                      or a full queue, are rare: those lanes append at once, re-reading the flagged columns from
//                      TMEM with warp-wide single-column loads.
//   flush_pending_hits (stage released) key construction, self handling, the list append and, when a list is
//                      about to overflow, its compaction.
// ------------------------------------------------------------------------------------------------------------
struct PendingHits {
  uint32_t col[2];  // local gallery row
  float val[2];     // filter value (dot, or -(d^2) for L2)
  int n;
};

template <bool kL2>
__device__ __forceinline__ void append_candidate(RowState& st, float f, uint32_t row, uint32_t self_row,
                                                 int self_mode) {
  float sc = exact_score<kL2>(f);
  bool take = true;
  if (row == self_row) {
    if (self_mode == KNN_SELF_EXCLUDE) take = false;
    else if (self_mode == KNN_SELF_MINUS1) sc = -1.0f;
  }
  const uint64_t key = make_key(sc, row);
  if (take && key > st.taukey) {
    __stcg(st.list + st.cnt, key);
    ++st.cnt;
  }
}

template <int E, bool kL2>
__device__ __forceinline__ void select_tile_tmem(RowState& st, PendingHits& pend, uint32_t taddr, int first, int step,
                                                 int nchunks, int64_t col0, int64_t c_end, const float* gst, float qn,
                                                 uint32_t self_row, int self_mode, int k, int lane,
                                                 uint32_t* tau_row, bool row_valid, bool stats_on,
                                                 long long& e_slow, unsigned long long& n_slow) {
#pragma unroll 1
  for (int ch = first; ch < nchunks; ch += step) {
    const int cb = ch * 32;
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)cb, v);
    ptx::tmem_ld_fence(v);
    const int64_t cg = col0 + cb;
    const int64_t rem = c_end - cg;
    const uint32_t nvalid = rem <= 0 ? 0u : (rem >= 32 ? 32u : (uint32_t)rem);
    const float* gn = gst + cb;
    auto fv = [&](int j) -> float {
      const float dot = __uint_as_float(v[j]);
      if (kL2) return fmaf(2.0f, dot, -(qn + gn[j]));
      return dot;
    };
    const float m = chunk_max32(fv);
    const bool hit = row_valid && m >= st.ftau;
    if (__any_sync(kFullMask, hit)) {
      const long long c0 = stats_on ? clock64() : 0;
      uint32_t mask = hit ? chunk_mask32(fv, st.ftau, nvalid) : 0u;
      if (mask != 0u && (mask & (mask - 1u)) == 0u && pend.n < 2 && nvalid == 32u) {  // (tail chunk: the maximum may sit in a padding column)
        // the single passing score is the chunk maximum: queue it, the list work waits until the stage is released
        const uint32_t row = (uint32_t)cg + (uint32_t)(__ffs(mask) - 1);
        if (pend.n == 0) { pend.col[0] = row; pend.val[0] = m; }
        else { pend.col[1] = row; pend.val[1] = m; }
        ++pend.n;
        mask = 0u;
      }
      uint32_t uni = __reduce_or_sync(kFullMask, mask);
      if (uni) {  // rare: some lane has several hits in this chunk (or a full queue) -> append them now
        while (uni) {
          const int j = __ffs(uni) - 1;
          uni &= uni - 1;
          const uint32_t x = ptx::tmem_ld_32x32_x1(taddr + (uint32_t)(cb + j));
          if ((mask >> j) & 1u) {
            float f = __uint_as_float(x);
            if (kL2) f = fmaf(2.0f, f, -(qn + gn[j]));
            append_candidate<kL2>(st, f, (uint32_t)cg + (uint32_t)j, self_row, self_mode);
          }
        }
        warp_compact_if_needed<E, 32, kL2>(st, k, lane, tau_row);
      }
      if (stats_on) {
        e_slow += clock64() - c0;
        ++n_slow;
      }
    }
  }
}

template <int E, bool kL2>
__device__ __forceinline__ void flush_pending_hits(RowState& st, PendingHits& pend, uint32_t self_row, int self_mode,
                                                   int k, int lane, uint32_t* tau_row, bool stats_on,
                                                   long long& e_slow) {
  if (__any_sync(kFullMask, pend.n > 0)) {
    const long long c0 = stats_on ? clock64() : 0;
    if (pend.n > 0) append_candidate<kL2>(st, pend.val[0], pend.col[0], self_row, self_mode);
    if (pend.n > 1) append_candidate<kL2>(st, pend.val[1], pend.col[1], self_row, self_mode);
    pend.n = 0;
    warp_compact_if_needed<E, 32, kL2>(st, k, lane, tau_row);
    if (stats_on) e_slow += clock64() - c0;
  }
}

// Threshold-seeding pass in MAXIMA mode (SearchParams::seed_stride > 0).  The pre-pass only has to produce a valid LOWER
// BOUND of every query's final k-th best score, not the sample's top-k: the best score of each of n DISJOINT groups of
// gallery rows belongs to n distinct rows, so the k-th largest of those maxima is reached by at least k rows.  A
// selection thread therefore keeps ONE running maximum over `stride` consecutive chunks of its own (the chunk maximum
// is what the fast path of select_tile_tmem computes anyway) and appends it as a key -- no threshold, no compaction,
// none of the accept-everything warm-up a selecting unit pays (which was 85 % of the pre-pass at 8192 x 50 M).  With
// 32..64 rows per group the k-th largest maximum sits within a few ranks of the sample's exact k-th best score.
// Chunks that hold the query's own row (self modes) or padding columns contribute nothing: still a valid bound.
// The key's row is the first column of the chunk that gave the maximum -- distinct for distinct groups.
struct SeedRun {
  float best;
  uint32_t col;
  int since;
};
template <bool kL2>
__device__ __forceinline__ void seed_flush(RowState& st, SeedRun& run, bool row_valid) {
  if (row_valid && run.best > -INFINITY) {
    __stcg(st.list + st.cnt, make_key(exact_score<kL2>(run.best), run.col));
    ++st.cnt;
  }
  run.best = -INFINITY;
  run.since = 0;
}
template <bool kL2>
__device__ __forceinline__ void seed_tile_tmem(RowState& st, SeedRun& run, int stride, uint32_t taddr, int first,
                                               int step, int nchunks, int64_t col0, int64_t c_end, const float* gst,
                                               float qn, uint32_t self_row, bool row_valid) {
#pragma unroll 1
  for (int ch = first; ch < nchunks; ch += step) {
    const int cb = ch * 32;
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)cb, v);
    ptx::tmem_ld_fence(v);
    const int64_t cg = col0 + cb;
    const float* gn = gst + cb;
    auto fv = [&](int j) -> float {
      const float dot = __uint_as_float(v[j]);
      if (kL2) return fmaf(2.0f, dot, -(qn + gn[j]));
      return dot;
    };
    const float m = chunk_max32(fv);
    const bool whole = cg + 32 <= c_end;                            // no padding column in this chunk
    const bool has_self = self_row - (uint32_t)cg < 32u;            // self_row = 0xFFFFFFFF: never
    if (whole && !has_self && m > run.best) {                       // (a NaN maximum fails the compare: skipped)
      run.best = m;
      run.col = (uint32_t)cg;
    }
    if (++run.since >= stride) seed_flush<kL2>(st, run, row_valid);
  }
}

// Dense mode of the tcgen05 kernels: instead of selecting, the owners of a query row write its scores of this tile to the
// dense [nq, ld] block (similarity, or the positive L2 distance; the query's own column masked as knn_scores_dense does).
// Every lane executes the (warp-collective) tensor-memory loads; a lane without a valid row just does not store.
template <bool kL2>
__device__ __forceinline__ void dense_store_tile_tmem(uint32_t taddr, int first, int step, int nchunks, int64_t col0,
                                                      int64_t c_end, const float* gst, float qn, int64_t row,
                                                      int64_t self_col, int self_mode, float* __restrict__ out,
                                                      int64_t ld) {
#pragma unroll 1
  for (int ch = first; ch < nchunks; ch += step) {
    const int cb = ch * 32;
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)cb, v);
    ptx::tmem_ld_fence(v);
    if (row < 0) continue;
    const int64_t cg = col0 + cb;
    if (cg >= c_end) continue;
    float* dst = out + row * ld + cg;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = __uint_as_float(v[j]);
      if (kL2) x = __fsqrt_rn(fmaxf(-fmaf(2.0f, x, -(qn + gst[cb + j])), 0.0f));
      f[j] = x;
    }
    if (self_col >= cg && self_col < cg + 32) {
      const float m = self_mode == KNN_SELF_EXCLUDE ? (kL2 ? INFINITY : -INFINITY) : (kL2 ? 1.0f : -1.0f);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cg + j == self_col) f[j] = m;
    }
    if (cg + 32 <= c_end && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        __stcs(reinterpret_cast<float4*>(dst + j), make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cg + j < c_end) __stcs(dst + j, f[j]);
    }
  }
}

}  // namespace knn
