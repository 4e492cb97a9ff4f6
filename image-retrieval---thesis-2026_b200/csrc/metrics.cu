// On-device retrieval metrics over ranked lists.  Per-query quantities are computed here with the
// reference's own operation order in IEEE double (explicit _rn intrinsics: no FMA contraction), so the
// Python layer only has to take the final mean the way the reference does (np.mean / Python sum).
#include "common.cuh"

namespace knn {
namespace {

__global__ void relevance_single_kernel(const int64_t* __restrict__ idx, int64_t total, int k,
                                        const int64_t* __restrict__ qlab, const int64_t* __restrict__ glab,
                                        int64_t ng, uint8_t* __restrict__ rel, int64_t* __restrict__ rlab) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t id = idx[t];
  uint8_t r = 0;
  int64_t lab = INT64_MIN;
  if (id >= 0 && id < ng) {
    lab = glab[id];
    r = (lab == qlab[t / k]) ? 1 : 0;
  }
  if (rel) rel[t] = r;
  if (rlab) rlab[t] = lab;
}

// Jaccard exactly as the reference evaluates it:
//  arith 0 (fp32): intersect / (union + 1e-8) > thr with fp32 tensors and the threshold rounded to fp32
//                  (train.py:462-466, nih_multilabel_training.py:90-93, test.py:956-965)
//  arith 1 (fp64): Python floats (evaluate_nih_zilliz.py:12-17,44)
__global__ void relevance_multilabel_kernel(const int64_t* __restrict__ idx, int64_t total, int k,
                                            const uint64_t* __restrict__ qmask, const uint64_t* __restrict__ gmask,
                                            int64_t ng, double thr, int arith, uint8_t* __restrict__ rel_j,
                                            uint8_t* __restrict__ rel_any) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t id = idx[t];
  uint8_t rj = 0, ra = 0;
  if (id >= 0 && id < ng) {
    const uint64_t a = qmask[t / k], b = gmask[id];
    const int inter = __popcll(a & b), uni = __popcll(a | b);
    ra = inter > 0;
    if (arith == 0) {
      const float j = __fdiv_rn((float)inter, __fadd_rn((float)uni, 1e-8f));
      rj = j > (float)thr;
    } else {
      const double j = __ddiv_rn((double)inter, __dadd_rn((double)uni, 1e-8));
      rj = j > thr;
    }
  }
  if (rel_j) rel_j[t] = rj;
  if (rel_any) rel_any[t] = ra;
}

__global__ void ranked_stats_kernel(const uint8_t* __restrict__ rel, int64_t nq, int k, int kk,
                                    int32_t* __restrict__ hits, int32_t* __restrict__ first,
                                    double* __restrict__ ap_topk, double* __restrict__ prec_sum) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const uint8_t* r = rel + q * k;
  int pos = 0, f = 0;
  double ps = 0.0;
  for (int j = 0; j < kk; ++j) {
    if (r[j]) {
      ++pos;
      ps = __dadd_rn(ps, __ddiv_rn((double)pos, (double)(j + 1)));  // precision_sum += positives / rank
      if (f == 0) f = j + 1;
    }
  }
  if (hits) hits[q] = pos;
  if (first) first[q] = f;
  if (prec_sum) prec_sum[q] = ps;
  if (ap_topk) ap_topk[q] = pos > 0 ? __ddiv_rn(ps, (double)pos) : 0.0;
}

// All cut-offs of a metric table in ONE launch: thread q walks its row once and snapshots the running statistics at
// every kks[t] (any order, each <= k).  Outputs are [nq, nk] (first: [nq], the rank of the first relevant item).
constexpr int kMaxCutoffs = 16;
__global__ void ranked_stats_multi_kernel(const uint8_t* __restrict__ rel, int64_t nq, int k,
                                          const int32_t* __restrict__ kks, int nk, int32_t* __restrict__ hits,
                                          int32_t* __restrict__ first, double* __restrict__ ap_topk,
                                          double* __restrict__ prec_sum) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const uint8_t* r = rel + q * k;
  int cut[kMaxCutoffs], kmax = 0;
  for (int t = 0; t < nk; ++t) {
    cut[t] = kks[t] < k ? kks[t] : k;
    kmax = cut[t] > kmax ? cut[t] : kmax;
  }
  int pos = 0, f = 0;
  double ps = 0.0;
  for (int j = 0; j < kmax; ++j) {
    if (r[j]) {
      ++pos;
      ps = __dadd_rn(ps, __ddiv_rn((double)pos, (double)(j + 1)));  // precision_sum += positives / rank
      if (f == 0) f = j + 1;
    }
    for (int t = 0; t < nk; ++t)
      if (cut[t] == j + 1) {
        if (hits) hits[q * nk + t] = pos;
        if (prec_sum) prec_sum[q * nk + t] = ps;
        if (ap_topk) ap_topk[q * nk + t] = pos > 0 ? __ddiv_rn(ps, (double)pos) : 0.0;
      }
  }
  if (first) first[q] = f;   // over the first max(kks) items
}

// R@K in one launch: counts[t] = number of queries with a label match among their first kks[t] retrieved rows
// (retrieval_accuracy, test.py:38-54: `correct[:k].any()` summed over the batch).  counts must be zeroed by the caller.
__global__ void recall_counts_kernel(const int64_t* __restrict__ idx, int64_t nq, int k, const int64_t* __restrict__ qlab,
                                     const int64_t* __restrict__ glab, int64_t ng, const int32_t* __restrict__ kks,
                                     int nk, int32_t* __restrict__ counts) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int first = 0;
  if (q < nq) {
    const int64_t ql = qlab[q];
    for (int j = 0; j < k && first == 0; ++j) {
      const int64_t id = idx[q * k + j];
      if (id >= 0 && id < ng && glab[id] == ql) first = j + 1;
    }
  }
  for (int t = 0; t < nk; ++t) {
    const unsigned b = __ballot_sync(0xFFFFFFFFu, first > 0 && first <= kks[t]);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(counts + t, __popc(b));
  }
}

__global__ void majority_vote_kernel(const int64_t* __restrict__ lab, int64_t nq, int k, int kk, int tie_mode,
                                     int64_t* __restrict__ vote) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int64_t* l = lab + q * k;
  int best = 0;
  int64_t bl = INT64_MIN;
  for (int j = 0; j < kk; ++j) {
    const int64_t lj = l[j];
    bool seen = false;  // count each distinct label once, at its first occurrence (Counter insertion order)
    for (int i = 0; i < j; ++i) seen |= (l[i] == lj);
    if (seen) continue;
    int c = 0;
    for (int i = j; i < kk; ++i) c += (l[i] == lj);
    if (c > best || (tie_mode == 1 && c == best && lj < bl)) { best = c; bl = lj; }
  }
  vote[q] = bl;
}

// one thread per (query, cut-off): votes [nq, nk]
__global__ void majority_vote_multi_kernel(const int64_t* __restrict__ lab, int64_t nq, int k,
                                           const int32_t* __restrict__ kks, int nk, int tie_mode,
                                           int64_t* __restrict__ vote) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nq * nk) return;
  const int64_t q = e / nk;
  const int kk = kks[e % nk] < k ? kks[e % nk] : k;
  const int64_t* l = lab + q * k;
  int best = 0;
  int64_t bl = INT64_MIN;
  for (int j = 0; j < kk; ++j) {
    const int64_t lj = l[j];
    bool seen = false;  // count each distinct label once, at its first occurrence (Counter insertion order)
    for (int i = 0; i < j; ++i) seen |= (l[i] == lj);
    if (seen) continue;
    int c = 0;
    for (int i = j; i < kk; ++i) c += (l[i] == lj);
    if (c > best || (tie_mode == 1 && c == best && lj < bl)) { best = c; bl = lj; }
  }
  vote[e] = bl;
}

// compute_ap / compute_map (test.py:58-146), one warp per query.
__global__ void map_full_kernel(const int64_t* __restrict__ ranks, int64_t nq, int64_t ng,
                                const int64_t* __restrict__ qlab, const int64_t* __restrict__ glab,
                                const int32_t* __restrict__ kappas, int nkappa, double* __restrict__ ap_out,
                                double* __restrict__ prs, int32_t* __restrict__ npos) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  const int64_t ql = qlab[q];
  const int64_t* rk = ranks + q * ng;
  // nres = len(qgnd): all gallery rows with the query's label (the query itself included)
  int cnt = 0;
  for (int64_t j = lane; j < ng; j += 32) cnt += (glab[j] == ql);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, off);
  const int nres = cnt;
  if (nres == 0) {
    if (lane == 0) {
      ap_out[q] = __longlong_as_double(0x7FF8000000000000ll);
      for (int j = 0; j < nkappa; ++j) prs[q * nkappa + j] = __longlong_as_double(0x7FF8000000000000ll);
      if (npos) npos[q] = 0;
    }
    return;
  }
  const double recall_step = __ddiv_rn(1.0, (double)nres);
  double ap = 0.0;
  int j = 0;            // positives seen so far
  int64_t maxpos = 0;   // 1-based position of the last positive
  int within[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // positives with 1-based position <= kappa
  for (int64_t base = 0; base < ng; base += 32) {
    const int64_t pos = base + lane;
    bool m = false;
    if (pos < ng) {
      const int64_t id = rk[pos];
      m = (id >= 0 && id < ng) && (glab[id] == ql);
    }
    unsigned bal = __ballot_sync(0xFFFFFFFFu, m);
    if (lane == 0) {
      while (bal) {
        const int b = __ffs(bal) - 1;
        bal &= bal - 1;
        const int64_t rank = base + b;
        const double p0 = (rank == 0) ? 1.0 : __ddiv_rn((double)j, (double)rank);
        const double p1 = __ddiv_rn((double)(j + 1), (double)(rank + 1));
        ap = __dadd_rn(ap, __ddiv_rn(__dmul_rn(__dadd_rn(p0, p1), recall_step), 2.0));
        ++j;
        maxpos = rank + 1;
        for (int t = 0; t < nkappa && t < 8; ++t) within[t] += (rank + 1 <= kappas[t]);
      }
    }
  }
  if (lane == 0) {
    ap_out[q] = ap;
    if (npos) npos[q] = nres;
    for (int t = 0; t < nkappa && t < 8; ++t) {
      // kq = min(max(pos), kappa); prs = (pos <= kq).sum() / kq
      const int64_t kq = maxpos < kappas[t] ? maxpos : kappas[t];
      int c = within[t];
      if (kq < kappas[t]) c = j;  // every positive lies at or before max(pos)
      prs[q * nkappa + t] = __ddiv_rn((double)c, (double)kq);
    }
  }
}

// numpy's pairwise summation of n doubles (numpy/_core/src/umath/loops_utils.h.src), as used by
// np.sum inside sklearn.metrics.average_precision_score.  One block of the tree (n <= 128):
__device__ double np_block_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  double r[8];
  for (int i = 0; i < 8; ++i) r[i] = a[i];
  int i;
  for (i = 8; i < n - (n % 8); i += 8)
    for (int t = 0; t < 8; ++t) r[t] = __dadd_rn(r[t], a[i + t]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}
// sum(a, n) = sum(a, n2) + sum(a + n2, n - n2), n2 = n/2 rounded down to a multiple of 8, walked with an explicit stack
// (device recursion needs a run-time stack size: the recursive form overflowed the default 1 KB at k ~ 3000).
__device__ double np_pairwise_sum(const double* a, int n) {
  int lo_s[24], n_s[24];
  double left_s[24];
  unsigned char phase_s[24];
  int sp = 0;
  lo_s[0] = 0; n_s[0] = n; phase_s[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    const int lo = lo_s[sp], m = n_s[sp];
    if (phase_s[sp] == 0) {
      if (m <= 128) { ret = np_block_sum(a + lo, m); --sp; continue; }
      int n2 = m / 2;
      n2 -= n2 % 8;
      phase_s[sp] = 1;
      ++sp;
      lo_s[sp] = lo; n_s[sp] = n2; phase_s[sp] = 0;
    } else if (phase_s[sp] == 1) {
      left_s[sp] = ret;
      int n2 = m / 2;
      n2 -= n2 % 8;
      phase_s[sp] = 2;
      ++sp;
      lo_s[sp] = lo + n2; n_s[sp] = m - n2; phase_s[sp] = 0;
    } else {
      ret = __dadd_rn(left_s[sp], ret);
      --sp;
    }
  }
  return ret;
}

// average_precision_score over one ranked list (scores non-increasing): thresholds at the last index of
// every run of equal scores; AP = -sum(diff(recall_rev) * precision_rev[:-1]).
__global__ void ap_sklearn_kernel(const float* __restrict__ val, const uint8_t* __restrict__ rel, int64_t nq, int k,
                                  double* __restrict__ scratch, double* __restrict__ ap) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const float* v = val + q * k;
  const uint8_t* r = rel + q * k;
  double* terms = scratch + q * (int64_t)(k + 1);
  int total = 0;
  for (int j = 0; j < k; ++j) total += r[j] ? 1 : 0;
  if (total == 0) { ap[q] = __longlong_as_double(0x7FF8000000000000ll); return; }
  // forward pass: thresholds t = 0..T-1 (descending score), store precision and recall
  // terms are consumed in reversed threshold order, so first collect (tps, idx) compactly in `terms`
  int T = 0, tps = 0;
  for (int j = 0; j < k; ++j) {
    tps += r[j] ? 1 : 0;
    const bool last_of_run = (j == k - 1) || (v[j + 1] != v[j]);
    if (last_of_run) {
      // pack tps (<= 2^20) and j (<= 2^20) exactly into one double slot
      terms[T++] = (double)tps * 1048576.0 + (double)j;
    }
  }
  // reversed order: i = 0 is the lowest-score threshold
  // term_i = (R_{t-1} - R_t) * P_t  with t = T-1-i, R_{-1} = 0; ap = -sum(terms)
  // compute in place from the back: need R_{t-1}, so walk t descending while reading packed slots
  // into registers first.
  // Unpack into precision / recall on the fly (two passes to keep the packed values intact).
  for (int i = 0; i < T / 2; ++i) { double tmp = terms[i]; terms[i] = terms[T - 1 - i]; terms[T - 1 - i] = tmp; }
  // now terms[i] holds threshold t = T-1-i
  for (int i = 0; i < T; ++i) {
    const double packed = terms[i];
    const int tp = (int)(packed / 1048576.0);
    const int j = (int)(packed - (double)tp * 1048576.0);
    const double prec = __ddiv_rn((double)tp, (double)(j + 1));   // tps / (tps + fps), tps + fps = j + 1
    const double rec = __ddiv_rn((double)tp, (double)total);
    double rec_prev = 0.0;                                        // recall of threshold t-1 (next slot)
    if (i + 1 < T) {
      const double pk = terms[i + 1];
      const int tp2 = (int)(pk / 1048576.0);
      rec_prev = __ddiv_rn((double)tp2, (double)total);
    }
    terms[i] = __dmul_rn(__dadd_rn(rec_prev, -rec), prec);
  }
  ap[q] = -np_pairwise_sum(terms, T);
}

}  // namespace
}  // namespace knn

using namespace knn;

static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" int knn_relevance_single(const int64_t* idx, int64_t nq, int k, const int64_t* qlab, const int64_t* glab,
                                    int64_t ng, uint8_t* rel, int64_t* retrieved_lab, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && ng >= 0, "knn_relevance_single: bad sizes");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(idx && qlab && (glab || ng == 0), "knn_relevance_single: null pointer");
  const int64_t total = nq * k;
  relevance_single_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(idx, total, k, qlab, glab, ng,
                                                                                   rel, retrieved_lab);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_relevance_multilabel(const int64_t* idx, int64_t nq, int k, const uint64_t* qmask,
                                        const uint64_t* gmask, int64_t ng, double jaccard_thr, int arith,
                                        uint8_t* rel_jaccard, uint8_t* rel_any, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && ng >= 0, "knn_relevance_multilabel: bad sizes");
  KNN_REQUIRE(arith == 0 || arith == 1, "knn_relevance_multilabel: arith must be 0 (fp32) or 1 (fp64)");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(idx && qmask && (gmask || ng == 0), "knn_relevance_multilabel: null pointer");
  const int64_t total = nq * k;
  relevance_multilabel_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      idx, total, k, qmask, gmask, ng, jaccard_thr, arith, rel_jaccard, rel_any);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_ranked_stats(const uint8_t* rel, int64_t nq, int k, int kk, int32_t* hits, int32_t* first,
                                double* ap_topk, double* prec_sum, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && kk >= 1 && kk <= k, "knn_ranked_stats: bad sizes k=%d kk=%d", k, kk);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(rel, "knn_ranked_stats: null pointer");
  ranked_stats_kernel<<<blocks_for(nq, 128), 128, 0, (cudaStream_t)stream>>>(rel, nq, k, kk, hits, first, ap_topk,
                                                                            prec_sum);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_ranked_stats_multi(const uint8_t* rel, int64_t nq, int k, const int32_t* kks, int nk, int32_t* hits,
                                      int32_t* first, double* ap_topk, double* prec_sum, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && nk >= 1 && nk <= kMaxCutoffs, "knn_ranked_stats_multi: bad sizes (1 <= nk <= %d)",
              kMaxCutoffs);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(rel && kks, "knn_ranked_stats_multi: null pointer");
  ranked_stats_multi_kernel<<<blocks_for(nq, 128), 128, 0, (cudaStream_t)stream>>>(rel, nq, k, kks, nk, hits, first,
                                                                                  ap_topk, prec_sum);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_recall_counts(const int64_t* idx, int64_t nq, int k, const int64_t* qlab, const int64_t* glab,
                                 int64_t ng, const int32_t* kks, int nk, int32_t* counts, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && ng >= 0 && nk >= 1, "knn_recall_counts: bad sizes");
  KNN_REQUIRE(counts && kks, "knn_recall_counts: null pointer");
  KNN_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)nk, (cudaStream_t)stream));
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(idx && qlab && (glab || ng == 0), "knn_recall_counts: null pointer");
  recall_counts_kernel<<<blocks_for(nq, 128), 128, 0, (cudaStream_t)stream>>>(idx, nq, k, qlab, glab, ng, kks, nk, counts);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_majority_vote_multi(const int64_t* lab, int64_t nq, int k, const int32_t* kks, int nk, int tie_mode,
                                       int64_t* vote, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && nk >= 1, "knn_majority_vote_multi: bad sizes");
  KNN_REQUIRE(tie_mode == 0 || tie_mode == 1, "knn_majority_vote_multi: bad tie_mode");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(lab && kks && vote, "knn_majority_vote_multi: null pointer");
  majority_vote_multi_kernel<<<blocks_for(nq * nk, 128), 128, 0, (cudaStream_t)stream>>>(lab, nq, k, kks, nk, tie_mode,
                                                                                        vote);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_majority_vote(const int64_t* lab, int64_t nq, int k, int kk, int tie_mode, int64_t* vote,
                                 void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && kk >= 1 && kk <= k, "knn_majority_vote: bad sizes k=%d kk=%d", k, kk);
  KNN_REQUIRE(tie_mode == 0 || tie_mode == 1, "knn_majority_vote: bad tie_mode");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(lab && vote, "knn_majority_vote: null pointer");
  majority_vote_kernel<<<blocks_for(nq, 128), 128, 0, (cudaStream_t)stream>>>(lab, nq, k, kk, tie_mode, vote);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_map_full(const int64_t* ranks, int64_t nq, int64_t ng, const int64_t* qlab, const int64_t* glab,
                            const int32_t* kappas, int nkappa, double* ap, double* prs, int32_t* npos,
                            void* stream) {
  KNN_REQUIRE(nq >= 0 && ng >= 1 && nkappa >= 0 && nkappa <= 8, "knn_map_full: bad sizes (nkappa <= 8)");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(ranks && qlab && glab && ap && (prs || nkappa == 0), "knn_map_full: null pointer");
  map_full_kernel<<<blocks_for(nq, 4), 128, 0, (cudaStream_t)stream>>>(ranks, nq, ng, qlab, glab, kappas, nkappa, ap,
                                                                      prs, npos);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" size_t knn_ap_sklearn_workspace(int64_t nq, int k) {
  return nq > 0 ? (size_t)nq * (size_t)(k + 1) * sizeof(double) : 0;
}

extern "C" int knn_ap_sklearn(const float* val, const uint8_t* rel, int64_t nq, int k, double* ap, void* workspace,
                              size_t workspace_bytes, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && k < (1 << 20), "knn_ap_sklearn: bad sizes");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(val && rel && ap, "knn_ap_sklearn: null pointer");
  if (!workspace || workspace_bytes < knn_ap_sklearn_workspace(nq, k)) {
    set_error("knn_ap_sklearn: workspace too small");
    return KNN_E_WORKSPACE;
  }
  ap_sklearn_kernel<<<blocks_for(nq, 64), 64, 0, (cudaStream_t)stream>>>(val, rel, nq, k,
                                                                        reinterpret_cast<double*>(workspace), ap);
  KNN_LAUNCHED();
  return KNN_OK;
}
