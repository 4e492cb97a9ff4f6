/*
 * b200knn.h -- C ABI of the B200-native exact k-NN retrieval + retrieval-metrics engine.
 *
 * This library replaces ONE hot path of CrispyChillies/Image-Retrieval---Thesis-2026:
 *   L2-normalise -> cosine / inner-product / L2 distance -> top-k ranking -> retrieval metrics.
 * The reference has no FFI for this path (it is inline torch/numpy); each entry point below cites the
 * reference call it stands in for (paths relative to the reference repo root).
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`
 *   - return 0 on success, a negative KNN_E_* code otherwise; knn_last_error() gives a thread-local message
 *   - no allocation / ownership transfer inside the library: the caller passes outputs and a workspace
 *     sized by the matching *_workspace() query
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*), no hidden syncs
 *   - no global mutable state that a result depends on: the error string and the opt-in profile ring (knn_profile_*)
 *     are thread-local, the launch counter (knn_launch_count) is a process-wide atomic, experiment knobs are read from
 *     the environment once; entry points may be called from several threads on different streams / devices
 */
#ifndef B200KNN_H_
#define B200KNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KNN_ABI_VERSION 1

#if defined(__GNUC__)
#define KNN_API __attribute__((visibility("default")))
#else
#define KNN_API
#endif

/* error codes */
#define KNN_OK              0
#define KNN_E_INVALID      -1   /* bad argument (shape, dtype, k, alignment) */
#define KNN_E_WORKSPACE    -2   /* workspace too small / null */
#define KNN_E_CUDA         -3   /* a CUDA runtime / driver call failed */
#define KNN_E_UNSUPPORTED  -4   /* combination not implemented */

/* element types */
#define KNN_F32   0
#define KNN_BF16  1
#define KNN_BF16X3 2  /* knn_search only: bf16 rows written by knn_split_bf16x3 (d = 3 * dpad); same result as KNN_BF16
                         over those rows, but the kernels may load each hi / lo part once for the three products */
#define KNN_BF16X2 4  /* knn_search only: the SAME rows as KNN_BF16X3, but the CTA-pair kernel computes only the two products
                         q_hi.g_hi + q_lo.g_hi (the gallery lo part is neither loaded nor multiplied: 2/3 of the work).
                         The dropped product is bounded by |q| * max|g - g_hi|: knn_filter_error_bound2 */
#define KNN_F32_PACKED 3  /* knn_search / knn_scores_dense / knn_score_stats: q and g are knn_pack_f32 output (fp32 rows
                         transposed into 128-row tiles); d stays the ORIGINAL row width.  Same scores, same bits as
                         KNN_F32 -- the FFMA kernel fills its operand ring with bulk copies instead of transposing every
                         tile through registers (~1.3x the throughput); never takes the small-problem path */

/* metrics: score ordering is always "best first" in the outputs */
#define KNN_COSINE 0   /* inner product of (already normalised) rows, larger = better (test.py:1006, train.py:405) */
#define KNN_IP     1   /* inner product, larger = better (test.py:296 `embeds @ embeds.t()`)                         */
#define KNN_L2     2   /* Euclidean distance sqrt(max(|q|^2+|g|^2-2q.g,0)), smaller = better (test_ath.py:87)         */

/* eps modes of the three normalisation conventions on the path */
#define KNN_EPS_CLAMP 0  /* x / max(||x||, eps)   F.normalize, test.py:1005, fusion_eval/fuse.py:11-15 */
#define KNN_EPS_NONE  1  /* x / ||x||             test.py:251 (zero row -> NaN, reproduced)             */
#define KNN_EPS_ADD   2  /* x / (||x|| + eps)     test.py:444                                           */
#define KNN_CAST_ONLY 3  /* y = x (cast to out_dtype only; rows already normalised by the model, model.py:38) */

/* self handling of self-retrieval (query i is gallery row self_offset+i) */
#define KNN_SELF_KEEP    0
#define KNN_SELF_EXCLUDE 1  /* fill_diagonal_(-inf): test.py:1081, train.py:406 -- the row never appears in top-k */
#define KNN_SELF_MINUS1  2  /* fill_diagonal_(-1):   nih_multilabel_training.py:86 -- row kept with score -1        */

KNN_API int         knn_version(void);
KNN_API const char* knn_last_error(void);

/* Row-wise L2 normalisation fused with the cast to the search dtype.
 * Replaces F.normalize(x, p=2, dim=1) (test.py:1005; train.py:404,450; nih_multilabel_training.py:83;
 * milvus/milvus_retrieval.py:63), `x / x.norm(dim=-1, keepdim=True)` (test.py:251) and l2_normalize
 * (fusion_eval/fuse.py:11-15).  x: [n,d] in_dtype; y: [n,d] out_dtype (must not alias x);
 * sqnorm (nullable): [n] fp32, squared norm of the OUTPUT row as stored (used by the L2 metric). */
KNN_API int knn_normalize(const void* x, void* y, float* sqnorm, int64_t n, int d,
                  int in_dtype, int out_dtype, float eps, int eps_mode, void* stream);

/* Squared row norms of a stored matrix (fp32 accumulate): the |g|^2 term of torch.cdist's GEMM form. */
KNN_API int knn_row_sqnorm(const void* x, float* sqnorm, int64_t n, int d, int dtype, void* stream);

/* Fused distance + top-k.  Replaces `S = q @ g.T` / `-torch.cdist(q, g)` followed by
 * `S.topk(k, 1, True, True)` (test.py:44,1080; train.py:405-409) / `torch.argsort(dist, dim=1)[:, :k]`
 * (test_ath.py:100-114) / a Milvus FLAT `collection.search(limit=k)` (milvus/milvus_retrieval.py:80-86)
 * without materialising the Q x N matrix.
 *   q [nq,d], g [ng,d]: row-major, same dtype (KNN_F32 exact FFMA path, KNN_BF16 tcgen05 path)
 *   q_sqnorm [nq], g_sqnorm [ng]: fp32 squared norms, required for KNN_L2, ignored otherwise
 *   self_mode / self_offset: query i is gallery row (self_offset + i - index_base) of this shard
 *   index_base: added to the local gallery row index in out_idx (row-sharded galleries)
 *   out_val [nq,k] fp32: cosine/ip similarity (descending) or L2 distance (ascending)
 *   out_idx [nq,k] int64: gallery row (+index_base); ties broken by ascending gallery row;
 *                         if fewer than k candidates exist the tail is (-inf | +inf, -1)
 */
KNN_API int knn_search(const void* q, const void* g, const float* q_sqnorm, const float* g_sqnorm,
               int64_t nq, int64_t ng, int d, int dtype, int k, int metric,
               int self_mode, int64_t self_offset, int64_t index_base,
               float* out_val, int64_t* out_idx,
               void* workspace, size_t workspace_bytes, void* stream);
KNN_API size_t knn_search_workspace(int64_t nq, int64_t ng, int d, int dtype, int k);

/* Exact fp32 search on the tensor cores (the KNN_F32 definition of knn_search, reached through a bounded-error
 * filter; csrc/exact_tc.cu).  Same reference calls as knn_search (test.py:44,1006,1080; train.py:405-409).
 * knn_split_bf16x3: error-free split of fp32 rows x [n,d] into bf16 parts hi = bf16(x), lo = bf16(x - hi), written
 *   as ONE row of 3*dpad bf16 (dpad = d rounded up to 8, zero padded): role 0 (queries) [hi|lo|hi], role 1 (gallery)
 *   [hi|hi|lo], so that knn_search(KNN_BF16) over the split rows computes qhi.ghi + qlo.ghi + qhi.glo, which differs
 *   from q.g by at most 8.04 * 2^-18 * |q||g| (bf16 keeps 8 significant bits: |x - hi| <= 2^-8 |x|,
 *   |x - hi - lo| <= 2^-17 |x|; dropped: qlo.glo, (q - qhi - qlo).g, (qhi + qlo).(g - ghi - glo)) plus the accumulation
 *   error of the tensor cores.
 * knn_rescore_exact: cand_val / cand_idx [nq,kc] = the filter's top-kc (kc > k, best first, global rows, -1 = empty).
 *   Every candidate is re-scored with the exact fp32 chain of the KNN_F32 path, the best k are written to
 *   out_val / out_idx exactly as knn_search(KNN_F32) would, and unverified[q] = 0 iff the result is PROVEN to be the
 *   exact top-k: either the filter returned fewer than kc rows (the set is the whole gallery), or the k-th best exact
 *   score beats (worst approximate score in the set) + eps[q], where eps[q] >= |approximate - exact| of the filter
 *   value (the dot product; -(|q|^2 + |g|^2 - 2 q.g) for KNN_L2) for every gallery row.  Queries flagged 1 must be
 *   re-run through knn_search(KNN_F32). */
KNN_API int knn_split_bf16x3(const float* x, int64_t n, int d, int role, void* out, void* stream);
/* The error bound knn_rescore_exact needs.  knn_max_sqnorm: out[0] = max of the gallery's squared row norms (device
 * scalar).  knn_filter_error_bound: eps[q] (rounded up) = u * |q| * max|g|  with
 *   u = 8.04 * 2^-18 (dropped terms of the split) + (3 * dpad / 16 + 1) * 2^-21 * 1.012 (tensor-core accumulation:
 *   at most 2^-21 of the magnitude sum per K = 16 MMA step -- the one hardware assumption, checked against observed
 *   errors by tests/test_gpu_exact_tensor.py AND at run time: knn_rescore_exact flags a query whose candidates show
 *   |filter value - exact value| > eps / 2) + d * 2^-24 * 1.001 (rounding of the exact fp32 chain itself)
 *   + (d / 32 + 6) * 2^-24 (rounding of the fp32 squared norms the bound is built from);
 * KNN_L2: 2 * that + 2^-21 * 1.01 * (|q|^2 + max|g|^2) for the two fp32 roundings of -(|q|^2 + |g|^2 - 2 q.g). */
KNN_API int knn_max_sqnorm(const float* sqnorm, int64_t n, float* out, void* stream);
KNN_API int knn_filter_error_bound(const float* q_sqnorm, int64_t nq, const float* g_sqnorm_max, int d, int metric,
                           float* eps, void* stream);
/* The bound of the TWO-product filter (KNN_BF16X2).  knn_split_lo_max_sqnorm: out[0] = max over the rows of a gallery
 * split (role 1 of knn_split_bf16x3, n rows of 3 * dpad bf16) of the squared norm of the lo part.
 * knn_filter_error_bound2: eps[q] = |q| * (1.0001 * max|g_lo| + u2 * max|g|), u2 = 4.04 * 2^-18 (g - g_hi - g_lo and
 * q - q_hi - q_lo) + (2 * dpad / 16 + 1) * 2^-21 * 1.012 + the chain / norm roundings of knn_filter_error_bound;
 * ~10 x the three-product bound at d = 1024 on typical rows (|g_lo| ~ 0.38 * 2^-8 |g|): the caller re-runs the queries it cannot
 * prove through the three-product filter (search._search_exact_tensor). */
KNN_API int knn_split_lo_max_sqnorm(const void* gallery_split_rows, int64_t n, int d, float* out, void* stream);
KNN_API int knn_filter_error_bound2(const float* q_sqnorm, int64_t nq, const float* g_sqnorm_max,
                            const float* g_lo_sqnorm_max, int d, int metric, float* eps, void* stream);
KNN_API int knn_rescore_exact(const float* q, const float* g, const float* q_sqnorm, const float* g_sqnorm,
                      int64_t nq, int64_t ng, int d, int metric, int self_mode, int64_t self_offset,
                      int64_t index_base, const float* cand_val, const int64_t* cand_idx, int kc, int k,
                      const float* eps, float* out_val, int64_t* out_idx, int32_t* unverified, void* stream);

/* Opt-in, per calling thread: knn_search records CUDA events on its stream around (s) the threshold-seeding
 * pre-pass + seeding merge, (a) the main distance+select kernel and (b) the unit-merge kernel.  Recording does not
 * synchronise, so a
 * measurement harness can leave it on during a timed region and read the per-call durations afterwards:
 * knn_profile_enable(1) resets the call counter, knn_profile_count() = calls recorded since, knn_profile_read(i)
 * synchronises on call i's last event and returns its durations (the last 64 calls are kept; host pointers,
 * each nullable); knn_profile_last = (a), (b) of the most recent call. */
/* Kernels this library has launched in this process so far (every launch site counts itself): the measurement
 * harness reports the difference over its timed region as "gpu_launches". */
KNN_API long long knn_launch_count(void);
KNN_API int knn_profile_enable(int on);
KNN_API int knn_profile_count(void);
KNN_API int knn_profile_read(int call, float* seed_ms_host, float* distance_ms_host, float* merge_ms_host);
KNN_API int knn_profile_last(float* distance_ms_host, float* merge_ms_host);
/* Diagnostics, off unless the process was started with KNN_PAIR_STATS=1: the tcgen05 kernels then add the cycles
 * their TMA / MMA / selection roles spend waiting on each other to 32 device counters (layout: DESIGN.md
 * "stall counters").  Synchronises the device, copies the counters to out32 (host), optionally clears them. */
KNN_API int knn_debug_stats(unsigned long long* out32_host, int reset);
/* Work decomposition knn_search / knn_search_workspace use for this problem (host arithmetic only, no device call):
 * out8_host = {query blocks of 128 rows, gallery splits, candidate lists per (split, row), gallery rows per split,
 * list capacity, units per query block of the threshold-seeding pre-pass (0 = none), gallery rows per seeding unit,
 * chunks per appended maximum of the pre-pass (0 = the pre-pass selects like the main pass)}.  For measurement
 * harnesses and tests; no reference counterpart (torch.mm + topk has no decomposition, test.py:1006,44). */
KNN_API int knn_search_geometry(int64_t nq, int64_t ng, int d, int dtype, int k, int64_t* out8_host);

/* Hamming distance over binary codes + fused top-k: the replacement of
 * `(q[:, None, :] != g[None, :, :]).sum(dim=2).float()` + `argsort(dim=1)`, test_ath.py:80-100, train_ath.py:162-175.
 * knn_pack_bits: x [n, bits] fp32 (a position is set iff the value is non-zero) -> out [n, ceil(bits/64)] uint64.
 * knn_search_hamming: q/g packed words [rows, words] (words in 1, 2, 3, 4, 8); out_val = distance as fp32 ascending,
 * out_idx = gallery row (+index_base), ties by ascending gallery row; KNN_SELF_KEEP / KNN_SELF_EXCLUDE. */
KNN_API int knn_pack_bits(const void* x, int64_t n, int bits, int in_dtype, void* out_words, void* stream);
/* Packed codes -> rows of +1 / -1 in bf16, [n, dpad] with dpad = bits rounded up to 8 (padding columns 0): the
 * operand form of the TENSOR-CORE Hamming search -- <q, g> = bits - 2 * hamming(q, g) exactly (integers far below
 * 2^24), so knn_search(KNN_BF16, KNN_IP) over these rows ranks like knn_search_hamming, ties included, and
 * distance = (bits - score) / 2.  For batches of >= 32 queries the popcount kernel is issue-bound; the MMA form is
 * ~8x faster at 16x the gallery bytes. */
KNN_API int knn_unpack_bits_pm1(const void* words, int64_t n, int bits, void* out_bf16, void* stream);
/* out[i] = (bits - score[i]) / 2 (the Hamming distance behind a +-1 inner product; -inf = empty slot -> +inf). */
KNN_API int knn_hamming_from_scores(const float* score, int64_t n, int bits, float* out, void* stream);
KNN_API int knn_search_hamming(const void* q_words, const void* g_words, int64_t nq, int64_t ng, int words, int k,
                       int self_mode, int64_t self_offset, int64_t index_base,
                       float* out_val, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);
KNN_API size_t knn_search_hamming_workspace(int64_t nq, int64_t ng, int k);

/* Per-query statistics of the scores against the WHOLE gallery without materialising them: out [nq,4] double =
 * sum, sum of squares, min, max of score(q, g) over g (KNN_SELF_EXCLUDE leaves the query's own row out).  Replaces
 * the row-wise mean / std / min / max of a full similarity matrix in `normalize_similarity_matrix`,
 * fusion_eval/evaluate.py:152-177.  fp32 inputs, exact-fp32 scores (same definition as knn_scores_dense). */
KNN_API int knn_score_stats(const void* q, const void* g, const float* q_sqnorm, const float* g_sqnorm,
                    int64_t nq, int64_t ng, int d, int dtype, int metric, int self_mode, int64_t self_offset,
                    double* out, void* workspace, size_t workspace_bytes, void* stream);
KNN_API size_t knn_score_stats_workspace(int64_t nq, int64_t ng);

/* Re-score the first `first_m` candidates of every query with a per-(candidate, query column) table, the text
 * re-ranking of test.py:612-621 / 769-777: out[q][j] = rn(rn(alpha * vals[q][j]) + rn(beta * table[idx[q][j]][qcol[q]]))
 * for j < first_m and idx[q][j] != self_offset + q (self_offset < 0: no self), else vals[q][j] unchanged;
 * mask_self != 0 writes -inf for the query's own entry (the `fill_diagonal_(-inf)` that follows, test.py:623).
 * The caller re-sorts the row afterwards (knn_sort_topk). */
KNN_API int knn_rescore_topk(const float* vals, const int64_t* idx, int64_t nq, int k, const float* table,
                     int64_t table_rows, int table_cols, const int64_t* qcol, float alpha, float beta,
                     int first_m, int64_t self_offset, int mask_self, float* out_vals, void* stream);
/* Region (lesion) re-ranking of the first `first_m` candidates of every query: rerank_with_specific_lesion /
 * rerank_with_adaptive_lesion, ChestMIR/chestmir_eval.py:507-650, on top-k lists.  Images carry ragged sets of
 * normalised region vectors per lesion slot, CSR over (image, slot): vectors[offsets[img*n_slots+slot] ..
 * offsets[img*n_slots+slot+1]) of dl floats.  Query q uses qvec[q] of slot qslot[q] (-1: none).  Candidate j gets
 * region = max <qvec, v> over its vectors of that slot (-1.0 if none), score = gw*base + (1-gw)*region in IEEE double;
 * the first_m candidates are stably re-ordered by (score, base) descending.  matched[q] = candidates with region >= 0
 * (0: list left in global order, as the reference's fallback; -1: the query has no vector).  out_score [nq,first_m]
 * (nullable) = the combined scores in the new order. */
KNN_API int knn_lesion_rerank(const float* cand_val, const int64_t* cand_idx, int64_t nq, int k, int first_m,
                      const float* qvec, const int32_t* qslot, const int64_t* offsets, const float* vectors,
                      int64_t n_img, int n_slots, int dl, double global_weight, int64_t* out_idx,
                      double* out_score, int32_t* matched, void* stream);
/* Order k <= 4096 (value, index) candidates per row best-first (largest != 0: descending values), ties by ascending
 * index, entries with index < 0 last; indices must be < 2^32 - 1.  The deterministic form of the `argsort` that
 * follows a re-scoring (test.py:633). */
KNN_API int knn_sort_topk(const float* vals, const int64_t* idx, int64_t nq, int k, int largest,
                  float* out_vals, int64_t* out_idx, void* stream);

/* fp32 rows [n,d] row-major -> the KNN_F32_PACKED operand layout: 128-row tiles [ceil(n/128)][dpad][128] fp32 with
 * dpad = d rounded up to 16, zeros beyond n and d (what `torch.mm(q, g.t())`, test.py:1006, would read as g.t() tile by
 * tile).  out: knn_pack_f32_bytes(n, d) bytes, 16-byte aligned.  Pack a gallery once, search it with many batches. */
KNN_API size_t knn_pack_f32_bytes(int64_t n, int d);
KNN_API int knn_pack_f32(const float* x, int64_t n, int d, float* out, void* stream);

/* Dense score block (callers that want the full `dists` matrix of test.py:1080 / fusion_eval/metrics.py:15, and the
 * per-chunk input of knn_rank_of_positives).  out [nq,ng] fp32, same score definition and self handling as knn_search
 * (KNN_SELF_EXCLUDE writes -inf for similarity, +inf for L2).  dtype KNN_F32: the exact fp32 chain (FFMA kernel);
 * KNN_BF16 / KNN_BF16X3 (d a multiple of 8 / rows of knn_split_bf16x3): the tcgen05 CTA-pair kernel with a store
 * epilogue -- bf16 inputs with fp32 accumulation, or the three-product split that reproduces the fp32 inner product to
 * ~1e-5 |q||g| (q_sqnorm / g_sqnorm = the norms of the fp32 rows for KNN_L2). */
KNN_API int knn_scores_dense(const void* q, const void* g, const float* q_sqnorm, const float* g_sqnorm,
                     int64_t nq, int64_t ng, int d, int dtype, int metric,
                     int self_mode, int64_t self_offset, float* out, void* stream);

/* Full ranking of every row of a dense score matrix: stable order (score best-first, then ascending
 * column).  Replaces torch.argsort(dists, dim=1, descending=True) (test.py:1018) and the column-wise
 * variant of test.py:1090 when called on the transposed matrix.  scores [nq,ng] fp32; ranks [nq,ng] int64.
 * largest_first=1 for similarities, 0 for distances. */
KNN_API int    knn_rank_rows(const float* scores, int64_t nq, int64_t ng, int largest_first,
                     int64_t* ranks, void* workspace, size_t workspace_bytes, void* stream);
KNN_API size_t knn_rank_rows_workspace(int64_t nq, int64_t ng);

/* k-way merge of `parts` per-shard candidate lists (after the all-gather of row-sharded search results).
 * vals [parts,nq,k], idx [parts,nq,k] (global indices, -1 = empty) -> out [nq,k]; order as knn_search. */
KNN_API int knn_merge_topk(const float* vals, const int64_t* idx, int parts, int64_t nq, int k, int metric,
                   float* out_val, int64_t* out_idx, void* stream);
/* The same merge over `parts` (<= 16) separately placed lists: val_parts_host[p] / idx_parts_host[p] are DEVICE
 * pointers to part p's [nq,k] candidates, the two arrays themselves live on the host.  The parts may be slices of an
 * all-gathered buffer (no unpacking copy) or the symmetric-memory buffers of peer GPUs mapped into this process:
 * then the kernel reads the other shards' candidates straight over NVLink -- exchange and merge in ONE kernel
 * (the caller orders it after every peer's search with a device-side barrier). */
KNN_API int knn_merge_topk_parts(const float* const* val_parts_host, const int64_t* const* idx_parts_host, int parts,
                         int64_t nq, int k, int metric, float* out_val, int64_t* out_idx, void* stream);
/* ... with the synchronisation of the exchange inside: flags_host[p] = DEVICE address of rank p's flag array
 * (int32 [parts], zero-initialised peer-mapped memory), rank = this process, epoch = the number of this search (strictly
 * increasing from 1).  A one-warp kernel first publishes `epoch` into slot [rank] of EVERY rank's array -- ordered after
 * this rank's candidates, which the preceding kernels on `stream` wrote -- and every CTA of the merge kernel waits until
 * all `parts` slots of THIS rank's array have reached it before reading the peers' lists: no separate barrier, the wait of
 * one query's merge overlaps nothing but the slowest shard.  (A buffer may be rewritten two searches later: a peer
 * publishes epoch e+1 only behind its own merge of epoch e.) */
KNN_API int knn_merge_topk_parts_sync(const float* const* val_parts_host, const int64_t* const* idx_parts_host, int parts,
                              int64_t nq, int k, int metric, int32_t* const* flags_host, int rank, int epoch,
                              float* out_val, int64_t* out_idx, void* stream);
/* The two halves of knn_merge_topk_parts_sync as separate calls, for a PIPELINED exchange (sharded.py,
 * `ShardedFlatIndex(exchange="peer", pipeline=True)`): knn_peer_publish runs on the search stream right behind the local
 * search; knn_merge_topk_parts_wait (wait for every peer's `epoch`, then merge over NVLink -- no publish) runs on a side
 * stream, so the NEXT local search neither waits for the slowest shard nor for the merge.  Buffer reuse is then the
 * caller's protocol: with the publish on the search stream and "search j starts behind this rank's own merge j-2", a
 * candidate buffer may be rewritten FOUR searches later (proof in DESIGN section 6). */
KNN_API int knn_peer_publish(int32_t* const* flags_host, int parts, int rank, int epoch, void* stream);
KNN_API int knn_merge_topk_parts_wait(const float* const* val_parts_host, const int64_t* const* idx_parts_host, int parts,
                              int64_t nq, int k, int metric, int32_t* const* flags_host, int rank, int epoch,
                              float* out_val, int64_t* out_idx, void* stream);

/* Single-label relevance of retrieved lists: rel[i,j] = (glab[idx[i,j]] == qlab[i]), 0 for idx < 0.
 * Building block of retrieval_accuracy (test.py:38-54), compute_metrics (test_ath.py:90-172),
 * _compute_single_label_retrieval_metrics (train.py:399-441), fusion_eval/metrics.py:41-94. */
KNN_API int knn_relevance_single(const int64_t* idx, int64_t nq, int k, const int64_t* qlab, const int64_t* glab,
                         int64_t ng, uint8_t* rel, int64_t* retrieved_lab, void* stream);

/* Multi-label relevance with labels packed to 64-bit masks: Jaccard (intersection / (union + 1e-8)) in
 * the reference's own arithmetic -- arith 0: fp32 tensors, threshold rounded to fp32 (train.py:462-466,
 * nih_multilabel_training.py:90-93, test.py:956-965); arith 1: Python doubles (evaluate_nih_zilliz.py:12-17)
 * -- and the "shares >= 1 label" match of test.py:1045.  rel_jaccard / rel_any: [nq,k] uint8 (nullable). */
KNN_API int knn_relevance_multilabel(const int64_t* idx, int64_t nq, int k, const uint64_t* qmask,
                             const uint64_t* gmask, int64_t ng, double jaccard_thr, int arith,
                             uint8_t* rel_jaccard, uint8_t* rel_any, void* stream);

/* Per-query ranked-list statistics from a relevance matrix rel [nq,k] at cut-off kk <= k:
 *   hits[i]      int32  number of relevant items in the first kk
 *   first[i]     int32  1-based rank of the first relevant item (0 = none)
 *   ap_topk[i]   f64    sum_{hits}(cum/rank) / hits             (test_ath.py:138-151; 0 when no hit)
 *   prec_sum[i]  f64    sum_{hits}(cum/rank)  (caller divides by the relevant count: train.py:431-433,
 *                       fusion_eval/metrics.py:78-83, test.py:974-981)
 * All in IEEE double with the reference's operation order (no FMA contraction). */
KNN_API int knn_ranked_stats(const uint8_t* rel, int64_t nq, int k, int kk,
                     int32_t* hits, int32_t* first, double* ap_topk, double* prec_sum, void* stream);

/* The same statistics at nk <= 16 cut-offs kks[t] (device int32, each clamped to k) in ONE launch -- the k-loops of
 * test_ath.py:112-170 (`for topk in topk_values`), evaluate_nih_zilliz.py:56-62, test.py:1031-1056 without a launch and
 * a read-back per cut-off.  hits / ap_topk / prec_sum: [nq,nk]; first: [nq] over the first max(kks) items. */
KNN_API int knn_ranked_stats_multi(const uint8_t* rel, int64_t nq, int k, const int32_t* kks, int nk, int32_t* hits,
                           int32_t* first, double* ap_topk, double* prec_sum, void* stream);

/* R@K straight from retrieved indices, one launch: counts[t] (device int32 [nk], zeroed by the call) = queries with a
 * label match among their first kks[t] rows -- `correct[:k].any()` summed over the batch (retrieval_accuracy,
 * test.py:38-54).  The caller multiplies by 100 / nq in fp32 as the reference does. */
KNN_API int knn_recall_counts(const int64_t* idx, int64_t nq, int k, const int64_t* qlab, const int64_t* glab, int64_t ng,
                      const int32_t* kks, int nk, int32_t* counts, void* stream);

/* Majority vote over the first kk retrieved labels (lab [nq,k] int64).
 * tie_mode 0: first label reaching the max count in rank order (collections.Counter.most_common,
 *             test.py:149-161, test_ath.py:153);  tie_mode 1: smallest label (torch.mode train_ath.py:208,
 *             np.unique+argmax evaluate_medsiglip.py:156-157). vote [nq] int64. */
KNN_API int knn_majority_vote(const int64_t* lab, int64_t nq, int k, int kk, int tie_mode, int64_t* vote,
                      void* stream);
/* ... for nk cut-offs kks[t] (device int32) in one launch: vote [nq,nk] (`for k in k_values`, test.py:186-205). */
KNN_API int knn_majority_vote_multi(const int64_t* lab, int64_t nq, int k, const int32_t* kks, int nk, int tie_mode,
                            int64_t* vote, void* stream);

/* Trapezoidal AP of compute_ap/compute_map (test.py:58-146) from a full ranking.
 * ranks [nq,ng] int64 row-major = for query i the gallery rows best-first (the reference passes the
 * transposed [ng,nq] layout; the Python wrapper transposes).  Positives of query i are all j with
 * glab[j] == qlab[i] (the query itself included when it is part of the gallery).
 * ap [nq] f64, prs [nq,nkappa] f64 (precision at kappas, test.py:137-140), npos [nq] int32. */
KNN_API int knn_map_full(const int64_t* ranks, int64_t nq, int64_t ng, const int64_t* qlab, const int64_t* glab,
                 const int32_t* kappas, int nkappa, double* ap, double* prs, int32_t* npos, void* stream);

/* sklearn.metrics.average_precision_score over a ranked list (scores descending, tied scores grouped;
 * evaluate_nih_zilliz.py:50, train.py:473-475, nih_multilabel_training.py:95).  val [nq,k] fp32 scores
 * (best first), rel [nq,k]; ap [nq] f64 (NaN when the list has no relevant item). */
KNN_API int    knn_ap_sklearn(const float* val, const uint8_t* rel, int64_t nq, int k, double* ap,
                      void* workspace, size_t workspace_bytes, void* stream);
KNN_API size_t knn_ap_sklearn_workspace(int64_t nq, int k);

/* Full-ranking metrics without the N x N ranking (SURVEY 8(b) knn_rank_of_positives).  The reference sorts whole score
 * rows -- torch.argsort(dists, dim=0) test.py:1090, np.argsort(-dists, axis=0) test.py:962, topk(N-1) train.py:409,455,
 * sklearn's internal argsort nih_multilabel_training.py:95 -- only to read off where the RELEVANT rows ended up.  For
 * every row of a dense score block scores [nq, ld_scores >= ng] (one query each) this call returns the 0-based ranks of
 * the relevant gallery rows in the full ranking (best first: largest score, or smallest with largest_first = 0; ties by
 * ascending gallery row -- the order of knn_search / knn_rank_rows), ascending, in pos_ranks [nq, ld_out] with
 * npos[q] of them valid (ld_out >= the largest npos; ng always suffices).
 *   rel_mode 0: single label, relevant iff g_rel[g] == q_rel[q] (int64 labels; test.py:119, train.py:412)
 *            1: multi-label masks (uint64), Jaccard |a&b| / (|a|b| + 1e-8) > jaccard_thr in fp32 tensor arithmetic
 *               (train.py:462-466, nih_multilabel_training.py:90-93, test.py:956-965)
 *            2: the same in Python-float arithmetic (evaluate_nih_zilliz.py:12-17);  3: shares >= 1 label (test.py:1045)
 *   drop_self != 0: gallery row self_offset + q is neither ranked nor relevant for query q (the fill_diagonal_(-inf) /
 *               masked self of test.py:1081, train.py:406,452,468); 0: it is an ordinary row (nih_multilabel_training.py:86
 *               keeps it at score -1 and relevant).
 *   q_group [nq], g_group [ng] (int64, both or neither): gallery rows whose group equals the query's are not ranked
 *               either -- fusion_eval/metrics.py:67 drops every row that shares the query's image path.
 *   nranked[q] (nullable) = rows ranked (ng minus the dropped ones).
 *   pos_ge, pos_tgroup [nq, ld_out], ngroups [nq] (all three or none): per positive the number of ranked rows whose score
 *               is >= its own (the end of its run of equal scores) and the number of DISTINCT score values above it;
 *               ngroups = distinct score values in the row -- the threshold structure of sklearn's
 *               average_precision_score, consumed by knn_ap_sklearn_from_ranks.
 * Workspace: knn_rank_of_positives_workspace(nq, ng) bytes (two 64-bit key rows per resident CTA). */
KNN_API int knn_rank_of_positives(const float* scores, int64_t ld_scores, int64_t nq, int64_t ng, int largest_first,
                          int rel_mode, const void* q_rel, const void* g_rel, double jaccard_thr,
                          int64_t self_offset, int drop_self, const int64_t* q_group, const int64_t* g_group,
                          int32_t* pos_ranks, int64_t ld_out,
                          int32_t* pos_ge, int32_t* pos_tgroup, int32_t* npos, int32_t* nranked, int32_t* ngroups,
                          void* workspace, size_t workspace_bytes, void* stream);
KNN_API size_t knn_rank_of_positives_workspace(int64_t nq, int64_t ng);

/* AP variants from the ranks of the positives (pos_ranks / npos / nranked of knn_rank_of_positives), IEEE double in the
 * reference's operation order, accumulated in rank order.  All outputs nullable.
 *   ap_trapz [nq], prs [nq,nkappa], nres [nq]: compute_ap / compute_map, test.py:58-146 (trapezoidal AP, precision at
 *       kappas with kq = min(max(pos), kappa)); self_last_positive != 0 appends the query itself as one more positive at
 *       rank nranked[q] -- test.py:119 counts the query among its positives and its -inf score ranks it last.
 *       NaN when a query has no positive.
 *   prec_sum [nq]: sum over the positives of (positives so far) / (1-based rank) -- the numerator of the rank-by-rank AP
 *       of test.py:974-981, train.py:424-433, fusion_eval/metrics.py:78-83 (the caller divides by its relevant count).
 *   first [nq]: 1-based rank of the best positive, 0 = none (R@K of train.py:436-439).
 *   hits_at [nq,nkappa]: positives with rank < kappas[t] (mP@k of fusion_eval/metrics.py:85-88). */
KNN_API int knn_ap_from_ranks(const int32_t* pos_ranks, int64_t ld, const int32_t* npos, const int32_t* nranked,
                      int64_t nq, int self_last_positive, const int32_t* kappas, int nkappa, double* ap_trapz,
                      double* prs, int32_t* nres, double* prec_sum, int32_t* first, int32_t* hits_at, void* stream);

/* sklearn.metrics.average_precision_score over the FULL ranking (train.py:473, nih_multilabel_training.py:95) from the
 * tie structure of the positives: thresholds are the runs of equal scores, AP = -sum(diff(recall) * precision[:-1]) on
 * the reversed curves, summed with numpy's pairwise summation over all ngroups thresholds (only runs that hold a positive
 * contribute a non-zero term, but every run keeps its position in the summation tree).  ap [nq] f64, NaN = no positive.
 * Workspace: knn_ap_sklearn_from_ranks_workspace(nq, ld) bytes. */
KNN_API int knn_ap_sklearn_from_ranks(const int32_t* pos_ge, const int32_t* pos_tgroup, int64_t ld, const int32_t* npos,
                              const int32_t* ngroups, int64_t nq, double* ap, void* workspace, size_t workspace_bytes,
                              void* stream);
KNN_API size_t knn_ap_sklearn_from_ranks_workspace(int64_t nq, int64_t ld);

/* Training-time pairwise operations over a batch, forward values only.
 * knn_triplet_mine: dist [n,n] = the dense pairwise distance matrix of the batch (knn_scores_dense, KNN_L2 -- what
 *   torch.cdist(embeddings, embeddings) returns in loss.py:61,87), labels [n] int64.
 *     hard [n] fp32 (nullable): max((hardest positive - hardest negative) + margin, 0) per anchor with the reference's masks
 *       and fp32 operation order (batch_hard_triplet_loss, loss.py:60-83; the loss is their mean);
 *     all_sum [n] f64, all_pos [n], all_valid [n] int64 (all three or none): per anchor the sum of the positive triplet
 *       terms max((d(a,p) - d(a,n)) + margin, 0), how many exceed 1e-16 and how many triplets are valid
 *       (batch_all_triplet_loss, loss.py:86-112: loss = sum / (positives + 1e-16), fraction = positives / (valid + 1e-16)). */
KNN_API int knn_triplet_mine(const float* dist, const int64_t* labels, int64_t n, float margin, float* hard,
                     double* all_sum, long long* all_pos, long long* all_valid, void* stream);
/* out [nq,ng] fp32 = |a & b| / (|a| + |b| - |a & b| + eps) over 64-bit multi-hot label masks, fp32 tensor arithmetic:
 * JaccardSupConLoss.compute_jaccard_sim (loss.py:237-242), WeightedMultiLabelTripletLoss.compute_jaccard_sim
 * (loss.py:158-173). */
KNN_API int knn_jaccard_matrix(const uint64_t* q_masks, const uint64_t* g_masks, int64_t nq, int64_t ng, float eps,
                       float* out, void* stream);
/* Nearest-centroid anomaly score (anomaly/test_anomaly.py:31-48).  knn_class_means: means [nclasses,d] fp32 = numpy's
 * `x[labels == c].mean(axis=0)` for every c in classes (rows added in index order, fp32), counts [nclasses] (nullable).
 * knn_centroid_min_dist: out [n] f64 = scipy `cdist(x, centroids).min(axis=1)` (float64, direct form). */
KNN_API int knn_class_means(const float* x, const int64_t* labels, int64_t n, int d, const int64_t* classes, int nclasses,
                    float* means, int64_t* counts, void* stream);
KNN_API int knn_centroid_min_dist(const float* x, const float* centroids, int64_t n, int d, int ncentroids, double* out,
                          void* stream);

/* 1-based rank of the best-scoring gallery row that carries the query's label over a dense score row (0 = none), ranks
 * as knn_rank_rows orders them -- everything R@K reads off `output.topk(maxk)` + `target[pred] == target`
 * (retrieval_accuracy, test.py:38-54; train.py:560-577) in one launch: R@K = mean(0 < first <= K).
 * scores [nq, ld_scores >= ng] fp32, q_labels [nq], g_labels [ng] int64, first [nq] int32. */
KNN_API int knn_first_relevant_rank(const float* scores, int64_t ld_scores, int64_t nq, int64_t ng, int largest_first,
                            const int64_t* q_labels, const int64_t* g_labels, int32_t* first, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200KNN_H_ */
