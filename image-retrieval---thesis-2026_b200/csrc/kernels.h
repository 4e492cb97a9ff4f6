// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace knn {

struct SearchParams {
  const void* q;
  const void* g;
  const float* qsq;   // |q|^2 (L2 only)
  const float* gsq;   // |g|^2 (L2 only)
  int64_t nq, ng;
  int d;
  int k, kp;
  int metric;
  int self_mode;
  int64_t self_offset;  // LOCAL gallery row of query 0 (query i is local row self_offset + i)
  int64_t split_len;    // gallery rows per split
  int splits;
  int groups;            // candidate lists per (split, row): 2 for the TMEM-resident kernel, else 1
  int qblocks;
  int split3;            // rows are knn_split_bf16x3 output (3 parts of d/3 columns): q [hi|lo|hi], g [hi|hi|lo];
                         // 2 = the same rows, TWO products only (q_hi.g_hi + q_lo.g_hi: KNN_BF16X2)
  int f32_packed;        // fp32 rows are knn_pack_f32 output (128-row tiles [tile][dpad][128]); d is the original width
  uint64_t* lists;       // [splits * groups][qblocks][128][2*kp] candidate keys (unordered)
  int32_t* counts;       // [splits * groups][qblocks][128] keys in each list when its unit finished
  uint32_t* tau_global;  // [qblocks*128] shared thresholds (order-preserving encoding, 0 = none)
  // small query batches (one CTA per row would leave the GPU idle): scratch of the two-phase unit merge, nullptr = off
  uint64_t* pre;         // [qblocks*128][pre_cap] keys >= tau_global gathered by a list-parallel pre-filter
  int32_t* precount;     // [qblocks*128] keys the pre-filter found (may exceed pre_cap: the merge then re-reads the lists)
  int pre_cap;
  uint32_t* maxima;      // [qblocks*128][splits*groups] best score of every list (threshold seeding from list maxima)
  int32_t* progress;     // [splits][qblocks/2] tile index every CTA pair last published (soft lock-step of the pairs that
                         // stream the same gallery range, pair kernel only); -1 = not started; nullptr = off
  int seed_stride;       // > 0: threshold-seeding pass in MAXIMA mode: every selection thread appends the best score of
                         // each run of `seed_stride` of its 32-column chunks instead of selecting; kSeedWholeUnit
                         // (TMEM-resident kernel): one maximum per unit, written to `maxima` directly
  float* dense_out;      // dense mode only
  double* stats_out;     // statistics mode only: [splits][qblocks][128][4] = sum, sum of squares, min, max
};

int launch_search_f32(const SearchParams& p, bool dense, cudaStream_t stream);
// fp32 rows [n, d] -> 128-row tiles [ceil(n/128)][dpad][128] (dpad = d rounded up to 16, zero padded): the operand
// layout of the bulk-copy pipeline of the FFMA kernel (KNN_F32_PACKED)
size_t pack_f32_bytes(int64_t n, int d);
int launch_pack_f32(const float* x, int64_t n, int d, float* out, cudaStream_t stream);
int launch_search_bf16(const SearchParams& p, cudaStream_t stream);       // dispatch (pair kernel by default)
int launch_search_bf16_pair(const SearchParams& p, cudaStream_t stream);  // cta_group::2, csrc/search_tc2.cu
int bf16_tile_cols();  // gallery rows per tile of the shared-memory-A tcgen05 kernels
int launch_search_bf16_ts(const SearchParams& p, cudaStream_t stream);    // query tile in TMEM, csrc/search_ts.cu
int ts_tile_cols(int d);  // gallery rows per tile of the TMEM-resident kernel, 0 = d does not fit

// out[q][0..3] = sum, sum of squares, min, max over the splits' partials (splits summed in ascending order)
int launch_stats_reduce(const double* partials, int splits, int qblocks, int64_t nq, double* out, cudaStream_t stream);
// vals[q][j] <- rn(rn(alpha*vals[q][j]) + rn(beta*table[idx[q][j]][qcol[q]])) for j < first_m, idx != self (test.py:612-621)
int launch_rescore_topk(const float* vals, const int64_t* idx, int64_t nq, int k, const float* table, int64_t table_rows,
                        int table_cols, const int64_t* qcol, float alpha, float beta, int first_m, int64_t self_offset,
                        int mask_self, float* out_vals, cudaStream_t stream);
// threshold seeding from list maxima: tau_out[row] = k-th largest of the row's list maxima (needs >= k lists per row)
// reduced = true: the search kernel already wrote p.maxima (seed_stride == kSeedWholeUnit)
int launch_seed_from_maxima(const SearchParams& p, uint32_t* tau_out, cudaStream_t stream, bool reduced = false);
constexpr int kSeedWholeUnit = 1 << 30;
// order k (value, index) candidates per row best-first, ties by ascending index
int launch_sort_topk(const float* vals, const int64_t* idx, int64_t nq, int k, int largest, float* out_vals,
                     int64_t* out_idx, cudaStream_t stream);

// Small problems (csrc/small.cu): dense scores on 32 x 32 tiles (same bits as launch_search_f32's dense mode) and the
// best k of every dense row (<= 4096 columns) in knn_search's output format
int launch_dense_small(const SearchParams& p, cudaStream_t stream);
int launch_topk_dense(const float* dense, int64_t nq, int64_t ng, int k, int metric, int self_mode, int64_t self_offset,
                      int64_t index_base, float* out_val, int64_t* out_idx, cudaStream_t stream);

// Hamming distance over packed 64-bit code words (csrc/search_hamming.cu); p.q / p.g point at uint64 [rows, words]
int launch_search_hamming(const SearchParams& p, int words, cudaStream_t stream);
int launch_hamming_from_scores(const float* score, int64_t n, int bits, float* out, cudaStream_t stream);
int launch_unpack_pm1(const uint64_t* words, int64_t n, int bits, int nwords, void* out_bf16, cudaStream_t stream);
int launch_pack_bits(const void* x, int dtype, int64_t n, int bits, int words, uint64_t* out, cudaStream_t stream);

// [rows, d] bf16 row-major -> tensor map with a {64, box_rows} box, 128-byte swizzle (api.cu)
// pitch = elements between consecutive rows (0: d) -- a map over a column range of wider rows reads zeros past d
int make_tmap_bf16_rows(CUtensorMap* map, const void* base, int64_t rows, int d, int box_rows, int64_t pitch = 0,
                        int box_cols = 64);   // box_cols 64 -> 128-byte swizzle, 32 -> 64-byte swizzle

// Diagnostics (KNN_PAIR_STATS=1): 32 device counters the tcgen05 kernels add stall cycles to; nullptr when off.
unsigned long long* debug_stats_buffer();

// tau_out != nullptr: seeding mode (publish each row's k-th best score instead of writing results)
int launch_merge_units(const SearchParams& p, int64_t index_base, float* out_val, int64_t* out_idx,
                       uint32_t* tau_out, cudaStream_t stream);

}  // namespace knn
