// C-ABI entry points: argument validation, work decomposition, workspace carving, launches.
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "kernels.h"

namespace knn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Optional per-thread profiling: CUDA events recorded on the caller's stream right around the distance kernel
// and the unit-merge kernel of knn_search (bench.py derives the roofline fraction of the dominant kernel from
// these; nothing is recorded and nothing synchronises unless the caller opted in).
constexpr int kProfSlots = 64;
struct Profile {
  bool on = false;
  cudaEvent_t ev[kProfSlots][4] = {};  // per recorded call: start, after seeding, after the main kernel, after the merge
  int calls = 0;                        // knn_search calls recorded since knn_profile_enable(1)
};
static thread_local Profile g_prof;

static int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    return 148;  // B200
  }
  return sms;
}

// Which bf16 kernel serves a problem.
//   one 128-row query block (HBM-bound streaming), d <= 768 : the TMEM-resident-query kernel (search_ts.cu) -- every
//       byte that crosses L2 -> SM is a gallery byte;
//   one query block, d > 768                                 : the one-CTA shared-memory-A kernel (search_tc.cu);
//   several query blocks (tensor-bound)                      : the CTA-pair kernel (search_tc2.cu).  Its N = 256 tiles
//       beat the N = 128 tiles the TMEM-resident layout leaves room for (measured).
// KNN_BF16_TS=0 / 1 forces the TMEM-resident kernel off / on wherever it fits.
static bool use_ts(int dtype, int d, int qblocks) {
  static const int forced = [] {
    const char* e = getenv("KNN_BF16_TS");
    return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }();
  if (dtype != KNN_BF16 || ts_tile_cols(d) == 0) return false;
  if (forced >= 0) return forced == 1;
  return qblocks == 1;
}

// Work decomposition: a unit = (128-query block, gallery split).  Splits are contiguous gallery ranges,
// multiples of the column tile.  Query block is the fast grid index so concurrently resident CTAs walk the SAME
// gallery range (L2 reuse of the gallery stream).  Units all take the same time and CTAs are scheduled in order,
// so the number of splits is chosen to make qblocks * splits fill WHOLE waves of the machine: between one wave
// and ~4 waves' worth of splits, the count with the best (work / waves) efficiency wins (8192 queries on 148 SMs:
// 37 splits = 16 full waves instead of 10 splits = 4.32 waves).
static void seed_geometry(SearchGeom& g, int64_t ng, int tile, int64_t slots, bool pair, bool ts);

static SearchGeom make_geom(int64_t nq, int64_t ng, int d, int dtype, int k) {
  SearchGeom g;
  g.kp = kpad_for(k);
  g.L = 2 * g.kp;
  g.qblocks = (int)((nq + kRowsPerUnit - 1) / kRowsPerUnit);
  if (g.qblocks < 1) g.qblocks = 1;
  const bool ts = use_ts(dtype, d, g.qblocks);
  // every tcgen05 kernel runs two selection threads (two candidate lists) per query row
  g.groups = dtype == KNN_BF16 ? 2 : 1;
  // bf16: more than one 128-row block -> CTA pairs own 256 query rows (cta_group::2 kernels); a single block
  // (small-batch, HBM-bound regime) runs one CTA per unit: every SM streams its own gallery tiles
  if (dtype == KNN_BF16 && g.qblocks > 1 && (g.qblocks & 1)) g.qblocks += 1;
  const int tile = dtype == KNN_BF16 ? (ts ? ts_tile_cols(d) : bf16_tile_cols()) : 128;
  const int64_t ntiles = (ng + tile - 1) / tile;
  const int64_t slots = (int64_t)sm_count() * (dtype == KNN_BF16 ? 1 : 2);  // co-resident CTAs
  static const int env_waves = [] {
    const char* e = getenv("KNN_WAVES");
    return e ? atoi(e) : 0;
  }();
  // Lower end: one full wave, and units short enough (<= kMaxUnitTiles tiles) that the CTAs streaming the same
  // gallery range stay within an L2-sized window of each other -- they start together but nothing re-synchronises
  // them, and over longer units their random drift turned 2/3 of the L2 hits into HBM re-reads (measured).
  // Upper end: ~4 waves' worth for many query blocks, ~2 waves for a single block (every extra split adds a
  // candidate list per query to the final merge), at least twice the lower end.
  static const int64_t kMaxUnitTiles = [] {
    const char* e = getenv("KNN_MAX_UNIT_TILES");   // experiment knob
    return e ? (int64_t)atoll(e) : (int64_t)6144;
  }();
  const int waves = env_waves > 0 ? env_waves : (g.qblocks <= 2 ? 2 : 4);
  int64_t lo = (slots + g.qblocks - 1) / g.qblocks;
  const int64_t lo_len = (ntiles + kMaxUnitTiles - 1) / kMaxUnitTiles;
  if (lo < lo_len) lo = lo_len;
  int64_t hi = ((int64_t)waves * slots + g.qblocks - 1) / g.qblocks;
  if (hi < 2 * lo_len) hi = 2 * lo_len;
  const int64_t min_tiles = (32 * 256) / tile;                          // >= 8192 gallery rows per unit
  int64_t by_len = ntiles / min_tiles;
  if (by_len < (slots + g.qblocks - 1) / g.qblocks) by_len = (slots + g.qblocks - 1) / g.qblocks;
  if (hi > by_len) hi = by_len;
  if (hi > ntiles) hi = ntiles;
  if (hi > 2048) hi = 2048;
  if (hi < 1) hi = 1;
  if (lo > hi) lo = hi;
  int64_t best = hi;
  double best_eff = -1.0;
  for (int64_t s = hi; s >= lo && ntiles > 0; --s) {  // ties -> more splits (smaller tail in absolute time)
    const int64_t tps = (ntiles + s - 1) / s;
    const int64_t real = (ntiles + tps - 1) / tps;   // splits actually used
    const int64_t ctas = real * g.qblocks;
    const int64_t nw = (ctas + slots - 1) / slots;
    // time ~ waves * tiles-per-split; work ~ ntiles * qblocks / slots
    const double eff = (double)ntiles * g.qblocks / ((double)nw * slots * tps);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  const int64_t tiles_per_split = ntiles > 0 ? (ntiles + best - 1) / best : 1;
  g.split_len = tiles_per_split * tile;
  g.splits = ntiles > 0 ? (int)((ntiles + tiles_per_split - 1) / tiles_per_split) : 0;
  seed_geometry(g, ng, tile, slots, dtype == KNN_BF16 && !ts && g.qblocks > 1, ts);
  return g;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Gallery splits of the dense / statistics modes of the FFMA kernel (units = 128-row query block x split, all equally
// long, `slots` of them resident): the count that wastes the least of the last wave -- 10 query blocks on 296 slots:
// 29 splits = 290 CTAs = 0.98 of ONE wave, where "2 waves' worth" (30 splits = 300 CTAs) ran a second wave for 4 CTAs.
static void dense_splits(int64_t qblocks, int64_t ntiles, int64_t slots, int64_t* tiles_per_split, int* splits) {
  int64_t hi = (4 * slots + qblocks - 1) / qblocks;   // up to ~4 waves
  if (hi > ntiles) hi = ntiles;
  if (hi < 1) hi = 1;
  int64_t best = 1;
  double best_eff = -1.0;
  for (int64_t s = 1; s <= hi; ++s) {
    const int64_t tps = (ntiles + s - 1) / s;
    const int64_t real = (ntiles + tps - 1) / tps;
    const int64_t waves = (real * qblocks + slots - 1) / slots;
    double eff = (double)ntiles * (double)qblocks / ((double)waves * (double)slots * (double)tps);
    if (real * qblocks < slots) eff *= 0.999;   // prefer filling the machine among (near) ties
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  *tiles_per_split = (ntiles + best - 1) / best;
  *splits = (int)((ntiles + *tiles_per_split - 1) / *tiles_per_split);
}

// fp32 problems whose 128 x 128 tiling would occupy less than half of the SMs (and whose score rows fit the shared-memory
// sort): dense scores on 32 x 32 tiles + one sort per row (csrc/small.cu).  Same scores, same order, same bits.
static bool small_problem(int64_t nq, int64_t ng, int dtype) {
  if (dtype != KNN_F32 || nq <= 0 || ng <= 0 || ng > 4096) return false;
  static const int forced = [] {
    const char* e = getenv("KNN_SMALL_PATH");   // 0 / 1: off / on wherever it fits (tests compare the two paths)
    return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }();
  if (forced >= 0) return forced == 1 && nq * ng <= (1ll << 24);
  const int64_t tiles = ((nq + 127) / 128) * ((ng + 127) / 128);
  return tiles * 2 < sm_count();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows, d] bf16 row-major -> 2-D tensor map with a {64 elements, box_rows} box and 128-byte swizzle (the layout
// the UMMA shared-memory descriptors of ptx.cuh expect).  Rows past the end read as zeros.
int make_tmap_bf16_rows(CUtensorMap* map, const void* base, int64_t rows, int d, int box_rows, int64_t pitch,
                        int box_cols) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    KNN_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return KNN_E_CUDA;
    }
    enc = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)(pitch > 0 ? pitch : d) * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d=%d)", (int)r, (long long)rows, d);
    return KNN_E_CUDA;
  }
  return KNN_OK;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

unsigned long long* debug_stats_buffer() {
  static unsigned long long* buf = [] {
    const char* e = getenv("KNN_PAIR_STATS");
    unsigned long long* b = nullptr;
    if (e && e[0] == '1' && cudaMalloc(&b, 32 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(b, 0, 32 * sizeof(unsigned long long));
    return b;
  }();
  return buf;
}

// Threshold-seeding pre-pass: a sample of the gallery's first rows is searched by seed_splits units per query block
// and merged; the sample's exact k-th best score becomes every row's starting threshold (a valid lower bound of the
// final k-th best), so the main pass starts with a pass rate of ~k / sample instead of accepting everything.
//   * sample size: <= ~0.5 % of the main pass's work, 4096 .. 65536 rows (KNN_SEED_ROWS overrides; 0 = no pre-pass);
//   * spread over ~2 waves of CTAs (a single query block gets 2 * #SM units), but never fewer than kSeedUnitMin rows
//     per unit: a unit that short fills its candidate lists without a single compaction.  With one query block the
//     sample therefore grows to 2 * #SM * 224 = 66 k rows -- searched in parallel it costs about one tile time.
constexpr int64_t kTwoPhaseMaxRowsSeed = 1024;   // = kTwoPhaseMaxRows: small batches keep seeding from list maxima
constexpr int64_t kSeedUnitMin = 224;  // = list capacity (256 for k <= 128) - one 32-column chunk
static void seed_geometry(SearchGeom& g, int64_t ng, int tile, int64_t slots, bool pair, bool ts) {
  static const int64_t forced = [] {
    const char* e = getenv("KNN_SEED_ROWS");
    return e ? (int64_t)atoll(e) : (int64_t)-1;
  }();
  static const int maxima_mode = [] {
    const char* e = getenv("KNN_SEED_MAXIMA");   // 0: the pair kernel's pre-pass selects like the main pass (round-1 form)
    return e ? atoi(e) : 1;
  }();
  g.seed_splits = 0;
  g.seed_len = 0;
  g.seed_stride = 0;
  // the sample's threshold is shared by every list of a row: it pays for itself only when there are many of them
  // (8192 x 50M: 74 lists per row; 25000 x 112k: 6 -- there the pre-pass cost half as much as the main pass)
  if (forced < 0 && g.splits * g.groups < 8) return;
  int64_t rows = forced;
  if (rows < 0) {
    const double budget = 0.005 * (double)ng * (double)g.qblocks / (double)slots;
    rows = 4096;
    while (rows * 2 <= 65536 && (double)(rows * 2) <= budget) rows *= 2;
  }
  if (rows <= 0) return;
  // MAXIMA mode (CTA-pair kernel, large query batches; select.cuh: seed_tile_tmem): the pre-pass appends the best score
  // of every `stride` chunks per selection thread and the seeding merge takes the k-th largest of those maxima.  Nothing
  // is selected, so a unit has no warm-up and the pre-pass costs its MMA time: the unit count is free to fit the
  // machine -- the (units, tiles per unit) pair with the fewest waves x (tiles + fixed cost of a wave) wins
  // (8192 queries: 2 units of 128 tiles = one wave at 86 %, where 3 units of 86 tiles ran a second wave for 44 CTAs).
  if (pair && maxima_mode && (int64_t)g.qblocks * kRowsPerUnit > kTwoPhaseMaxRowsSeed) {
    // Sample size: the pre-pass now costs rows / ng of the main pass and a tighter starting threshold saves the main pass
    // more than that up to ~1 % of the gallery (8192 x 6.25 M, 16 k / 32 k / 64 k rows: main pass 35.4 / 35.0 / 34.6 ms,
    // pre-pass 0.32 / 0.34 / 0.56 ms): 1 % of the gallery up to 65536 rows, and at least 128 * kp rows (4 * kp chunks --
    // with fewer groups than ~2k the k-th largest maximum is a loose bound) where 2 % of the gallery allows it.
    int64_t mrows = rows;
    if (forced < 0) {
      mrows = ng / 100 < 65536 ? ng / 100 : 65536;
      if (mrows < 128 * (int64_t)g.kp) mrows = 128 * (int64_t)g.kp;
    }
    if (mrows > ng / 50) mrows = ng / 50 / tile * tile;
    const int64_t stiles = (mrows + tile - 1) / tile;
    int64_t best_u = 0, best_t = 0;
    double best_cost = 0.0;
    for (int64_t u = 1; u <= 64 && u <= stiles; ++u) {
      const int64_t t = (stiles + u - 1) / u;
      if (u * t * tile > ng / 50) continue;   // the sample stays a small prefix of the gallery (see below)
      const int64_t waves = (u * g.qblocks + slots - 1) / slots;
      const double cost = (double)waves * (double)(t + 6);
      if (best_u == 0 || cost < best_cost - 1e-9) { best_u = u; best_t = t; best_cost = cost; }
    }
    if (best_u > 0) {
      const int64_t per_thread = best_t * (tile / 32) / g.groups;   // chunks a selection thread sees in a unit
      const int64_t stride = (per_thread + g.L - 1) / g.L;          // ceil(per_thread / stride) <= L keys per list
      const int64_t maxima = best_u * g.groups * ((per_thread + stride - 1) / stride);
      if (maxima >= 2 * (int64_t)g.kp) {   // enough groups for the k-th largest maximum to be a useful bound
        g.seed_splits = (int)best_u;
        g.seed_len = best_t * tile;
        g.seed_stride = (int)stride;
        return;
      }
    }
  }
  static const int64_t forced_units = [] {
    const char* e = getenv("KNN_SEED_UNITS");   // experiment knob
    return e ? (int64_t)atoll(e) : (int64_t)0;
  }();
  int64_t units = forced_units > 0 ? forced_units : (2 * slots + g.qblocks - 1) / g.qblocks;
  if (units < 1) units = 1;
  int64_t len = (rows + units - 1) / units;
  if (len < kSeedUnitMin) len = kSeedUnitMin;
  if (g.L - 32 < kSeedUnitMin && len < 4096) len = 4096;  // small k: short lists compact anyway, keep units long
  (void)tile;
  // the sample stays a small prefix of the gallery: at most 2 % of it (the minimum useful sample, 4096 rows, is
  // not worth scanning twice in a gallery of 100 k rows)
  // (maxima mode below has no list to fill: a small gallery shard keeps its unit count with shorter units instead --
  // 64 x 1.25 M, the 8-GPU shard of config 4: 296 units of 64 rows rather than 111 of 224)
  const bool ts_maxima = ts && maxima_mode && (int64_t)g.qblocks * kRowsPerUnit <= kTwoPhaseMaxRowsSeed;
  if (ts_maxima && units * len > ng / 50) {
    const int64_t fit = ng / 50 / units / 64 * 64;   // whole chunks for both selection threads of a row
    if (fit >= 64) len = fit;
  }
  while (units > 1 && units * len > ng / 50) --units;
  if (units * len > ng / 50) return;
  g.seed_splits = (int)units;
  g.seed_len = len;
  // Small batches on the TMEM-resident kernel seed from the LIST MAXIMA (ws_layout: seed_maxima; >= 2k lists per row):
  // only the best score of every (unit, selection thread) is ever read, so the unit keeps just that -- one running
  // maximum per thread, one key per list, no appends (seed_stride = "the whole unit").
  if (ts_maxima && units * g.groups >= 2 * (int64_t)g.kp && units * g.groups <= 4096)
    g.seed_stride = kSeedWholeUnit;
}

}  // namespace knn

using namespace knn;

extern "C" int knn_version(void) { return KNN_ABI_VERSION; }
extern "C" const char* knn_last_error(void) { return g_err; }

// Workspace layout: tau_global [qblocks*128] u32 | precount [qblocks*128] i32 (two-phase merge only) |
//                   counts [(splits+seed_splits)*groups][qblocks*128] i32 |
//                   lists [(splits+seed_splits)*groups][qblocks*128][L] u64 | pre [qblocks*128][pre_cap] u64 |
//                   maxima [qblocks*128][seed_splits*groups] u32.  The extra splits are the scratch of the
//                   threshold-seeding pre-pass; pre / precount / maxima exist for small query batches only.
constexpr int kTwoPhaseMaxRows = 1024;  // up to 8 query blocks: one merge CTA per row cannot fill 148 SMs
constexpr int kTwoPhaseMinLists = 16;
constexpr int kPreCap = 2048;           // keys per row the pre-filter may gather (more: the merge re-reads the lists)
struct WsLayout {
  size_t tau_bytes, precount_bytes, counts_bytes, lists_bytes, pre_bytes, maxima_bytes, progress_bytes;
  int pre_cap;        // 0: one-phase merge
  bool seed_maxima;   // threshold seeding from list maxima
  size_t total() const {
    return tau_bytes + precount_bytes + counts_bytes + lists_bytes + pre_bytes + maxima_bytes + progress_bytes;
  }
};
static WsLayout ws_layout(const SearchGeom& g, int k) {
  WsLayout w;
  const size_t rows = (size_t)g.qblocks * kRowsPerUnit;
  const size_t vsplits = (size_t)((g.splits > 0 ? g.splits : 1) + g.seed_splits) * g.groups;
  w.tau_bytes = align_up(rows * sizeof(uint32_t), 256);
  w.counts_bytes = align_up(vsplits * rows * sizeof(int32_t), 256);
  w.lists_bytes = align_up(vsplits * rows * (size_t)g.L * sizeof(uint64_t), 256);
  const bool two_phase = rows <= (size_t)kTwoPhaseMaxRows && g.splits * g.groups >= kTwoPhaseMinLists;
  w.pre_cap = two_phase ? kPreCap : 0;
  w.precount_bytes = two_phase ? align_up(rows * sizeof(int32_t), 256) : 0;
  w.pre_bytes = two_phase ? rows * (size_t)kPreCap * sizeof(uint64_t) : 0;
  // list maxima: threshold seeding (pre-pass lists) and tightening before the pre-filter (main lists) -- each needs
  // at least 2k lists per row so that the k-th largest maximum is a useful bound
  const int seed_lists = g.seed_splits * g.groups, main_lists = g.splits * g.groups;
  w.seed_maxima = rows <= (size_t)kTwoPhaseMaxRows && seed_lists >= 2 * k && seed_lists <= 4096;
  const bool main_maxima = two_phase && main_lists >= 2 * k && main_lists <= 4096;
  const int mx = (w.seed_maxima ? seed_lists : 0) > (main_maxima ? main_lists : 0) ? seed_lists
                                                                                    : (main_maxima ? main_lists : 0);
  w.maxima_bytes = mx > 0 ? align_up(rows * (size_t)mx * sizeof(uint32_t), 256) : 0;
  // soft lock-step of the CTA pairs sharing a gallery range (pair kernel: several query blocks)
  w.progress_bytes = g.qblocks > 1 ? align_up((size_t)(g.splits > 0 ? g.splits : 1) * (g.qblocks / 2 + 1) * sizeof(int32_t), 256)
                                   : 0;
  return w;
}

extern "C" size_t knn_search_workspace(int64_t nq, int64_t ng, int d, int dtype, int k) {
  if (nq <= 0 || k < 1 || k > kMaxFusedK) return 0;
  if (dtype == KNN_BF16X3 || dtype == KNN_BF16X2) dtype = KNN_BF16;  // same geometry: bf16 rows of 3 * dpad columns
  const bool packed = dtype == KNN_F32_PACKED;  // same geometry as KNN_F32; never the small-problem path
  if (packed) dtype = KNN_F32;
  const SearchGeom g = make_geom(nq, ng < 0 ? 0 : ng, d, dtype, k);
  size_t need = ws_layout(g, k).total();
  if (!packed && small_problem(nq, ng, dtype) && (size_t)nq * (size_t)ng * sizeof(float) > need)
    need = (size_t)nq * (size_t)ng * sizeof(float);   // the dense score block of the small path
  return need;
}

extern "C" int knn_search_geometry(int64_t nq, int64_t ng, int d, int dtype, int k, int64_t* out8) {
  KNN_REQUIRE(out8 != nullptr && nq > 0 && ng >= 0 && d >= 1 && k >= 1 && k <= kMaxFusedK,
              "knn_search_geometry: bad arguments nq=%lld ng=%lld d=%d k=%d", (long long)nq, (long long)ng, d, k);
  if (dtype == KNN_BF16X3 || dtype == KNN_BF16X2) dtype = KNN_BF16;
  if (dtype == KNN_F32_PACKED) dtype = KNN_F32;
  KNN_REQUIRE(dtype == KNN_F32 || dtype == KNN_BF16, "knn_search_geometry: bad dtype %d", dtype);
  const SearchGeom g = make_geom(nq, ng, d, dtype, k);
  out8[0] = g.qblocks; out8[1] = g.splits; out8[2] = g.groups; out8[3] = g.split_len; out8[4] = g.L;
  out8[5] = g.seed_splits; out8[6] = g.seed_len; out8[7] = g.seed_stride;
  return KNN_OK;
}

static int check_common(const void* q, const void* g, const float* qs, const float* gs, int64_t nq, int64_t ng,
                        int d, int dtype, int metric, int self_mode) {
  KNN_REQUIRE(nq >= 0 && ng >= 0 && d >= 1, "bad shape nq=%lld ng=%lld d=%d", (long long)nq, (long long)ng, d);
  KNN_REQUIRE(ng < 0xFFFFFFFEll, "gallery shard too large for 32-bit local rows: %lld", (long long)ng);
  KNN_REQUIRE(dtype == KNN_F32 || dtype == KNN_BF16, "bad dtype %d", dtype);
  // the tcgen05 kernels address gallery rows with int32 TMA coordinates
  KNN_REQUIRE(dtype != KNN_BF16 || ng <= 0x7FFFFFFFll, "bf16 gallery shard too large for int32 TMA row coordinates: %lld",
              (long long)ng);
  KNN_REQUIRE(metric == KNN_COSINE || metric == KNN_IP || metric == KNN_L2, "bad metric %d", metric);
  KNN_REQUIRE(self_mode >= KNN_SELF_KEEP && self_mode <= KNN_SELF_MINUS1, "bad self_mode %d", self_mode);
  KNN_REQUIRE(!(metric == KNN_L2 && self_mode == KNN_SELF_MINUS1), "KNN_SELF_MINUS1 is a similarity convention");
  if (nq > 0 && ng > 0) {
    KNN_REQUIRE(q && g, "null q/g pointer");
    KNN_REQUIRE(metric != KNN_L2 || (qs && gs), "KNN_L2 needs q_sqnorm and g_sqnorm");
  }
  return KNN_OK;
}

extern "C" int knn_search(const void* q, const void* g, const float* q_sqnorm, const float* g_sqnorm, int64_t nq,
                          int64_t ng, int d, int dtype, int k, int metric, int self_mode, int64_t self_offset,
                          int64_t index_base, float* out_val, int64_t* out_idx, void* workspace,
                          size_t workspace_bytes, void* stream) {
  const bool two = dtype == KNN_BF16X2;
  const bool split3 = dtype == KNN_BF16X3 || two;
  if (split3) {
    KNN_REQUIRE(d % 24 == 0, "KNN_BF16X3 rows are 3 parts of a multiple of 8 columns, got d=%d", d);
    dtype = KNN_BF16;
  }
  const bool packed = dtype == KNN_F32_PACKED;
  if (packed) dtype = KNN_F32;
  int rc = check_common(q, g, q_sqnorm, g_sqnorm, nq, ng, d, dtype, metric, self_mode);
  if (rc != KNN_OK) return rc;
  if (packed) KNN_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(g)) & 15) == 0,
                          "KNN_F32_PACKED rows must be 16-byte aligned");
  KNN_REQUIRE(k >= 1, "k must be >= 1, got %d", k);
  if (k > kMaxFusedK) {
    set_error("knn_search: k=%d exceeds the fused limit %d; use knn_scores_dense + knn_rank_rows", k, kMaxFusedK);
    return KNN_E_UNSUPPORTED;
  }
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(out_val && out_idx, "null output pointer");
  const size_t need = knn_search_workspace(nq, ng, d, packed ? KNN_F32_PACKED : dtype, k);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("knn_search: workspace too small (%zu < %zu)", workspace_bytes, need);
    return KNN_E_WORKSPACE;
  }
  KNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (!split3 && !packed && small_problem(nq, ng, dtype)) {
    SearchParams ps;
    memset(&ps, 0, sizeof(ps));
    ps.q = q; ps.g = g; ps.qsq = q_sqnorm; ps.gsq = g_sqnorm;
    ps.nq = nq; ps.ng = ng; ps.d = d; ps.metric = metric; ps.self_mode = self_mode;
    ps.self_offset = self_offset - index_base;
    ps.dense_out = reinterpret_cast<float*>(workspace);
    rc = launch_dense_small(ps, s);
    if (rc != KNN_OK) return rc;
    return launch_topk_dense(ps.dense_out, nq, ng, k, metric, self_mode, ps.self_offset, index_base, out_val, out_idx, s);
  }
  const SearchGeom geo = make_geom(nq, ng, d, dtype, k);
  const bool ts = use_ts(dtype, d, geo.qblocks);

  SearchParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.g = g; p.qsq = q_sqnorm; p.gsq = g_sqnorm;
  p.nq = nq; p.ng = ng; p.d = d; p.k = k; p.kp = geo.kp;
  p.metric = metric; p.self_mode = self_mode;
  p.self_offset = self_offset - index_base;
  p.split_len = geo.split_len; p.splits = geo.splits; p.qblocks = geo.qblocks; p.groups = geo.groups;
  p.split3 = split3 ? (two ? 2 : 1) : 0;
  p.f32_packed = packed ? 1 : 0;
  const WsLayout wl = ws_layout(geo, k);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
  p.tau_global = reinterpret_cast<uint32_t*>(wsb);
  p.precount = wl.pre_cap ? reinterpret_cast<int32_t*>(wsb + wl.tau_bytes) : nullptr;
  p.counts = reinterpret_cast<int32_t*>(wsb + wl.tau_bytes + wl.precount_bytes);
  p.lists = reinterpret_cast<uint64_t*>(wsb + wl.tau_bytes + wl.precount_bytes + wl.counts_bytes);
  p.pre = wl.pre_cap ? reinterpret_cast<uint64_t*>(wsb + wl.tau_bytes + wl.precount_bytes + wl.counts_bytes + wl.lists_bytes)
                     : nullptr;
  p.pre_cap = wl.pre_cap;
  p.maxima = wl.maxima_bytes ? reinterpret_cast<uint32_t*>(wsb + wl.tau_bytes + wl.precount_bytes + wl.counts_bytes +
                                                           wl.lists_bytes + wl.pre_bytes)
                             : nullptr;
  static const int lockstep = [] {
    const char* e = getenv("KNN_PAIR_LOCKSTEP");   // window in tiles (default 32), 0 = off
    return e ? atoi(e) : 32;
  }();
  p.progress = (lockstep > 0 && wl.progress_bytes && dtype == KNN_BF16)
                   ? reinterpret_cast<int32_t*>(wsb + wl.tau_bytes + wl.precount_bytes + wl.counts_bytes + wl.lists_bytes +
                                                wl.pre_bytes + wl.maxima_bytes)
                   : nullptr;
  const size_t tau_bytes = wl.tau_bytes + wl.precount_bytes;   // thresholds and pre-filter counts: zeroed together
  p.dense_out = nullptr;

  const bool prof = g_prof.on;
  cudaEvent_t* pev = g_prof.ev[g_prof.calls % kProfSlots];
  if (prof) {
    for (int i = 0; i < 4; ++i)
      if (!pev[i]) KNN_CHECK_CUDA(cudaEventCreate(&pev[i]));
  }
  if (geo.splits > 0) {
    KNN_CHECK_CUDA(cudaMemsetAsync(p.tau_global, 0, tau_bytes, s));
    if (prof) KNN_CHECK_CUDA(cudaEventRecord(pev[0], s));
    // Threshold seeding (see seed_geometry): search the sample into the scratch lists, merge in seeding mode.
    // The sample's candidates are discarded -- the main pass visits those rows again.
    if (p.progress) KNN_CHECK_CUDA(cudaMemsetAsync(p.progress, 0xFF, wl.progress_bytes, s));   // -1 = not started
    if (geo.seed_splits > 0) {
      SearchParams ps = p;
      ps.progress = nullptr;
      ps.ng = (int64_t)geo.seed_splits * geo.seed_len;
      ps.splits = geo.seed_splits;
      ps.split_len = geo.seed_len;
      // (whole-unit maxima are written to the maxima table: only with the list-maxima seeding that reads it)
      ps.seed_stride = (geo.seed_stride == kSeedWholeUnit && !(wl.seed_maxima && ts)) ? 0 : geo.seed_stride;
      const size_t scratch_rows = (size_t)geo.splits * geo.groups * geo.qblocks * kRowsPerUnit;
      ps.lists = p.lists + scratch_rows * (size_t)geo.L;
      ps.counts = p.counts + scratch_rows;
      rc = ts ? launch_search_bf16_ts(ps, s)
              : (dtype == KNN_BF16) ? launch_search_bf16(ps, s) : launch_search_f32(ps, false, s);
      if (rc != KNN_OK) return rc;
      // single query block: hundreds of short lists per row -> seed from the list maxima (list-parallel)
      rc = wl.seed_maxima ? launch_seed_from_maxima(ps, p.tau_global, s, ts && ps.seed_stride == kSeedWholeUnit)
                          : launch_merge_units(ps, 0, nullptr, nullptr, p.tau_global, s);
      if (rc != KNN_OK) return rc;
    }
    if (prof) KNN_CHECK_CUDA(cudaEventRecord(pev[1], s));
    rc = ts ? launch_search_bf16_ts(p, s)
            : (dtype == KNN_BF16) ? launch_search_bf16(p, s) : launch_search_f32(p, false, s);
    if (rc != KNN_OK) return rc;
  } else if (prof) {
    KNN_CHECK_CUDA(cudaEventRecord(pev[0], s));
    KNN_CHECK_CUDA(cudaEventRecord(pev[1], s));
  }
  if (prof) KNN_CHECK_CUDA(cudaEventRecord(pev[2], s));
  rc = launch_merge_units(p, index_base, out_val, out_idx, nullptr, s);
  if (rc != KNN_OK) return rc;
  if (prof) {
    KNN_CHECK_CUDA(cudaEventRecord(pev[3], s));
    ++g_prof.calls;
  }
  return KNN_OK;
}

extern "C" long long knn_launch_count(void) { return launches_so_far(); }

extern "C" int knn_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.calls = 0;
  return KNN_OK;
}

extern "C" int knn_profile_count(void) { return g_prof.calls; }

extern "C" int knn_profile_read(int call, float* seed_ms, float* distance_ms, float* merge_ms) {
  if (call < 0 || call >= g_prof.calls || call < g_prof.calls - kProfSlots) {
    set_error("knn_profile_read: call %d not recorded (%d calls since enable, last %d kept)", call, g_prof.calls,
              kProfSlots);
    return KNN_E_INVALID;
  }
  cudaEvent_t* pev = g_prof.ev[call % kProfSlots];
  KNN_CHECK_CUDA(cudaEventSynchronize(pev[3]));
  float sd = 0.f, a = 0.f, b = 0.f;
  KNN_CHECK_CUDA(cudaEventElapsedTime(&sd, pev[0], pev[1]));
  KNN_CHECK_CUDA(cudaEventElapsedTime(&a, pev[1], pev[2]));
  KNN_CHECK_CUDA(cudaEventElapsedTime(&b, pev[2], pev[3]));
  if (seed_ms) *seed_ms = sd;
  if (distance_ms) *distance_ms = a;
  if (merge_ms) *merge_ms = b;
  return KNN_OK;
}

extern "C" int knn_profile_last(float* distance_ms, float* merge_ms) {
  if (g_prof.calls == 0) {
    set_error("knn_profile_last: no profiled knn_search call on this thread");
    return KNN_E_INVALID;
  }
  return knn_profile_read(g_prof.calls - 1, nullptr, distance_ms, merge_ms);
}

extern "C" int knn_debug_stats(unsigned long long* out32, int reset) {
  unsigned long long* b = debug_stats_buffer();
  if (!b) {
    set_error("knn_debug_stats: diagnostics are off (set KNN_PAIR_STATS=1 before the first search)");
    return KNN_E_INVALID;
  }
  KNN_CHECK_CUDA(cudaDeviceSynchronize());
  if (out32) KNN_CHECK_CUDA(cudaMemcpy(out32, b, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) KNN_CHECK_CUDA(cudaMemset(b, 0, 32 * sizeof(unsigned long long)));
  return KNN_OK;
}

// ---------------------------------------------------------------------------------------------- Hamming
static SearchGeom hamming_geom(int64_t nq, int64_t ng, int k) {
  SearchGeom g;
  memset(&g, 0, sizeof(g));
  g.kp = kpad_for(k);
  g.L = 2 * g.kp;
  g.groups = 1;
  g.qblocks = (int)((nq + kRowsPerUnit - 1) / kRowsPerUnit);
  if (g.qblocks < 1) g.qblocks = 1;
  const int64_t nchunks = (ng + 31) / 32;
  int64_t want = ((int64_t)8 * sm_count() + g.qblocks - 1) / g.qblocks;   // 4 CTAs per SM, ~2 waves
  const int64_t by_len = nchunks / 64 > 0 ? nchunks / 64 : 1;            // >= 2048 gallery rows per unit
  if (want > by_len) want = by_len;
  if (want < 1) want = 1;
  const int64_t cps = nchunks > 0 ? (nchunks + want - 1) / want : 1;
  g.split_len = cps * 32;
  g.splits = nchunks > 0 ? (int)((nchunks + cps - 1) / cps) : 0;
  g.seed_splits = 0;
  g.seed_len = 0;
  return g;
}

extern "C" size_t knn_search_hamming_workspace(int64_t nq, int64_t ng, int k) {
  if (nq <= 0 || k < 1 || k > kMaxFusedK) return 0;
  return ws_layout(hamming_geom(nq, ng < 0 ? 0 : ng, k), k).total();
}

extern "C" int knn_pack_bits(const void* x, int64_t n, int bits, int in_dtype, void* out_words, void* stream) {
  KNN_REQUIRE(n >= 0 && bits >= 1, "knn_pack_bits: bad shape n=%lld bits=%d", (long long)n, bits);
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(x && out_words, "knn_pack_bits: null pointer");
  return launch_pack_bits(x, in_dtype, n, bits, (bits + 63) / 64, reinterpret_cast<uint64_t*>(out_words),
                          (cudaStream_t)stream);
}

extern "C" int knn_unpack_bits_pm1(const void* words, int64_t n, int bits, void* out_bf16, void* stream) {
  KNN_REQUIRE(n >= 0 && bits >= 1, "knn_unpack_bits_pm1: bad shape n=%lld bits=%d", (long long)n, bits);
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(words && out_bf16, "knn_unpack_bits_pm1: null pointer");
  KNN_REQUIRE((reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0, "out must be 16-byte aligned");
  return launch_unpack_pm1(reinterpret_cast<const uint64_t*>(words), n, bits, (bits + 63) / 64, out_bf16,
                           (cudaStream_t)stream);
}

extern "C" int knn_hamming_from_scores(const float* score, int64_t n, int bits, float* out, void* stream) {
  KNN_REQUIRE(n >= 0 && bits >= 1, "knn_hamming_from_scores: bad sizes");
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(score && out, "knn_hamming_from_scores: null pointer");
  return launch_hamming_from_scores(score, n, bits, out, (cudaStream_t)stream);
}

extern "C" int knn_search_hamming(const void* q_words, const void* g_words, int64_t nq, int64_t ng, int words, int k,
                                  int self_mode, int64_t self_offset, int64_t index_base, float* out_val,
                                  int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
  KNN_REQUIRE(nq >= 0 && ng >= 0 && words >= 1, "bad shape nq=%lld ng=%lld words=%d", (long long)nq, (long long)ng, words);
  KNN_REQUIRE(ng < 0xFFFFFFFEll, "gallery shard too large for 32-bit local rows: %lld", (long long)ng);
  KNN_REQUIRE(self_mode == KNN_SELF_KEEP || self_mode == KNN_SELF_EXCLUDE, "bad self_mode %d", self_mode);
  KNN_REQUIRE(k >= 1, "k must be >= 1, got %d", k);
  if (k > kMaxFusedK) {
    set_error("knn_search_hamming: k=%d exceeds the fused limit %d", k, kMaxFusedK);
    return KNN_E_UNSUPPORTED;
  }
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(out_val && out_idx && (ng == 0 || (q_words && g_words)), "null pointer");
  const size_t need = knn_search_hamming_workspace(nq, ng, k);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("knn_search_hamming: workspace too small (%zu < %zu)", workspace_bytes, need);
    return KNN_E_WORKSPACE;
  }
  KNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const SearchGeom geo = hamming_geom(nq, ng, k);
  SearchParams p;
  memset(&p, 0, sizeof(p));
  p.q = q_words; p.g = g_words;
  p.nq = nq; p.ng = ng; p.d = words * 64; p.k = k; p.kp = geo.kp;
  p.metric = KNN_L2;  // distances: smaller = better, emitted as positive values by the unit merge
  p.self_mode = self_mode;
  p.self_offset = self_offset - index_base;
  p.split_len = geo.split_len; p.splits = geo.splits; p.qblocks = geo.qblocks; p.groups = 1;
  const WsLayout wl = ws_layout(geo, k);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
  p.tau_global = reinterpret_cast<uint32_t*>(wsb);
  p.precount = wl.pre_cap ? reinterpret_cast<int32_t*>(wsb + wl.tau_bytes) : nullptr;
  p.counts = reinterpret_cast<int32_t*>(wsb + wl.tau_bytes + wl.precount_bytes);
  p.lists = reinterpret_cast<uint64_t*>(wsb + wl.tau_bytes + wl.precount_bytes + wl.counts_bytes);
  p.pre = wl.pre_cap ? reinterpret_cast<uint64_t*>(wsb + wl.tau_bytes + wl.precount_bytes + wl.counts_bytes + wl.lists_bytes)
                     : nullptr;
  p.pre_cap = wl.pre_cap;
  KNN_CHECK_CUDA(cudaMemsetAsync(p.tau_global, 0, wl.tau_bytes + wl.precount_bytes, s));
  if (geo.splits > 0) {
    int rc = launch_search_hamming(p, words, s);
    if (rc != KNN_OK) return rc;
  }
  return launch_merge_units(p, index_base, out_val, out_idx, nullptr, s);
}

extern "C" size_t knn_score_stats_workspace(int64_t nq, int64_t ng) {
  if (nq <= 0 || ng <= 0) return 0;
  const int64_t qblocks = (nq + kRowsPerUnit - 1) / kRowsPerUnit;
  const int64_t ntiles = (ng + 127) / 128;
  int64_t tps;
  int splits;
  dense_splits(qblocks, ntiles, (int64_t)2 * sm_count(), &tps, &splits);
  return (size_t)splits * qblocks * kRowsPerUnit * 4 * sizeof(double);
}

extern "C" int knn_score_stats(const void* q, const void* g, const float* q_sqnorm, const float* g_sqnorm, int64_t nq,
                               int64_t ng, int d, int dtype, int metric, int self_mode, int64_t self_offset,
                               double* out, void* workspace, size_t workspace_bytes, void* stream) {
  const bool packed = dtype == KNN_F32_PACKED;
  if (packed) dtype = KNN_F32;
  int rc = check_common(q, g, q_sqnorm, g_sqnorm, nq, ng, d, dtype, metric, self_mode);
  if (rc != KNN_OK) return rc;
  if (dtype != KNN_F32) {
    set_error("knn_score_stats: fp32 inputs only");
    return KNN_E_UNSUPPORTED;
  }
  if (packed) KNN_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(g)) & 15) == 0,
                          "KNN_F32_PACKED rows must be 16-byte aligned");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(out, "null output pointer");
  KNN_REQUIRE(ng > 0, "knn_score_stats: empty gallery");
  const size_t need = knn_score_stats_workspace(nq, ng);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("knn_score_stats: workspace too small (%zu < %zu)", workspace_bytes, need);
    return KNN_E_WORKSPACE;
  }
  SearchParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.g = g; p.qsq = q_sqnorm; p.gsq = g_sqnorm;
  p.nq = nq; p.ng = ng; p.d = d; p.k = 1; p.kp = 32;
  p.metric = metric; p.self_mode = self_mode; p.self_offset = self_offset;
  p.groups = 1;
  p.qblocks = (int)((nq + kRowsPerUnit - 1) / kRowsPerUnit);
  const int64_t ntiles = (ng + 127) / 128;
  int64_t tps;
  dense_splits(p.qblocks, ntiles, (int64_t)2 * sm_count(), &tps, &p.splits);
  p.split_len = tps * 128;
  p.stats_out = reinterpret_cast<double*>(workspace);
  p.f32_packed = packed ? 1 : 0;
  rc = launch_search_f32(p, false, (cudaStream_t)stream);
  if (rc != KNN_OK) return rc;
  return launch_stats_reduce(p.stats_out, p.splits, p.qblocks, nq, out, (cudaStream_t)stream);
}

extern "C" int knn_rescore_topk(const float* vals, const int64_t* idx, int64_t nq, int k, const float* table,
                                int64_t table_rows, int table_cols, const int64_t* qcol, float alpha, float beta,
                                int first_m, int64_t self_offset, int mask_self, float* out_vals, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && table_rows >= 0 && table_cols >= 1, "knn_rescore_topk: bad sizes");
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(vals && idx && table && qcol && out_vals, "knn_rescore_topk: null pointer");
  return launch_rescore_topk(vals, idx, nq, k, table, table_rows, table_cols, qcol, alpha, beta, first_m, self_offset,
                             mask_self, out_vals, (cudaStream_t)stream);
}

extern "C" int knn_sort_topk(const float* vals, const int64_t* idx, int64_t nq, int k, int largest, float* out_vals,
                             int64_t* out_idx, void* stream) {
  KNN_REQUIRE(nq >= 0 && k >= 1 && k <= 4096, "knn_sort_topk: k must be in [1, 4096], got %d", k);
  if (nq == 0) return KNN_OK;
  KNN_REQUIRE(vals && idx && out_vals && out_idx, "knn_sort_topk: null pointer");
  return launch_sort_topk(vals, idx, nq, k, largest, out_vals, out_idx, (cudaStream_t)stream);
}

extern "C" int knn_scores_dense(const void* q, const void* g, const float* q_sqnorm, const float* g_sqnorm,
                                int64_t nq, int64_t ng, int d, int dtype, int metric, int self_mode,
                                int64_t self_offset, float* out, void* stream) {
  const bool split3 = dtype == KNN_BF16X3;
  if (split3) {
    KNN_REQUIRE(d % 24 == 0, "KNN_BF16X3 rows are 3 parts of a multiple of 8 columns, got d=%d", d);
    dtype = KNN_BF16;
  }
  const bool packed = dtype == KNN_F32_PACKED;
  if (packed) dtype = KNN_F32;
  int rc = check_common(q, g, q_sqnorm, g_sqnorm, nq, ng, d, dtype, metric, self_mode);
  if (rc != KNN_OK) return rc;
  if (nq == 0 || ng == 0) return KNN_OK;
  KNN_REQUIRE(out, "null output pointer");
  if (packed) KNN_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(g)) & 15) == 0,
                          "KNN_F32_PACKED rows must be 16-byte aligned");
  if (dtype == KNN_BF16) {
    // tensor-core dense mode: the CTA-pair tcgen05 kernel with a store epilogue (bf16 rows, or bf16x3 split rows whose
    // three products reproduce the fp32 inner product to ~1e-5 |q||g|)
    KNN_REQUIRE(d % 8 == 0, "bf16 rows must be padded to a multiple of 8 columns, got d=%d", d);
    SearchParams pb;
    memset(&pb, 0, sizeof(pb));
    pb.q = q; pb.g = g; pb.qsq = q_sqnorm; pb.gsq = g_sqnorm;
    pb.nq = nq; pb.ng = ng; pb.d = d; pb.k = 1; pb.kp = 32;
    pb.metric = metric; pb.self_mode = self_mode; pb.self_offset = self_offset;
    pb.groups = 2; pb.split3 = split3 ? 1 : 0;
    pb.qblocks = (int)((nq + kRowsPerUnit - 1) / kRowsPerUnit);
    if (pb.qblocks & 1) pb.qblocks += 1;                       // CTA pairs own 256 query rows
    const int tile = bf16_tile_cols();
    const int64_t ntiles_b = (ng + tile - 1) / tile;
    int64_t tps_b;
    dense_splits(pb.qblocks, ntiles_b, sm_count(), &tps_b, &pb.splits);   // one CTA per SM
    pb.split_len = tps_b * tile;
    pb.dense_out = out;
    return launch_search_bf16_pair(pb, (cudaStream_t)stream);
  }
  SearchParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.g = g; p.qsq = q_sqnorm; p.gsq = g_sqnorm;
  p.nq = nq; p.ng = ng; p.d = d; p.k = 1; p.kp = 32;
  p.metric = metric; p.self_mode = self_mode; p.self_offset = self_offset;
  p.groups = 1;
  p.qblocks = (int)((nq + kRowsPerUnit - 1) / kRowsPerUnit);
  const int64_t ntiles = (ng + 127) / 128;
  int64_t tps;
  dense_splits(p.qblocks, ntiles, (int64_t)2 * sm_count(), &tps, &p.splits);   // two resident CTAs per SM
  p.split_len = tps * 128;
  p.dense_out = out;
  p.f32_packed = packed ? 1 : 0;
  if (!packed && ((nq + 127) / 128) * ((ng + 127) / 128) * 2 < sm_count())
    return launch_dense_small(p, (cudaStream_t)stream);
  return launch_search_f32(p, true, (cudaStream_t)stream);
}

extern "C" size_t knn_pack_f32_bytes(int64_t n, int d) {
  if (n <= 0 || d < 1) return 0;
  return pack_f32_bytes(n, d);
}

extern "C" int knn_pack_f32(const float* x, int64_t n, int d, float* out, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1, "knn_pack_f32: bad shape n=%lld d=%d", (long long)n, d);
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(x && out, "knn_pack_f32: null pointer");
  KNN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "knn_pack_f32: out must be 16-byte aligned");
  return launch_pack_f32(x, n, d, out, (cudaStream_t)stream);
}
