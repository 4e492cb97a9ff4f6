// Training-time pairwise operations over a batch (SURVEY 8(f)-4 tail), forward values only.
//   * triplet mining on the pairwise distance matrix: batch-hard and batch-all (loss.py:60-133)
//   * Jaccard similarity matrix of multi-hot label rows (JaccardSupConLoss.compute_jaccard_sim, loss.py:237-242;
//     WeightedMultiLabelTripletLoss.compute_jaccard_sim, loss.py:158-173)
//   * nearest-centroid anomaly score (anomaly/test_anomaly.py:31-48): class means + distance to the nearest one
// The distance matrix itself comes from knn_scores_dense (KNN_L2): the same fused distance kernel as the search.
#include "common.cuh"

namespace knn {
namespace {

// One CTA per anchor i over the dense distance row dist[i, :] (fp32, as torch.cdist returns).
//   batch-hard: hardest positive = max_j mask_ap(i,j) * d(i,j), hardest negative = min_k d(i,k) + rowmax * (1 - mask_an(i,k)),
//               hard[i] = max((hp - hn) + margin, 0) in fp32 -- the reference's operations in its order (loss.py:61-83)
//   batch-all : over the valid triplets (j != i same label, k different label): sum of max((d(i,j) - d(i,k)) + margin, 0)
//               (fp32 terms, double accumulation), number of terms > 1e-16, number of valid triplets (loss.py:90-112)
__global__ void __launch_bounds__(256) triplet_mine_kernel(const float* __restrict__ dist, const int64_t* __restrict__ labels,
                                                          int64_t n, float margin, float* __restrict__ hard,
                                                          double* __restrict__ all_sum, long long* __restrict__ all_pos,
                                                          long long* __restrict__ all_valid) {
  __shared__ float red_f[3][8];
  __shared__ double red_d[8];
  __shared__ long long red_l[2][8];
  const int64_t i = blockIdx.x;
  const float* row = dist + i * n;
  const int64_t li = labels[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float hp = 0.0f, rmax = -INFINITY;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const float d = row[j];
    rmax = fmaxf(rmax, d);
    if (j != i && labels[j] == li) hp = fmaxf(hp, d);   // mask * d: every other entry contributes 0
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    hp = fmaxf(hp, __shfl_xor_sync(0xFFFFFFFFu, hp, off));
    rmax = fmaxf(rmax, __shfl_xor_sync(0xFFFFFFFFu, rmax, off));
  }
  if (lane == 0) { red_f[0][warp] = hp; red_f[1][warp] = rmax; }
  __syncthreads();
  hp = red_f[0][0]; rmax = red_f[1][0];
  for (int w = 1; w < 8; ++w) { hp = fmaxf(hp, red_f[0][w]); rmax = fmaxf(rmax, red_f[1][w]); }
  float hn = INFINITY;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
    const float d = row[k];
    hn = fminf(hn, labels[k] != li ? d : __fadd_rn(d, rmax));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) hn = fminf(hn, __shfl_xor_sync(0xFFFFFFFFu, hn, off));
  if (lane == 0) red_f[2][warp] = hn;
  __syncthreads();
  hn = red_f[2][0];
  for (int w = 1; w < 8; ++w) hn = fminf(hn, red_f[2][w]);
  if (threadIdx.x == 0 && hard) hard[i] = fmaxf(__fadd_rn(__fsub_rn(hp, hn), margin), 0.0f);
  if (all_sum == nullptr) return;
  double s = 0.0;
  long long npos = 0, nvalid = 0;
  for (int64_t j = 0; j < n; ++j) {
    if (j == i || labels[j] != li) continue;
    const float dij = row[j];
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      if (labels[k] == li) continue;   // k != i and k != j follow (different label)
      const float t = __fadd_rn(__fsub_rn(dij, row[k]), margin);
      ++nvalid;
      if (t > 0.0f) {
        s += (double)t;
        npos += t > 1e-16f;
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
    npos += __shfl_xor_sync(0xFFFFFFFFu, npos, off);
    nvalid += __shfl_xor_sync(0xFFFFFFFFu, nvalid, off);
  }
  __syncthreads();
  if (lane == 0) { red_d[warp] = s; red_l[0][warp] = npos; red_l[1][warp] = nvalid; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { s += red_d[w]; npos += red_l[0][w]; nvalid += red_l[1][w]; }
    all_sum[i] = s; all_pos[i] = npos; all_valid[i] = nvalid;
  }
}

// intersection / (|a| + |b| - intersection + eps) in fp32 tensor arithmetic over 64-bit label masks
__global__ void jaccard_matrix_kernel(const uint64_t* __restrict__ qm, const uint64_t* __restrict__ gm, int64_t nq,
                                      int64_t ng, float eps, float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * ng) return;
  const uint64_t a = qm[t / ng], b = gm[t % ng];
  const float inter = (float)__popcll(a & b);
  const float uni = __fsub_rn(__fadd_rn((float)__popcll(a), (float)__popcll(b)), inter);
  out[t] = __fdiv_rn(inter, __fadd_rn(uni, eps));
}

// mean of the rows of class c: numpy's `x[labels == c].mean(axis=0)` on float32 rows -- the members are added row by row in
// index order (fp32), then divided by their number.  One thread per (class, column).
__global__ void class_means_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, int64_t n, int d,
                                   const int64_t* __restrict__ classes, int nclasses, float* __restrict__ means,
                                   int64_t* __restrict__ counts) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nclasses * d) return;
  const int c = (int)(t / d), col = (int)(t % d);
  const int64_t cls = classes[c];
  float acc = 0.0f;
  int64_t cnt = 0;
  for (int64_t r = 0; r < n; ++r)
    if (labels[r] == cls) {
      acc = cnt == 0 ? x[r * d + col] : __fadd_rn(acc, x[r * d + col]);
      ++cnt;
    }
  means[t] = __fdiv_rn(acc, (float)cnt);   // 0 / 0 = NaN for an empty class, as numpy
  if (col == 0 && counts) counts[c] = cnt;
}

// scipy.spatial.distance.cdist(x, centroids).min(axis=1): float64, sqrt(sum_k (x_k - c_k)^2) summed over k ascending
__global__ void centroid_min_dist_kernel(const float* __restrict__ x, const float* __restrict__ cent, int64_t n, int d,
                                         int nc, double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double best = INFINITY;
  for (int c = 0; c < nc; ++c) {
    double s = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = __dsub_rn((double)x[r * d + k], (double)cent[(int64_t)c * d + k]);
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    const double dd = __dsqrt_rn(s);
    best = dd < best || dd != dd ? dd : best;   // np.min propagates NaN
  }
  out[r] = best;
}

}  // namespace
}  // namespace knn

using namespace knn;

extern "C" int knn_triplet_mine(const float* dist, const int64_t* labels, int64_t n, float margin, float* hard,
                                double* all_sum, long long* all_pos, long long* all_valid, void* stream) {
  KNN_REQUIRE(n >= 0, "knn_triplet_mine: bad size");
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(dist && labels && (hard || all_sum), "knn_triplet_mine: null pointer");
  KNN_REQUIRE((all_sum == nullptr) == (all_pos == nullptr) && (all_sum == nullptr) == (all_valid == nullptr),
              "knn_triplet_mine: the three batch-all outputs come together");
  triplet_mine_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(dist, labels, n, margin, hard, all_sum, all_pos,
                                                                    all_valid);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_jaccard_matrix(const uint64_t* q_masks, const uint64_t* g_masks, int64_t nq, int64_t ng, float eps,
                                  float* out, void* stream) {
  KNN_REQUIRE(nq >= 0 && ng >= 0, "knn_jaccard_matrix: bad sizes");
  if (nq == 0 || ng == 0) return KNN_OK;
  KNN_REQUIRE(q_masks && g_masks && out, "knn_jaccard_matrix: null pointer");
  const int64_t total = nq * ng;
  jaccard_matrix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_masks, g_masks, nq, ng, eps,
                                                                                          out);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_class_means(const float* x, const int64_t* labels, int64_t n, int d, const int64_t* classes,
                               int nclasses, float* means, int64_t* counts, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1 && nclasses >= 1, "knn_class_means: bad sizes");
  KNN_REQUIRE(x && labels && classes && means, "knn_class_means: null pointer");
  const int64_t total = (int64_t)nclasses * d;
  class_means_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, labels, n, d, classes, nclasses,
                                                                                       means, counts);
  KNN_LAUNCHED();
  return KNN_OK;
}

extern "C" int knn_centroid_min_dist(const float* x, const float* centroids, int64_t n, int d, int ncentroids,
                                     double* out, void* stream) {
  KNN_REQUIRE(n >= 0 && d >= 1 && ncentroids >= 1, "knn_centroid_min_dist: bad sizes");
  if (n == 0) return KNN_OK;
  KNN_REQUIRE(x && centroids && out, "knn_centroid_min_dist: null pointer");
  centroid_min_dist_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, centroids, n, d, ncentroids,
                                                                                         out);
  KNN_LAUNCHED();
  return KNN_OK;
}
