// Brute-force k-nearest-neighbour classification on the device: what the reference's evaluation does
// with sklearn on the host -- KNeighborsClassifier().fit(train embeddings, labels) in compute_knn
// (src/train_utils/knn.py:22-42) and estimator.predict(val embeddings) in eval_pretrained_model
// (src/train_utils/eval_functions.py:65-97): Euclidean metric, k = 5, uniform weights, majority vote, ties to the
// smallest label -- after a .cpu().numpy() of every batch.  Here the embeddings never leave HBM.
//   knn_dist_kernel    d2[q][t] = sum_f (Q[q][f] - X[t][f])^2 by direct differences (no |q|^2 + |t|^2 - 2 q.t
//                      cancellation: neighbour ranks must match a float64 host computation), 64 x 64 tiles, 4 x 4 per thread
//   knn_select_kernel  one warp per query: per-lane sorted top-k over a strided scan, k rounds of warp arg-min merge
//                      ((distance, index) lexicographic: deterministic), vote by class counts
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fb {

constexpr int kKnnMaxK = 16;

__global__ void __launch_bounds__(256) knn_dist_kernel(const float* __restrict__ Q, const float* __restrict__ X, int nq,
                                                       int nt, int dim, float* __restrict__ d2) {
  __shared__ float sq[16][64 + 1], sx[16][64 + 1];
  const int q0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int f0 = 0; f0 < dim; f0 += 16) {
    // 64 rows x 16 features of each side: thread -> (row = tid / 4, 4 consecutive features)
    const int row = threadIdx.x >> 2, fc = (threadIdx.x & 3) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = f0 + fc + e;
      sq[fc + e][row] = (q0 + row < nq && f < dim) ? __ldg(Q + (size_t)(q0 + row) * dim + f) : 0.f;
      sx[fc + e][row] = (t0 + row < nt && f < dim) ? __ldg(X + (size_t)(t0 + row) * dim + f) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < 16; ++f) {
      float a[4], b[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { a[e] = sq[f][ty * 4 + e]; b[e] = sx[f][tx * 4 + e]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float df = a[i] - b[j];
          acc[i][j] = fmaf(df, df, acc[i][j]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= nq) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx * 4 + j;
      if (t < nt) d2[(size_t)q * nt + t] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(128) knn_select_kernel(const float* __restrict__ d2, const int32_t* __restrict__ labels,
                                                         int nq, int nt, int k, int n_classes, int32_t* __restrict__ out,
                                                         int32_t* __restrict__ nbr_idx) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= nq) return;
  float bd[kKnnMaxK];
  int bi[kKnnMaxK];
#pragma unroll
  for (int e = 0; e < kKnnMaxK; ++e) { bd[e] = __int_as_float(0x7f800000); bi[e] = 0x7fffffff; }
  const float* row = d2 + (size_t)q * nt;
  for (int t = lane; t < nt; t += 32) {
    const float v = __ldg(row + t);
    // insert (v, t) into the lane's ascending list of the k best (indices ascend with the scan, so ties keep the older)
    if (v < bd[k - 1]) {
      float cd = v;
      int ci = t;
#pragma unroll
      for (int e = 0; e < kKnnMaxK; ++e) {
        if (e < k && (cd < bd[e])) {
          const float td = bd[e]; const int ti = bi[e];
          bd[e] = cd; bi[e] = ci; cd = td; ci = ti;
        }
      }
    }
  }
  // k rounds: the warp's smallest (distance, index) head wins and is popped from its lane's list
  int votes = 0;                                     // lane c counts class c (n_classes <= 32 per pass)
  for (int round = 0; round < k; ++round) {
    float hd = bd[0];
    int hi = bi[0];
    float md = hd; int mi = hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, md, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
    }
    if (hi == mi && hd == md) {                      // the winning lane pops its head
#pragma unroll
      for (int e = 0; e < kKnnMaxK - 1; ++e) { bd[e] = bd[e + 1]; bi[e] = bi[e + 1]; }
      bd[kKnnMaxK - 1] = __int_as_float(0x7f800000); bi[kKnnMaxK - 1] = 0x7fffffff;
    }
    if (mi < nt) {
      const int lab = __ldg(labels + mi);
      if (nbr_idx && lane == 0) nbr_idx[(size_t)q * k + round] = mi;
      if ((lab & 31) == lane && lab < 32) ++votes;
    } else if (nbr_idx && lane == 0) {
      nbr_idx[(size_t)q * k + round] = -1;
    }
  }
  // majority vote, ties to the smallest label (scipy.stats.mode / sklearn): max over (count, -label)
  int best = (votes << 8) | (31 - lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) out[q] = 31 - (best & 0xff);
  (void)n_classes;
}

}  // namespace fb
