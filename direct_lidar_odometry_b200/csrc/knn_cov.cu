// knn_cov.cu — K2+K3: k-nearest-neighbour search fused with the plane-regularised covariance
// estimate (reference include/nano_gicp/impl/nano_gicp_impl.hpp:298-357), plus the raw kNN query
// entry used by the parity tests (KdTreeFLANN::nearestKSearch, include/nano_gicp/nanoflann.hpp:141-152).
//
// One warp per query for the search (coalesced candidate scans, warp-distributed top-k list), then
// one THREAD per query for the fp64 statistics + 3x3 Jacobi SVD: a warp first finds the neighbour
// lists of 32 consecutive cell-ordered points (lists parked in shared memory), then its 32 lanes
// each finish one point.  Algorithmic bytes: 16 B point read + 48 B covariance write = 64 B/point.
#include "internal.h"
#include "grid_search.cuh"
#include "gicp_math.cuh"

namespace ngicp {

constexpr int KC_THREADS = 256;
constexpr int KC_WARPS = KC_THREADS / 32;
constexpr int KC_ROW = 33;  // padded row: lane j writes [j][t], lane t reads [j][t] — both conflict-free

// queries: arbitrary points (float4 xyz), results in ORIGINAL index order of the cloud
__global__ void __launch_bounds__(KC_THREADS) knn_query_kernel(GridView g, const float4* __restrict__ queries, int nq, int k,
                                                               int* __restrict__ idx, float* __restrict__ d2) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const GridParams gp = load_grid(g.desc);
  for (int q = warp; q < nq; q += nwarps) {
    const float4 qp = queries[q];
    WarpTopK rs;
    rs.init(k, lane);
    if (isfinite(qp.x) && isfinite(qp.y) && isfinite(qp.z)) grid_search_warp(g, gp, qp.x, qp.y, qp.z, FLT_MAX, rs);
    if (lane < k) {
      int o = -1;
      float dd = -1.f;
      if (rs.p >= 0) { o = __float_as_int(__ldg(g.sorted + rs.p).w); dd = rs.d; }
      idx[(size_t)q * k + lane] = o;
      d2[(size_t)q * k + lane] = dd;
    }
  }
}

__global__ void __launch_bounds__(KC_THREADS) knn_cov_kernel(GridView g, int n, int k, int method, double* __restrict__ covs6) {
  extern __shared__ int nbr_smem[];  // [KC_WARPS][k][KC_ROW]
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  int* nbr = nbr_smem + (size_t)w * k * KC_ROW;
  const GridParams gp = load_grid(g.desc);
  const int nbatches = (n + 31) >> 5;
  for (int batch = blockIdx.x * KC_WARPS + w; batch < nbatches; batch += gridDim.x * KC_WARPS) {
    const int q0 = batch << 5;
    // phase A: the warp searches the neighbours of 32 consecutive cell-ordered points
    const int cnt = min(32, n - q0);
    for (int t = 0; t < cnt; ++t) {
      const float4 qp = __ldg(g.sorted + q0 + t);
      WarpTopK rs;
      rs.init(k, lane);
      if (isfinite(qp.x) && isfinite(qp.y) && isfinite(qp.z)) grid_search_warp(g, gp, qp.x, qp.y, qp.z, FLT_MAX, rs);
      if (lane < k) nbr[lane * KC_ROW + t] = rs.p;
    }
    __syncwarp();
    // phase B: one lane per point — mean, covariance / k, regularisation (all fp64)
    if (lane < cnt) {
      const float4 self = __ldg(g.sorted + q0 + lane);
      const int orig = __float_as_int(self.w);
      double mx = 0.0, my = 0.0, mz = 0.0;
      for (int j = 0; j < k; ++j) {
        const int p = nbr[j * KC_ROW + lane];
        if (p >= 0) { const float4 c = __ldg(g.sorted + p); mx += (double)c.x; my += (double)c.y; mz += (double)c.z; }
      }
      const double kd = (double)k;
      mx /= kd; my /= kd; mz /= kd;
      double c[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 0; j < k; ++j) {
        const int p = nbr[j * KC_ROW + lane];
        double x = -mx, y = -my, z = -mz;   // a missing neighbour is a zero column minus the mean, like the reference's zero-initialised matrix would be
        if (p >= 0) { const float4 v = __ldg(g.sorted + p); x += (double)v.x; y += (double)v.y; z += (double)v.z; }
        c[0] += x * x; c[1] += x * y; c[2] += x * z; c[3] += y * y; c[4] += y * z; c[5] += z * z;
      }
#pragma unroll
      for (int i = 0; i < 6; i++) c[i] /= kd;
      double out[6];
      regularize_cov(c, method, out);
      double* dst = covs6 + (size_t)orig * 6;
#pragma unroll
      for (int i = 0; i < 6; i++) dst[i] = out[i];
    }
    __syncwarp();
  }
}

cudaError_t launch_knn_queries(const DevCloud& c, const float4* queries, int nq, int k, int* idx, float* d2, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  int blocks = (nq + KC_WARPS - 1) / KC_WARPS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  knn_query_kernel<<<blocks, KC_THREADS, 0, st>>>(c.view(), queries, nq, k, idx, d2);
  note_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_covariances(const DevCloud& c, int k, int method, double* covs6, cudaStream_t st) {
  if (c.n <= 0) return cudaSuccess;
  const int nbatches = (c.n + 31) / 32;
  int blocks = (nbatches + KC_WARPS - 1) / KC_WARPS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const size_t smem = sizeof(int) * (size_t)KC_WARPS * k * KC_ROW;
  knn_cov_kernel<<<blocks, KC_THREADS, smem, st>>>(c.view(), c.n, k, method, covs6);
  note_launches(1);
  return cudaGetLastError();
}

}  // namespace ngicp
