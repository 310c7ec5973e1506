// knn_cov.cu — K2+K3: k-nearest-neighbour search fused with the plane-regularised covariance
// estimate (reference include/nano_gicp/impl/nano_gicp_impl.hpp:298-357), plus the raw kNN query
// entry used by the parity tests (KdTreeFLANN::nearestKSearch, include/nano_gicp/nanoflann.hpp:141-152).
//
// Two launches: knn_lists_kernel — one WARP per point for the search (coalesced candidate scans,
// warp-distributed top-k list, few registers so that many warps hide the gather latency), neighbour
// slots parked in HBM (4k B/point of scratch, L2 resident); cov_from_lists_kernel — one THREAD per
// point for the fp64 statistics + 3x3 Jacobi SVD.  Algorithmic bytes of the pair: 16 B point read +
// 48 B covariance write = 64 B/point.
#include "internal.h"
#include "grid_search.cuh"
#include "gicp_math.cuh"

namespace ngicp {

constexpr int KC_THREADS = 256;
constexpr int KC_WARPS = KC_THREADS / 32;

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

// K2: neighbour lists of the cloud's own points, one warp per point, visited in cell order so that
// concurrently running warps read the same cells.  nbr[q*k + j] = sorted slot of the j-th neighbour.
__global__ void __launch_bounds__(KC_THREADS) knn_lists_kernel(GridView g, int n, int k, int* __restrict__ nbr) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const GridParams gp = load_grid(g.desc);
  for (int q = warp; q < n; q += nwarps) {
    const float4 qp = __ldg(g.sorted + q);
    WarpTopK rs;
    rs.init(k, lane);
    if (isfinite(qp.x) && isfinite(qp.y) && isfinite(qp.z)) grid_search_warp(g, gp, qp.x, qp.y, qp.z, FLT_MAX, rs, k >= 4);
    if (lane < k) nbr[(size_t)q * k + lane] = rs.p;
  }
}

// K3: one thread per point — mean, covariance / k, regularisation, all fp64 (nano_gicp_impl.hpp:315-353)
__global__ void __launch_bounds__(128) cov_from_lists_kernel(GridView g, int n, int k, int method, const int* __restrict__ nbr,
                                                             double* __restrict__ covs6) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int* my = nbr + (size_t)q * k;
  const int orig = __float_as_int(__ldg(g.sorted + q).w);
  double mx = 0.0, my_ = 0.0, mz = 0.0;
  for (int j = 0; j < k; ++j) {
    const int p = __ldg(my + j);
    if (p >= 0) { const float4 c = __ldg(g.sorted + p); mx += (double)c.x; my_ += (double)c.y; mz += (double)c.z; }
  }
  const double kd = (double)k;
  mx /= kd; my_ /= kd; mz /= kd;
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < k; ++j) {
    const int p = __ldg(my + j);
    double x = -mx, y = -my_, z = -mz;   // a missing neighbour is a zero column minus the mean, like the reference's zero-initialised matrix would be
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

cudaError_t launch_knn_queries(const DevCloud& c, const float4* queries, int nq, int k, int* idx, float* d2, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  int blocks = (nq + KC_WARPS - 1) / KC_WARPS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  knn_query_kernel<<<blocks, KC_THREADS, 0, st>>>(c.view(), queries, nq, k, idx, d2);
  note_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_covariances(const DevCloud& c, int k, int method, int* nbr_scratch, double* covs6, cudaStream_t st) {
  if (c.n <= 0) return cudaSuccess;
  int blocks = (c.n + KC_WARPS - 1) / KC_WARPS;
  if (blocks > 148 * 32) blocks = 148 * 32;
  knn_lists_kernel<<<blocks, KC_THREADS, 0, st>>>(c.view(), c.n, k, nbr_scratch);
  cov_from_lists_kernel<<<(c.n + 127) / 128, 128, 0, st>>>(c.view(), c.n, k, method, nbr_scratch, covs6);
  note_launches(2);
  return cudaGetLastError();
}

}  // namespace ngicp
