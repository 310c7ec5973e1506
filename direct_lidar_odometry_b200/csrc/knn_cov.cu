// knn_cov.cu — K2+K3: k-nearest-neighbour search fused with the plane-regularised covariance
// estimate (reference include/nano_gicp/impl/nano_gicp_impl.hpp:298-357), plus the raw kNN query
// entry used by the parity tests (KdTreeFLANN::nearestKSearch, include/nano_gicp/nanoflann.hpp:141-152).
//
// Two launches: knn_lists_kernel — one WARP per point for the search (coalesced candidate scans,
// warp-distributed top-k list, few registers so that many warps hide the gather latency), neighbour
// slots parked in HBM (4k B/point of scratch, L2 resident); cov_from_lists_kernel — one THREAD per
// point for the fp64 statistics + 3x3 Jacobi SVD.  Algorithmic bytes of the pair: 16 B point read +
// 48 B covariance write = 64 B/point.
#include <cstdlib>
#include "internal.h"
#include "grid_search.cuh"
#include "gicp_math.cuh"

namespace ngicp {

constexpr int KC_THREADS = 256;
constexpr int KC_WARPS = KC_THREADS / 32;
#ifndef KNN_MIN_BLOCKS
#define KNN_MIN_BLOCKS 5
#endif

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
__global__ void __launch_bounds__(KC_THREADS, KNN_MIN_BLOCKS) knn_lists_kernel(GridView g, int n, int k, int* __restrict__ nbr) {
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

// K2, staged variant (the one launch_covariances uses): consecutive cell-ordered points mostly share their cell, so a
// warp copies the 3x3x3 block of cells around the current cell into shared memory ONCE (9 coalesced runs, SoA) and
// answers every query of that cell from there.  Selection is threshold-and-verify instead of list maintenance:
// collect the candidates closer than a guessed T^2 (first guess from the block's population, then 1.3x the previous
// query's k-th distance), accept when between k and 32 were collected — one bitonic sort of 32 then yields the
// ascending k nearest — else bisect T^2.  T^2 never exceeds the squared distance to the nearest unexplored cell face,
// so an accepted answer is exact; queries that cannot be decided inside the block (sparse regions, very dense
// cells) fall back to the growing-cube search above.
constexpr int ST_WARPS = 4;
constexpr int ST_CMAX = 896;
struct StageSmem {
  float x[ST_CMAX], y[ST_CMAX], z[ST_CMAX];
  float sel_d[32];
  int sel_i[32];
  int row_a[9];
  int row_excl[9];
};

__global__ void __launch_bounds__(ST_WARPS * 32) knn_lists_staged_kernel(GridView g, int n, int k, int qch, int* __restrict__ nbr) {
  __shared__ StageSmem smem[ST_WARPS];
  const int lane = threadIdx.x & 31;
  StageSmem& S = smem[threadIdx.x >> 5];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const GridParams gp = load_grid(g.desc);
  const unsigned lt = (1u << lane) - 1u;
  const int nchunks = (n + qch - 1) / qch;
  for (int chunk = warp; chunk < nchunks; chunk += nwarps) {
    int scx = -1, scy = -1, scz = -1, C = 0;
    bool staged_ok = false;
    float t2_prev = 0.f;
    const int qend = min(n, (chunk + 1) * qch);
    for (int q = chunk * qch; q < qend; ++q) {
      const float4 qp = __ldg(g.sorted + q);
      if (!(isfinite(qp.x) && isfinite(qp.y) && isfinite(qp.z))) {
        if (lane < k) nbr[(size_t)q * k + lane] = -1;
        continue;
      }
      const int cx = cell_coord(qp.x, gp.ox, gp.inv, gp.dx);
      const int cy = cell_coord(qp.y, gp.oy, gp.inv, gp.dy);
      const int cz = cell_coord(qp.z, gp.oz, gp.inv, gp.dz);
      if (cx != scx || cy != scy || cz != scz) {
        // ---- stage the block of cells around (cx,cy,cz) ----
        __syncwarp();
        int a = 0, b = 0;
        if (lane < 9) {
          const int y = cy + (lane % 3) - 1, z = cz + (lane / 3) - 1;
          if (y >= 0 && y < gp.dy && z >= 0 && z < gp.dz) {
            const int* row = g.cell_start + (z * gp.dy + y) * gp.dx;
            a = __ldg(row + max(cx - 1, 0));
            b = __ldg(row + min(cx + 1, gp.dx - 1) + 1);
          }
        }
        const int len = b - a;
        int inc = len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        C = __shfl_sync(FULL, inc, 8);
        const int excl = inc - len;
        staged_ok = C <= ST_CMAX;
        if (staged_ok) {
          if (lane < 9) { S.row_a[lane] = a; S.row_excl[lane] = excl; }
          for (int t0 = 0; t0 < C; t0 += 32) {
            const int t = t0 + lane;
            int j = 0;
#pragma unroll
            for (int step = 8; step > 0; step >>= 1) {
              const int v = __shfl_sync(FULL, inc, j + step - 1);
              if (v <= t) j += step;
            }
            const int ja = __shfl_sync(FULL, a, j), je = __shfl_sync(FULL, excl, j);
            if (t < C) {
              const float4 c = __ldg(g.sorted + ja + (t - je));
              S.x[t] = c.x; S.y[t] = c.y; S.z[t] = c.z;
            }
          }
        }
        __syncwarp();
        scx = cx; scy = cy; scz = cz;
        t2_prev = 0.f;
      }
      bool done = false;
      if (staged_ok && C >= k) {
        const int rmax = max(max(max(cx, gp.dx - 1 - cx), max(cy, gp.dy - 1 - cy)), max(cz, gp.dz - 1 - cz));
        float m2 = FLT_MAX;   // the block already covers the whole grid
        if (rmax > 1) {
          const float m = cube_face_distance(gp, cx, cy, cz, 1, qp.x, qp.y, qp.z);
          m2 = m > 0.f ? m * m * 0.999999f : 0.f;
        }
        float T2 = t2_prev > 0.f ? t2_prev * 1.3f : 2.865f * (float)(k + 6) * gp.cell * gp.cell / (float)C;
        T2 = fminf(T2, m2);
        float lo = 0.f, hi = -1.f;
        for (int tries = 0; tries < 8; ++tries) {
          int cnt = 0;
          for (int t0 = 0; t0 < C; t0 += 32) {
            const int t = t0 + lane;
            float d = INFINITY;
            if (t < C) d = sqdist_unfused(qp.x, qp.y, qp.z, S.x[t], S.y[t], S.z[t]);
            const bool pass = d < T2;
            const unsigned mk = __ballot_sync(FULL, pass);
            if (pass) {
              const int pos = cnt + __popc(mk & lt);
              if (pos < 32) { S.sel_d[pos] = d; S.sel_i[pos] = t; }
            }
            cnt += __popc(mk);
          }
          __syncwarp();
          if (cnt >= k && cnt <= 32) {
            float d = lane < cnt ? S.sel_d[lane] : INFINITY;
            int ci = lane < cnt ? S.sel_i[lane] : -1;
#pragma unroll
            for (int size = 2; size <= 32; size <<= 1) {
              const bool asc = (size == 32) || ((lane & size) == 0);
#pragma unroll
              for (int stride = size >> 1; stride > 0; stride >>= 1) WarpTopK::cex(d, ci, stride, ((lane & stride) == 0) == asc);
            }
            t2_prev = __shfl_sync(FULL, d, k - 1);
            if (lane < k) {
              int j = 0;
#pragma unroll
              for (int r = 1; r < 9; ++r) if (ci >= S.row_excl[r]) j = r;
              nbr[(size_t)q * k + lane] = S.row_a[j] + (ci - S.row_excl[j]);
            }
            done = true;
            __syncwarp();
            break;
          }
          if (cnt < k) {
            if (T2 >= m2) { __syncwarp(); break; }   // not decidable inside the block
            lo = T2;
            T2 = hi > 0.f ? sqrtf(lo * hi) : T2 * 1.7f;
            T2 = fminf(T2, m2);
          } else {
            hi = T2;
            T2 = lo > 0.f ? sqrtf(lo * hi) : T2 * 0.6f;
          }
          __syncwarp();
        }
      }
      if (!done) {
        WarpTopK rs;
        rs.init(k, lane);
        grid_search_warp(g, gp, qp.x, qp.y, qp.z, FLT_MAX, rs, k >= 4);
        if (lane < k) nbr[(size_t)q * k + lane] = rs.p;
        t2_prev = (rs.kth < FLT_MAX) ? rs.kth : 0.f;
      }
    }
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
  // queries per warp chunk: long enough to reuse a staged block, short enough to fill the GPU on small clouds
  int qch = c.n / (148 * 20);
  qch = qch < 4 ? 4 : (qch > 32 ? 32 : qch);
  const int nchunks = (c.n + qch - 1) / qch;
  int blocks = (nchunks + ST_WARPS - 1) / ST_WARPS;
  if (blocks > 148 * 64) blocks = 148 * 64;
  // the shared-memory staged variant executes fewer instructions but (occupancy 14 warps/SM, frequent fall-backs in
  // sparse cells) is slower end to end on the C2 submap (1.26 ms vs 0.85 ms, profiles/); kept behind a switch
  static const bool use_legacy = getenv("NGICP_KNN_STAGED") == nullptr;
  if (use_legacy) knn_lists_kernel<<<(c.n + KC_WARPS - 1) / KC_WARPS, KC_THREADS, 0, st>>>(c.view(), c.n, k, nbr_scratch);
  else knn_lists_staged_kernel<<<blocks, ST_WARPS * 32, 0, st>>>(c.view(), c.n, k, qch, nbr_scratch);
  cov_from_lists_kernel<<<(c.n + 127) / 128, 128, 0, st>>>(c.view(), c.n, k, method, nbr_scratch, covs6);
  note_launches(2);
  return cudaGetLastError();
}

}  // namespace ngicp
