// grid_search.cuh — K2: exact nearest-neighbour search on the uniform grid, one warp per query.
//
// Replaces nanoflann's recursive kd-tree descent (reference impl/nanoflann_impl.hpp:1230-1250,
// searchLevel :1355-1418) and its KNNResultSet (:149-214).  Same result definition:
//   * metric  d = ((dx*dx) + dy*dy) + dz*dz in float, unfused (:441-449)
//   * a candidate enters only if d < current k-th distance (strict), list kept ascending,
//     an equal-distance newcomer never displaces an earlier entry (:193)
// so sorted distance lists are bit-identical to the kd-tree's and indices agree wherever distances
// are distinct (ties depend on visiting order in both implementations).
//
// Search = expanding Chebyshev shells of cells around the query cell.  Because cell_start is a
// lower-bound table over keys with x fastest, every x-run of cells is one contiguous slot range, so
// shell s costs 8s full rows + 2(2s-1)^2 end cells, each a coalesced range scan.
// Exactness: after shell s every unvisited point lies beyond an interior face of the visited cube;
// the search stops only when (distance to the nearest such face - margin)^2 >= current k-th distance
// (or >= the caller's cap, e.g. max-correspondence-distance^2), or when the cube covers the grid.
#pragma once
#include "common.cuh"

namespace ngicp {

struct GridParams {
  float ox, oy, oz, cell, inv, margin;
  int dx, dy, dz;
};

__device__ __forceinline__ GridParams load_grid(const GridDesc* __restrict__ d) {
  GridParams g;
  g.ox = d->origin[0]; g.oy = d->origin[1]; g.oz = d->origin[2];
  g.cell = d->cell; g.inv = d->inv_cell; g.margin = d->margin;
  g.dx = d->dim[0]; g.dy = d->dim[1]; g.dz = d->dim[2];
  return g;
}

// k-NN result set distributed over the warp: lane j holds the j-th best (k <= 32).  All 32 lanes
// always hold the 32 best seen so far (ascending); kth = entry k-1 is the acceptance bound.
struct WarpTopK {
  float d;
  int p;
  float kth;
  int k;
  int lane;
  __device__ __forceinline__ void init(int k_, int lane_) { k = k_; lane = lane_; d = FLT_MAX; p = -1; kth = FLT_MAX; }
  __device__ __forceinline__ float bound() { return kth; }

  // one compare-exchange step of a bitonic network on (cd, cp); keep_min lanes keep the smaller key
  static __device__ __forceinline__ void cex(float& cd, int& cp, int stride, bool keep_min) {
    const float od = __shfl_xor_sync(FULL, cd, stride);
    const int op = __shfl_xor_sync(FULL, cp, stride);
    const bool take = keep_min ? (od < cd) : (od > cd);
    if (take) { cd = od; cp = op; }
  }

  // many candidates pass the bound at once: sort the batch (bitonic, 15 steps) and merge it into the
  // list (reverse + elementwise min gives the 32 smallest as a bitonic sequence; 5 more steps sort it)
  __device__ __forceinline__ void merge_batch(float cd, int cp) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
      const bool asc = (size == 32) || ((lane & size) == 0);
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) cex(cd, cp, stride, ((lane & stride) == 0) == asc);
    }
    const float rd = __shfl_sync(FULL, cd, 31 - lane);
    const int rp = __shfl_sync(FULL, cp, 31 - lane);
    if (rd < d) { d = rd; p = rp; }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) cex(d, p, stride, (lane & stride) == 0);
    kth = __shfl_sync(FULL, d, k - 1);
  }

  __device__ __forceinline__ void offer(bool valid, float cd, int cp) {
    const bool pass = valid && cd < kth;
    unsigned m = __ballot_sync(FULL, pass);
    if (m == 0u) return;
    if (__popc(m) > 10) {
      merge_batch(pass ? cd : INFINITY, cp);
      return;
    }
    while (m) {
      const int l = __ffs(m) - 1;
      const float bd = __shfl_sync(FULL, cd, l);
      const int bp = __shfl_sync(FULL, cp, l);
      const float du = __shfl_up_sync(FULL, d, 1);
      const int pu = __shfl_up_sync(FULL, p, 1);
      const unsigned gtm = __ballot_sync(FULL, d > bd);   // a suffix of lanes: the list is ascending
      const int first = __ffs(gtm) - 1;                    // -1 when nothing is larger
      if (first >= 0) {
        if (lane > first) { d = du; p = pu; }
        else if (lane == first) { d = bd; p = bp; }
      }
      kth = __shfl_sync(FULL, d, k - 1);
      const unsigned above = (l == 31) ? 0u : (FULL << (l + 1));
      m = __ballot_sync(FULL, valid && cd < kth) & above;
    }
  }
};

// 1-NN: every lane keeps the best candidate it has seen; merged at the end
struct WarpBest1 {
  float d;
  int p;
  __device__ __forceinline__ void init() { d = FLT_MAX; p = -1; }
  __device__ __forceinline__ float bound() { return warp_min(d); }
  __device__ __forceinline__ void offer(bool valid, float cd, int cp) {
    if (valid && cd < d) { d = cd; p = cp; }
  }
  // warp-uniform result: smallest distance, lowest sorted slot among equals
  __device__ __forceinline__ void finalize() {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(FULL, d, o);
      const int op = __shfl_xor_sync(FULL, p, o);
      if (od < d || (od == d && (unsigned)op < (unsigned)p)) { d = od; p = op; }
    }
  }
};

template <class RS>
__device__ __forceinline__ void grid_search_warp(const GridView& g, const GridParams& gp, float qx, float qy, float qz,
                                                 float cap_d2, RS& rs) {
  const int lane = threadIdx.x & 31;
  const int cx = cell_coord(qx, gp.ox, gp.inv, gp.dx);
  const int cy = cell_coord(qy, gp.oy, gp.inv, gp.dy);
  const int cz = cell_coord(qz, gp.oz, gp.inv, gp.dz);
  const int rmax = max(max(max(cx, gp.dx - 1 - cx), max(cy, gp.dy - 1 - cy)), max(cz, gp.dz - 1 - cz));
  for (int s = 0; s <= rmax; ++s) {
    const int side = 2 * s - 1;
    const int nfull = s == 0 ? 1 : 8 * s;
    const int ncap = s == 0 ? 0 : side * side;
    const int nslots = nfull + 2 * ncap;
    for (int base = 0; base < nslots; base += 32) {
      const int slot = base + lane;
      int a = 0, b = 0;
      if (slot < nslots) {
        int yy, zz, x0, x1;
        if (slot < nfull) {
          if (s == 0) { yy = 0; zz = 0; }
          else {
            const int sd = slot / (2 * s), o = slot - sd * (2 * s);
            if (sd == 0) { yy = -s + o; zz = -s; }
            else if (sd == 1) { yy = s; zz = -s + o; }
            else if (sd == 2) { yy = s - o; zz = s; }
            else { yy = -s; zz = s - o; }
          }
          x0 = cx - s; x1 = cx + s;
        } else {
          const int u = slot - nfull;
          const int v = u >> 1;
          zz = v / side;
          yy = v - zz * side - (s - 1);
          zz -= (s - 1);
          x0 = x1 = (u & 1) ? cx + s : cx - s;
        }
        const int y = cy + yy, z = cz + zz;
        const int xa = max(x0, 0), xb = min(x1, gp.dx - 1);
        if (y >= 0 && y < gp.dy && z >= 0 && z < gp.dz && xa <= xb) {
          const int row = (z * gp.dy + y) * gp.dx;
          a = __ldg(g.cell_start + row + xa);
          b = __ldg(g.cell_start + row + xb + 1);
        }
      }
      unsigned ne = __ballot_sync(FULL, b > a);
      while (ne) {
        const int l = __ffs(ne) - 1;
        ne &= ne - 1;
        const int ra = __shfl_sync(FULL, a, l), rb = __shfl_sync(FULL, b, l);
        for (int p0 = ra; p0 < rb; p0 += 32) {
          const int p = p0 + lane;
          const bool valid = p < rb;
          float d = FLT_MAX;
          if (valid) {
            const float4 c = __ldg(g.sorted + p);
            d = sqdist_unfused(qx, qy, qz, c.x, c.y, c.z);
          }
          rs.offer(valid, d, p);
        }
      }
    }
    if (s == rmax) break;
    // distance from the query to the nearest face of the visited cube that still has cells behind it
    float m = FLT_MAX;
    if (cx - s > 0) m = fminf(m, qx - (gp.ox + (float)(cx - s) * gp.cell));
    if (cx + s + 1 < gp.dx) m = fminf(m, (gp.ox + (float)(cx + s + 1) * gp.cell) - qx);
    if (cy - s > 0) m = fminf(m, qy - (gp.oy + (float)(cy - s) * gp.cell));
    if (cy + s + 1 < gp.dy) m = fminf(m, (gp.oy + (float)(cy + s + 1) * gp.cell) - qy);
    if (cz - s > 0) m = fminf(m, qz - (gp.oz + (float)(cz - s) * gp.cell));
    if (cz + s + 1 < gp.dz) m = fminf(m, (gp.oz + (float)(cz + s + 1) * gp.cell) - qz);
    m -= gp.margin;
    if (m > 0.f) {
      const float m2 = m * m * 0.999999f;
      if (cap_d2 <= m2) break;
      if (rs.bound() <= m2) break;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// 1-NN with ONE THREAD per query (used by the registration kernels, where every source point needs
// its nearest target point within the max-correspondence distance).  Same exact stop rule as the
// warp version; the 3x3x3 block around the query cell is scanned first as 9 contiguous x-runs
// whose ranges are fetched up front, further shells (rare: only when the cell edge is smaller than
// the cap or the cap is unbounded) are walked slot by slot.
// Strict '<' keeps the first visited of equidistant points, like nanoflann's result set does.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void nn1_scan_range(const GridView& g, int a, int b, float qx, float qy, float qz, float& bd, int& bp) {
#pragma unroll 4
  for (int p = a; p < b; ++p) {
    const float4 c = __ldg(g.sorted + p);
    const float d = sqdist_unfused(qx, qy, qz, c.x, c.y, c.z);
    if (d < bd) { bd = d; bp = p; }
  }
}

__device__ __forceinline__ float cube_face_distance(const GridParams& gp, int cx, int cy, int cz, int s, float qx, float qy, float qz) {
  float m = FLT_MAX;
  if (cx - s > 0) m = fminf(m, qx - (gp.ox + (float)(cx - s) * gp.cell));
  if (cx + s + 1 < gp.dx) m = fminf(m, (gp.ox + (float)(cx + s + 1) * gp.cell) - qx);
  if (cy - s > 0) m = fminf(m, qy - (gp.oy + (float)(cy - s) * gp.cell));
  if (cy + s + 1 < gp.dy) m = fminf(m, (gp.oy + (float)(cy + s + 1) * gp.cell) - qy);
  if (cz - s > 0) m = fminf(m, qz - (gp.oz + (float)(cz - s) * gp.cell));
  if (cz + s + 1 < gp.dz) m = fminf(m, (gp.oz + (float)(cz + s + 1) * gp.cell) - qz);
  return m - gp.margin;
}

__device__ __forceinline__ void grid_nn1_thread(const GridView& g, const GridParams& gp, float qx, float qy, float qz,
                                                float cap_d2, float& best_d, int& best_p) {
  best_d = FLT_MAX;
  best_p = -1;
  const int cx = cell_coord(qx, gp.ox, gp.inv, gp.dx);
  const int cy = cell_coord(qy, gp.oy, gp.inv, gp.dy);
  const int cz = cell_coord(qz, gp.oz, gp.inv, gp.dz);
  const int rmax = max(max(max(cx, gp.dx - 1 - cx), max(cy, gp.dy - 1 - cy)), max(cz, gp.dz - 1 - cz));
  // cube of radius 1: nine x-runs, ranges loaded first so the table reads overlap
  {
    int ra[9], rb[9];
    const int xa = max(cx - 1, 0), xb = min(cx + 1, gp.dx - 1);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
      ra[r] = 0; rb[r] = 0;
      if (y >= 0 && y < gp.dy && z >= 0 && z < gp.dz) {
        const int row = (z * gp.dy + y) * gp.dx;
        ra[r] = __ldg(g.cell_start + row + xa);
        rb[r] = __ldg(g.cell_start + row + xb + 1);
      }
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) nn1_scan_range(g, ra[r], rb[r], qx, qy, qz, best_d, best_p);
  }
  for (int s = 1; s < rmax; ) {
    const float m = cube_face_distance(gp, cx, cy, cz, s, qx, qy, qz);
    if (m > 0.f) {
      const float m2 = m * m * 0.999999f;
      if (cap_d2 <= m2 || best_d <= m2) break;
    }
    ++s;
    // shell s: rows with max(|yy|,|zz|) == s in full, the two end cells of the inner rows
    for (int zz = -s; zz <= s; ++zz) {
      const int z = cz + zz;
      if (z < 0 || z >= gp.dz) continue;
      for (int yy = -s; yy <= s; ++yy) {
        const int y = cy + yy;
        if (y < 0 || y >= gp.dy) continue;
        const int row = (z * gp.dy + y) * gp.dx;
        if (max(abs(yy), abs(zz)) == s) {
          const int xa = max(cx - s, 0), xb = min(cx + s, gp.dx - 1);
          nn1_scan_range(g, __ldg(g.cell_start + row + xa), __ldg(g.cell_start + row + xb + 1), qx, qy, qz, best_d, best_p);
        } else {
          if (cx - s >= 0) nn1_scan_range(g, __ldg(g.cell_start + row + cx - s), __ldg(g.cell_start + row + cx - s + 1), qx, qy, qz, best_d, best_p);
          if (cx + s < gp.dx) nn1_scan_range(g, __ldg(g.cell_start + row + cx + s), __ldg(g.cell_start + row + cx + s + 1), qx, qy, qz, best_d, best_p);
        }
      }
    }
  }
}

}  // namespace ngicp
