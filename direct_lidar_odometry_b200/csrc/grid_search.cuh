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

// k > 32: the ascending list lives in the warp's shared memory (KNN_WIDE_MAX_K entries), one insertion at a time —
// count the entries not larger than the candidate (that is its place: equal distances keep their visiting order, like
// WarpTopK), move the tail up by one from the top down in warp-wide chunks, store.  Not a tuned path: DLO runs k = 10
// and 20; this keeps setCorrespondenceRandomness(k) meaningful for the values the reference accepts beyond 32.
struct WarpTopKWide {
  float* sd;
  int* sp;
  float kth;
  int k;
  int lane;
  __device__ __forceinline__ void init(int k_, int lane_, float* sd_, int* sp_) {
    k = k_; lane = lane_; sd = sd_; sp = sp_; kth = FLT_MAX;
    __syncwarp();
    for (int j = lane; j < k; j += 32) { sd[j] = FLT_MAX; sp[j] = -1; }
    __syncwarp();
  }
  __device__ __forceinline__ float bound() { return kth; }
  __device__ __forceinline__ void offer(bool valid, float cd, int cp) {
    unsigned m = __ballot_sync(FULL, valid && cd < kth);
    while (m) {
      const int l = __ffs(m) - 1;
      const float bd = __shfl_sync(FULL, cd, l);
      const int bp = __shfl_sync(FULL, cp, l);
      int pos = 0;
      for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        pos += __popc(__ballot_sync(FULL, j < k && sd[j] <= bd));
      }
      // bd < kth = sd[k-1], so pos <= k-1; entries [pos, k-2] move to [pos+1, k-1]
      for (int hi = k - 1; hi > pos; hi -= 32) {
        const int j = hi - lane;
        const bool mv = j > pos;
        float td = 0.f;
        int tp = 0;
        if (mv) { td = sd[j - 1]; tp = sp[j - 1]; }
        __syncwarp();
        if (mv) { sd[j] = td; sp[j] = tp; }
        __syncwarp();
      }
      if (lane == 0) { sd[pos] = bd; sp[pos] = bp; }
      __syncwarp();
      kth = sd[k - 1];
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

// scan sorted slots [ra, rb) with the whole warp
template <class RS>
__device__ __forceinline__ void scan_range_warp(const GridView& g, int ra, int rb, int lane, float qx, float qy, float qz, RS& rs) {
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

// The search grows a cube of cells around the query cell: radius s0 (already scanned, -1 = nothing) ->
// radius s1.  Each (y,z) row of the new cube is one lane's job: rows outside the old cube contribute the
// x-run [cx-s1, cx+s1], rows inside it only the two side runs [cx-s1, cx-s0-1] and [cx+s0+1, cx+s1].
// Radii go 0 (or 1 when `start1`), 1, 2, 3, 4, then grow by 50% per step so that isolated points in
// sparse regions do not pay O(r^3) single-cell probes.
// PRUNE (used by the 1-NN tail of the registration kernels): rows and side runs whose nearest possible point is already
// farther than min(current bound, cap) are skipped without touching memory — exact, the bounds are margin-shrunk.
// U = (y,z) row groups of 32 whose table lookups are issued together: the far shells of an isolated point are hundreds of
// mostly empty rows, each group a dependent round trip to the (uncached) cell table — with U = 4 four of them overlap.
template <class RS, bool PRUNE = false, int U = 1>
__device__ __forceinline__ void grid_search_warp(const GridView& g, const GridParams& gp, float qx, float qy, float qz,
                                                 float cap_d2, RS& rs, bool start1 = false, int resume_s0 = -1) {
  const int lane = threadIdx.x & 31;
  const int cx = cell_coord(qx, gp.ox, gp.inv, gp.dx);
  const int cy = cell_coord(qy, gp.oy, gp.inv, gp.dy);
  const int cz = cell_coord(qz, gp.oz, gp.inv, gp.dz);
  const int rmax = max(max(max(cx, gp.dx - 1 - cx), max(cy, gp.dy - 1 - cy)), max(cz, gp.dz - 1 - cz));
  int s0 = -1;
  int s1 = start1 ? min(1, rmax) : 0;
  if (resume_s0 >= 0) {   // the cube of radius resume_s0 was scanned by the caller, who also applied the stop test
    s0 = resume_s0;
    s1 = min((s0 < 4) ? s0 + 1 : s0 + (s0 >> 1), rmax);
  }
  for (;;) {
    const int w = 2 * s1 + 1;
    const int nrows = w * w;
    const int xlo = max(cx - s1, 0), xhi = min(cx + s1, gp.dx - 1);
    float lim = FLT_MAX, fx0 = 0.f, fx1 = 0.f, fy0 = 0.f, fy1 = 0.f, fz0 = 0.f, fz1 = 0.f;
    if (PRUNE) {
      lim = fminf(rs.bound(), cap_d2);
      fx0 = fmaxf(qx - (gp.ox + (float)cx * gp.cell) - gp.margin, 0.f);
      fx1 = fmaxf((gp.ox + (float)(cx + 1) * gp.cell) - qx - gp.margin, 0.f);
      fy0 = fmaxf(qy - (gp.oy + (float)cy * gp.cell) - gp.margin, 0.f);
      fy1 = fmaxf((gp.oy + (float)(cy + 1) * gp.cell) - qy - gp.margin, 0.f);
      fz0 = fmaxf(qz - (gp.oz + (float)cz * gp.cell) - gp.margin, 0.f);
      fz1 = fmaxf((gp.oz + (float)(cz + 1) * gp.cell) - qz - gp.margin, 0.f);
    }
    for (int base0 = 0; base0 < nrows; base0 += 32 * U) {
      int A1[U], B1[U], A2[U], B2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = base0 + u * 32 + lane;
        int a1 = 0, b1 = 0, a2 = 0, b2 = 0;
        if (r < nrows) {
          // first 3x3 step: own row first, then the edge-adjacent rows, then the corners (tighter bound earlier)
          const int rr = (s0 < 0 && s1 == 1) ? (int)((0x620837154ULL >> (4 * r)) & 15ULL) : r;
          const int rz = rr / w;
          const int yy = rr - rz * w - s1, zz = rz - s1;
          const int y = cy + yy, z = cz + zz;
          bool keep = (y >= 0 && y < gp.dy && z >= 0 && z < gp.dz);
          float dyz = 0.f;
          if (PRUNE) {
            const float dy = yy == 0 ? 0.f : (yy < 0 ? fy0 + (float)(-yy - 1) * gp.cell : fy1 + (float)(yy - 1) * gp.cell);
            const float dz = zz == 0 ? 0.f : (zz < 0 ? fz0 + (float)(-zz - 1) * gp.cell : fz1 + (float)(zz - 1) * gp.cell);
            dyz = (dy * dy + dz * dz) * 0.999999f;
            keep = keep && dyz < lim;
          }
          if (keep) {
            const int* row = g.cell_start + (z * gp.dy + y) * gp.dx;
            if (max(abs(yy), abs(zz)) > s0) {
              a1 = __ldg(row + xlo);
              b1 = __ldg(row + xhi + 1);
            } else {
              const int xl = cx - s0 - 1, xr = cx + s0 + 1;
              bool kl = xlo <= xl, kr = xr <= xhi;
              if (PRUNE) {   // nearest x distance to the side runs [.., cx-s0-1] and [cx+s0+1, ..]
                const float sxl = fx0 + (float)s0 * gp.cell, sxr = fx1 + (float)s0 * gp.cell;
                kl = kl && dyz + sxl * sxl * 0.999999f < lim;
                kr = kr && dyz + sxr * sxr * 0.999999f < lim;
              }
              if (kl) { a1 = __ldg(row + xlo); b1 = __ldg(row + xl + 1); }
              if (kr) { a2 = __ldg(row + xr); b2 = __ldg(row + xhi + 1); }
            }
          }
        }
        A1[u] = a1; B1[u] = b1; A2[u] = a2; B2[u] = b2;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (U > 1 && base0 + u * 32 >= nrows) break;
        const int a1 = A1[u], b1 = B1[u], a2 = A2[u], b2 = B2[u];
        // flatten the (mostly short) runs of the 32 rows into one candidate stream: inclusive scan of the run
        // lengths, then every lane finds the run its candidate index falls into (5-step search over lanes)
        const int len1 = b1 - a1;
        const int len = len1 + (b2 - a2);
        if (!__any_sync(FULL, len != 0)) continue;            // 32 empty rows: nothing to scan
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        const int total = __shfl_sync(FULL, inc, 31);
        const int excl = inc - len;
        for (int t0 = 0; t0 < total; t0 += 32) {
          const int t = t0 + lane;
          const bool valid = t < total;
          int j = 0;
#pragma unroll
          for (int step = 16; step > 0; step >>= 1) {
            const int v = __shfl_sync(FULL, inc, j + step - 1);
            if (v <= t) j += step;
          }
          const int off = t - __shfl_sync(FULL, excl, j);
          const int ja1 = __shfl_sync(FULL, a1, j), jl1 = __shfl_sync(FULL, len1, j), ja2 = __shfl_sync(FULL, a2, j);
          const int p = off < jl1 ? ja1 + off : ja2 + (off - jl1);
          float d = FLT_MAX;
          if (valid) {
            const float4 c = __ldg(g.sorted + p);
            d = sqdist_unfused(qx, qy, qz, c.x, c.y, c.z);
          }
          rs.offer(valid, d, p);
        }
      }
    }
    if (s1 >= rmax) break;
    // distance from the query to the nearest face of the scanned cube that still has cells behind it
    const float m = cube_face_distance(gp, cx, cy, cz, s1, qx, qy, qz);
    if (m > 0.f) {
      const float m2 = m * m * 0.999999f;
      if (cap_d2 <= m2) break;
      if (rs.bound() <= m2) break;
    }
    s0 = s1;
    s1 = (s1 < 4) ? s1 + 1 : s1 + (s1 >> 1);
    if (s1 > rmax) s1 = rmax;
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
// Candidates are fetched eight at a time with clamped addresses (no remainder loop: a one-by-one tail would be a chain
// of dependent round trips to L1/L2) and compared in slot order.
__device__ __forceinline__ void nn1_scan_range(const GridView& g, int a, int b, float qx, float qy, float qz, float& bd, int& bp) {
  for (int p0 = a; p0 < b; p0 += 8) {
    float4 c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = __ldg(g.sorted + min(p0 + u, b - 1));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float d = sqdist_unfused(qx, qy, qz, c[u].x, c[u].y, c[u].z);
      if (p0 + u < b && d < bd) { bd = d; bp = p0 + u; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Radius-1 cube search by a GROUP of LPP (1, 2 or 4) adjacent lanes per query.  Returns true when the search is NOT
// finished (a closer point may exist outside the cube and within the cap): the caller then hands the query to a warp
// (grid_search_warp with resume_s0 = 1) or continues with grid_nn1_grow_thread.
// One thread per query leaves the GPU almost empty for a 20k-point scan and every thread walks its candidates as a
// chain of dependent round trips; a group reads LPP x 8 candidates per round trip.  Strictly closer wins; among equals
// the earliest scanned range, inside a range the lowest slot.  All lanes of a group must call this together with the
// same query; results are group-uniform.
// ---------------------------------------------------------------------------------------------
template <int LPP>
__device__ __forceinline__ void nn1_group_scan(const GridView& g, int a, int b, int sub, unsigned gmask, float qx, float qy, float qz, float& bd, int& bp) {
  if (a >= b) return;                     // group-uniform
  float md = bd;
  int mp = bp;
  for (int p0 = a; p0 < b; p0 += 8 * LPP) {
    float4 c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = __ldg(g.sorted + min(p0 + u * LPP + sub, b - 1));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + u * LPP + sub;
      const float d = sqdist_unfused(qx, qy, qz, c[u].x, c[u].y, c[u].z);
      if (p < b && d < md) { md = d; mp = p; }
    }
  }
  // combine the group: smallest distance, lowest slot among equals.  (A lane that found nothing strictly closer still
  // holds the incumbent, so an incumbent at the minimum distance survives: every lane at that distance holds it.)
#pragma unroll
  for (int o = 1; o < LPP; o <<= 1) {
    const float od = __shfl_xor_sync(gmask, md, o);
    const int op = __shfl_xor_sync(gmask, mp, o);
    if (od < md || (od == md && op < mp)) { md = od; mp = op; }
  }
  bd = md; bp = mp;
}

template <int LPP>
__device__ __forceinline__ bool grid_nn1_cube1_group(const GridView& g, const GridParams& gp, int sub, unsigned gmask, float qx, float qy, float qz,
                                                     float cap_d2, float& best_d, int& best_p) {
  best_d = FLT_MAX;
  best_p = -1;
  const int cx = cell_coord(qx, gp.ox, gp.inv, gp.dx);
  const int cy = cell_coord(qy, gp.oy, gp.inv, gp.dy);
  const int cz = cell_coord(qz, gp.oz, gp.inv, gp.dz);
  const int rmax = max(max(max(cx, gp.dx - 1 - cx), max(cy, gp.dy - 1 - cy)), max(cz, gp.dz - 1 - cz));
  const int e0 = max(cx - 1, 0), e1 = cx, e2 = cx + 1, e3 = min(cx + 2, gp.dx);
  int v[9][4];
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
    const bool in = (y >= 0 && y < gp.dy && z >= 0 && z < gp.dz);
    const int* row = g.cell_start + (in ? (z * gp.dy + y) * gp.dx : 0);
    v[r][0] = in ? __ldg(row + e0) : 0;
    v[r][1] = in ? __ldg(row + e1) : 0;
    v[r][2] = in ? __ldg(row + e2) : 0;
    v[r][3] = in ? __ldg(row + e3) : 0;
  }
  const float fx0 = fmaxf(qx - (gp.ox + (float)cx * gp.cell) - gp.margin, 0.f);
  const float fx1 = fmaxf((gp.ox + (float)(cx + 1) * gp.cell) - qx - gp.margin, 0.f);
  const float fy0 = fmaxf(qy - (gp.oy + (float)cy * gp.cell) - gp.margin, 0.f);
  const float fy1 = fmaxf((gp.oy + (float)(cy + 1) * gp.cell) - qy - gp.margin, 0.f);
  const float fz0 = fmaxf(qz - (gp.oz + (float)cz * gp.cell) - gp.margin, 0.f);
  const float fz1 = fmaxf((gp.oz + (float)(cz + 1) * gp.cell) - qz - gp.margin, 0.f);
  const float gx[3] = {fx0 * fx0, 0.f, fx1 * fx1};
  const float gy[3] = {fy0 * fy0, 0.f, fy1 * fy1};
  const float gz[3] = {fz0 * fz0, 0.f, fz1 * fz1};
  // Row by row, not cell by cell: the three cells of a (y,z) row are one contiguous slot range, and every scan is a
  // dependent round trip (load -> compare -> prune test of the next one), so 9 steps instead of up to 27.  The own row
  // first, then the four rows sharing a face with it, then the four diagonal ones; a row is skipped when even its
  // nearest point is too far, and trimmed to the cells that can still hold a closer point (the middle cell has the
  // smallest bound, so the kept cells are contiguous).
  nn1_group_scan<LPP>(g, v[4][0], v[4][3], sub, gmask, qx, qy, qz, best_d, best_p);
  constexpr int kRowOrder[8] = {1, 3, 5, 7, 0, 2, 6, 8};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = kRowOrder[k];
    const float dyz = gy[r % 3] + gz[r / 3];
    const float lim = fminf(best_d, cap_d2);
    if (dyz * 0.999999f < lim) {
      const int a = (dyz + gx[0]) * 0.999999f < lim ? v[r][0] : v[r][1];
      const int b = (dyz + gx[2]) * 0.999999f < lim ? v[r][3] : v[r][2];
      nn1_group_scan<LPP>(g, a, b, sub, gmask, qx, qy, qz, best_d, best_p);
    }
  }
  if (rmax <= 1) return false;
  const float m = cube_face_distance(gp, cx, cy, cz, 1, qx, qy, qz);
  if (m > 0.f) {
    const float m2 = m * m * 0.999999f;
    if (cap_d2 <= m2 || best_d <= m2) return false;
  }
  return true;
}

__device__ __forceinline__ bool grid_nn1_cube1(const GridView& g, const GridParams& gp, float qx, float qy, float qz,
                                               float cap_d2, float& best_d, int& best_p) {
  return grid_nn1_cube1_group<1>(g, gp, 0, 0u, qx, qy, qz, cap_d2, best_d, best_p);
}

// further growth by the same thread (only when the cell edge is smaller than the cap, or the cap is unbounded):
// cube s0 -> s0+1, rows taken 8 at a time so that their table reads overlap, rows and side runs whose nearest point
// is already farther than min(best, cap) skipped without touching memory
__device__ __forceinline__ void grid_nn1_grow_thread(const GridView& g, const GridParams& gp, float qx, float qy, float qz,
                                                     float cap_d2, float& best_d, int& best_p) {
  const int cx = cell_coord(qx, gp.ox, gp.inv, gp.dx);
  const int cy = cell_coord(qy, gp.oy, gp.inv, gp.dy);
  const int cz = cell_coord(qz, gp.oz, gp.inv, gp.dz);
  const int rmax = max(max(max(cx, gp.dx - 1 - cx), max(cy, gp.dy - 1 - cy)), max(cz, gp.dz - 1 - cz));
  const float fy0 = fmaxf(qy - (gp.oy + (float)cy * gp.cell) - gp.margin, 0.f);
  const float fy1 = fmaxf((gp.oy + (float)(cy + 1) * gp.cell) - qy - gp.margin, 0.f);
  const float fz0 = fmaxf(qz - (gp.oz + (float)cz * gp.cell) - gp.margin, 0.f);
  const float fz1 = fmaxf((gp.oz + (float)(cz + 1) * gp.cell) - qz - gp.margin, 0.f);
  const float fx0 = fmaxf(qx - (gp.ox + (float)cx * gp.cell) - gp.margin, 0.f);
  const float fx1 = fmaxf((gp.ox + (float)(cx + 1) * gp.cell) - qx - gp.margin, 0.f);
  for (int s0 = 1; s0 < rmax; ++s0) {
    const float m = cube_face_distance(gp, cx, cy, cz, s0, qx, qy, qz);
    if (m > 0.f) {
      const float m2 = m * m * 0.999999f;
      if (cap_d2 <= m2 || best_d <= m2) break;
    }
    const int s1 = s0 + 1;
    const int w = 2 * s1 + 1, nrows = w * w;
    const int xlo = max(cx - s1, 0), xhi = min(cx + s1, gp.dx - 1);
    // lower bounds of the two side runs of an inner row (x distance to the cells cx -/+ s1)
    const float sxl = fx0 + (float)(s1 - 1) * gp.cell, sxr = fx1 + (float)(s1 - 1) * gp.cell;
    for (int base = 0; base < nrows; base += 8) {
      int a1[8], b1[8], a2[8], b2[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a1[u] = b1[u] = a2[u] = b2[u] = 0;
        const int r = base + u;
        if (r >= nrows) continue;
        const int rz = r / w;
        const int yy = r - rz * w - s1, zz = rz - s1;
        const int y = cy + yy, z = cz + zz;
        if (y < 0 || y >= gp.dy || z < 0 || z >= gp.dz) continue;
        const float dy = yy == 0 ? 0.f : (yy < 0 ? fy0 + (float)(-yy - 1) * gp.cell : fy1 + (float)(yy - 1) * gp.cell);
        const float dz = zz == 0 ? 0.f : (zz < 0 ? fz0 + (float)(-zz - 1) * gp.cell : fz1 + (float)(zz - 1) * gp.cell);
        const float lim = fminf(best_d, cap_d2);
        const float dyz = (dy * dy + dz * dz) * 0.999999f;
        if (dyz >= lim) continue;
        const int* row = g.cell_start + (z * gp.dy + y) * gp.dx;
        if (max(abs(yy), abs(zz)) == s1) {
          a1[u] = __ldg(row + xlo);
          b1[u] = __ldg(row + xhi + 1);
        } else {
          if (cx - s1 >= 0 && dyz + sxl * sxl * 0.999999f < lim) { a1[u] = __ldg(row + cx - s1); b1[u] = __ldg(row + cx - s1 + 1); }
          if (cx + s1 < gp.dx && dyz + sxr * sxr * 0.999999f < lim) { a2[u] = __ldg(row + cx + s1); b2[u] = __ldg(row + cx + s1 + 1); }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        nn1_scan_range(g, a1[u], b1[u], qx, qy, qz, best_d, best_p);
        nn1_scan_range(g, a2[u], b2[u], qx, qy, qz, best_d, best_p);
      }
    }
  }
}

__device__ __forceinline__ void grid_nn1_thread(const GridView& g, const GridParams& gp, float qx, float qy, float qz,
                                                float cap_d2, float& best_d, int& best_p) {
  if (grid_nn1_cube1(g, gp, qx, qy, qz, cap_d2, best_d, best_p)) grid_nn1_grow_thread(g, gp, qx, qy, qz, cap_d2, best_d, best_p);
}

}  // namespace ngicp
