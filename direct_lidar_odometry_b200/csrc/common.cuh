// common.cuh — shared device structures and helpers for libnanogicp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>

namespace ngicp {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Uniform search grid over one cloud ("the index" that replaces nanoflann's kd-tree,
// reference include/nano_gicp/nanoflann.hpp:132-138).  Lives in HBM; written by
// grid_setup_kernel so that no host round trip is needed between bbox and binning.
// ---------------------------------------------------------------------------------------------
struct GridDesc {
  // bounding box as order-preserving uint encodings of floats (atomicMin/Max friendly)
  unsigned bb_min[3];
  unsigned bb_max[3];
  float origin[3];
  float cell;
  float inv_cell;
  int dim[3];
  int ncells;      // dim[0]*dim[1]*dim[2]
  int n;           // points binned
  int max_dim;
  float margin;    // conservative slack (metres) used by the exact ring search stop test
  // voxel-filter fields (pcl::VoxelGrid): min_b, div_b, overflow flag, output count
  int vmin_b[3];
  int vdiv[3];
  int voverflow;
  int vcount;
  int nfinite;
  unsigned long long occ_sq;  // sum over points of the population of their (trial) cell
};

// Read-only view handed to search kernels by value.
struct GridView {
  const int* __restrict__ cell_start;   // ncells+1 entries: first sorted slot with key >= c
  const float4* __restrict__ sorted;    // points in cell order; .w = bit pattern of the ORIGINAL index
  const GridDesc* __restrict__ desc;
};

__host__ __device__ inline unsigned f2ord(float f) {
  unsigned u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  union { float f; unsigned u; } c; c.f = f; u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; unsigned u; } c; c.u = u; return c.f;
#endif
}

// nanoflann's metric (nanoflann_impl.hpp:441-449): d = 0; d += dx*dx; d += dy*dy; d += dz*dz,
// every operation rounded to float, never fused.
__device__ __forceinline__ float sqdist_unfused(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// cell coordinate of a point along one axis; identical formula for binning and for queries
__device__ __forceinline__ int cell_coord(float p, float origin, float inv_cell, int dim) {
  int c = (int)floorf((p - origin) * inv_cell);
  return min(max(c, 0), dim - 1);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// Sum NV (<= 32) per-lane values across the warp at once: at every butterfly step a lane keeps one half of its values,
// hands the other half to its partner and adds what it receives, so 32 + 16 + ... = 31 exchanges do the work of
// NV x 5.  Afterwards lane j holds the warp total of value j in v[0] (j < NV).  Fixed summation tree: deterministic.
template <int NV>
__device__ __forceinline__ double warp_sum_transposed(const double* acc) {
  const int lane = threadIdx.x & 31;
  double v[32];
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = j < NV ? acc[j] : 0.0;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; j++) {
      if (j >= NV) continue;                       // both halves are padding: nothing to move
      const bool upper_real = j + half < NV;       // compile-time after unrolling
      const double lo = v[j], hi = upper_real ? v[j + half] : 0.0;
      const double send = up ? lo : hi;
      const double keep = up ? hi : lo;
      v[j] = keep + __shfl_xor_sync(FULL, send, half);
    }
  }
  return v[0];
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

}  // namespace ngicp
