// sort_scan.cu — hand-written device-wide primitives used by the grid index and the voxel reduce:
//   * exclusive prefix sum over int32 (three-kernel tile scan, length read from device memory)
//   * stable LSD radix sort of (uint32 key, uint32 value) pairs, 9 bits per pass
// Stability matters: it makes "ascending original index inside a cell/voxel" the deterministic
// within-bucket order (pcl::VoxelGrid's own sort is unstable).
#include <cstdint>
#include "internal.h"

namespace ngicp {

// ------------------------------------------------------------------------------------------
// exclusive scan
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;  // 2048

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* smem /* 8 ints */, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  int wsum = (lane < 8) ? smem[lane] : 0;
  int winc = wsum;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) { int t = __shfl_up_sync(FULL, winc, o); if (lane >= o) winc += t; }
  const int woff = __shfl_sync(FULL, winc - wsum, w);
  total = __shfl_sync(FULL, winc, 7);
  __syncthreads();
  return woff + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int* __restrict__ data, const int* __restrict__ n_ptr, int n_add, int* __restrict__ tile_sums) {
  __shared__ int sm[8];
  const int n = *n_ptr + n_add;
  const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  const bool vec = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int base = t * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int s = 0;
    if (vec && base + SCAN_IPT <= n) {
      const int4 a = *reinterpret_cast<const int4*>(data + base), b = *reinterpret_cast<const int4*>(data + base + 4);
      s = (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
    } else {
#pragma unroll
      for (int j = 0; j < SCAN_IPT; j++) s += (base + j < n) ? data[base + j] : 0;
    }
    int total;
    block_exclusive_scan_256(s, sm, total);
    if (threadIdx.x == 0) tile_sums[t] = total;
  }
}

__global__ void __launch_bounds__(1024) scan_tile_offsets(int* __restrict__ tile_sums, const int* __restrict__ n_ptr, int n_add) {
  // single block: exclusive scan of tile_sums[0..ntiles)
  __shared__ int sm[32];
  __shared__ int carry_s;
  const int n = *n_ptr + n_add;
  const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < ntiles ? tile_sums[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
      int ws = sm[lane], wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
      sm[lane] = wi - ws;
      if (lane == 31) sm[31] = wi - ws;  // keep exclusive; total handled below
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + sm[w] + inc - v;
    if (i < ntiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(int* __restrict__ data, const int* __restrict__ n_ptr, int n_add, const int* __restrict__ tile_sums) {
  __shared__ int sm[8];
  const int n = *n_ptr + n_add;
  const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  const bool vec = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int base = t * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int v[SCAN_IPT];
    int s = 0;
    const bool full = vec && base + SCAN_IPT <= n;
    if (full) {
      const int4 a = *reinterpret_cast<const int4*>(data + base), b = *reinterpret_cast<const int4*>(data + base + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
      for (int j = 0; j < SCAN_IPT; j++) s += v[j];
    } else {
#pragma unroll
      for (int j = 0; j < SCAN_IPT; j++) { v[j] = (base + j < n) ? data[base + j] : 0; s += v[j]; }
    }
    int total;
    int off = block_exclusive_scan_256(s, sm, total) + tile_sums[t];
    if (full) {
      int o[SCAN_IPT];
#pragma unroll
      for (int j = 0; j < SCAN_IPT; j++) { o[j] = off; off += v[j]; }
      *reinterpret_cast<int4*>(data + base) = make_int4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<int4*>(data + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int j = 0; j < SCAN_IPT; j++) { if (base + j < n) data[base + j] = off; off += v[j]; }
    }
  }
}

void exclusive_scan_inplace(int* data, const int* n_dev, int n_add, int max_n, int* tile_sums, cudaStream_t st) {
  int ntiles = (max_n + SCAN_TILE - 1) / SCAN_TILE;
  int grid = ntiles < 148 * 8 ? ntiles : 148 * 8;
  if (grid < 1) grid = 1;
  scan_tile_sums<<<grid, SCAN_THREADS, 0, st>>>(data, n_dev, n_add, tile_sums);
  scan_tile_offsets<<<1, 1024, 0, st>>>(tile_sums, n_dev, n_add);
  scan_apply<<<grid, SCAN_THREADS, 0, st>>>(data, n_dev, n_add, tile_sums);
  note_launches(3);
}

size_t scan_scratch_ints(int max_n) { return (size_t)(max_n + SCAN_TILE - 1) / SCAN_TILE + 1; }

// ------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BITS = 9;                    // 9-bit digits: 27-bit keys (cell / voxel indices) need 3 passes
constexpr int RS_BINS = 1 << RS_BITS;
constexpr unsigned RS_MASK = RS_BINS - 1;

// per-block digit histogram, digit-major layout hist[d * nblk + blk].  Warps count with match_any (no contended
// shared-memory atomics: voxel / cell keys arrive nearly sorted, so neighbouring items share their digit).
template <int IPT>
__global__ void __launch_bounds__(RS_THREADS) rs_histogram(const unsigned* __restrict__ keys, int n, int shift, int nblk, int* __restrict__ hist) {
  __shared__ int cnt[RS_BINS];
  for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) cnt[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wbase = blockIdx.x * (RS_THREADS * IPT) + w * (32 * IPT);
#pragma unroll
  for (int j = 0; j < IPT; j++) {
    const int i = wbase + j * 32 + lane;
    const bool valid = i < n;
    const unsigned d = valid ? ((keys[i] >> shift) & RS_MASK) : (unsigned)RS_BINS;
    const unsigned peers = __match_any_sync(FULL, d);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&cnt[d], __popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RS_BINS; i += RS_THREADS) hist[i * nblk + blockIdx.x] = cnt[i];
}

// exclusive scan of the digit-major histogram (RS_BINS*nblk ints) by a single block
__global__ void __launch_bounds__(1024) rs_scan_hist(int* __restrict__ hist, int total) {
  __shared__ int sm[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  constexpr int PER = 4;
  for (int base = 0; base < total; base += 1024 * PER) {
    const int i0 = base + threadIdx.x * PER;
    int v[PER];
    int s = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) { v[j] = (i0 + j < total) ? hist[i0 + j] : 0; s += v[j]; }
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
      int ws = sm[lane], wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
      sm[lane] = wi - ws;
    }
    __syncthreads();
    int off = carry_s + sm[w] + inc - s;
#pragma unroll
    for (int j = 0; j < PER; j++) { if (i0 + j < total) hist[i0 + j] = off; off += v[j]; }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = off;
    __syncthreads();
  }
}

// stable scatter: each warp owns 32*IPT contiguous items and ranks them in order with match_any
template <int IPT>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                                                         unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                                                         int n, int shift, int nblk, const int* __restrict__ hist) {
  __shared__ int wcnt[RS_WARPS][RS_BINS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int wbase = blockIdx.x * (RS_THREADS * IPT) + w * (32 * IPT);
  unsigned key[IPT], val[IPT];
  int rank[IPT];
  const unsigned lt = (1u << lane) - 1u;
  // pass 1: per-warp digit counts in item order; remember each item's rank among equal digits of its warp
#pragma unroll
  for (int j = 0; j < IPT; j++) {
    const int i = wbase + j * 32 + lane;
    const bool valid = i < n;
    key[j] = valid ? keys_in[i] : 0u;
    val[j] = valid ? vals_in[i] : 0u;
    const unsigned d = valid ? ((key[j] >> shift) & RS_MASK) : (unsigned)RS_BINS;
    const unsigned peers = __match_any_sync(FULL, d);
    const int leader = __ffs(peers) - 1;
    int before = 0;
    if (valid && lane == leader) { before = wcnt[w][d]; wcnt[w][d] = before + __popc(peers); }
    before = __shfl_sync(FULL, before, leader);
    rank[j] = before + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  // exclusive offsets: global digit base for this block, then earlier warps of this block
  for (int d = threadIdx.x; d < RS_BINS; d += RS_THREADS) {
    int run = hist[d * nblk + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++) { const int t = wcnt[ww][d]; wcnt[ww][d] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < IPT; j++) {
    const int i = wbase + j * 32 + lane;
    if (i < n) {
      const unsigned d = (key[j] >> shift) & RS_MASK;
      const int pos = wcnt[w][d] + rank[j];
      keys_out[pos] = key[j];
      vals_out[pos] = val[j];
    }
  }
}

static inline int rs_ipt(int n) { return n <= (1 << 18) ? 4 : 16; }

size_t radix_sort_scratch_ints(int n) {
  const int tile = RS_THREADS * rs_ipt(n);
  int nblk = (n + tile - 1) / tile;
  if (nblk < 1) nblk = 1;
  return (size_t)RS_BINS * nblk;
}

// Sorts (keys_a, vals_a) using (keys_b, vals_b) as the ping-pong partner.  `bits` = number of
// significant key bits (rounded up to whole 9-bit passes).  Returns 0 if the sorted data ends in
// the *_a buffers, 1 if in *_b.
int radix_sort_pairs(unsigned* keys_a, unsigned* vals_a, unsigned* keys_b, unsigned* vals_b, int n, int bits, int* hist, cudaStream_t st) {
  if (n <= 0) return 0;
  const int ipt = rs_ipt(n);
  const int tile = RS_THREADS * ipt;
  const int nblk = (n + tile - 1) / tile;
  const int passes = (bits + RS_BITS - 1) / RS_BITS;
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    const unsigned* kin = cur ? keys_b : keys_a;
    const unsigned* vin = cur ? vals_b : vals_a;
    unsigned* kout = cur ? keys_a : keys_b;
    unsigned* vout = cur ? vals_a : vals_b;
    if (ipt == 4) rs_histogram<4><<<nblk, RS_THREADS, 0, st>>>(kin, n, p * RS_BITS, nblk, hist);
    else rs_histogram<16><<<nblk, RS_THREADS, 0, st>>>(kin, n, p * RS_BITS, nblk, hist);
    rs_scan_hist<<<1, 1024, 0, st>>>(hist, RS_BINS * nblk);
    if (ipt == 4) rs_scatter<4><<<nblk, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, p * RS_BITS, nblk, hist);
    else rs_scatter<16><<<nblk, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, p * RS_BITS, nblk, hist);
    note_launches(3);
    cur ^= 1;
  }
  return cur;
}

}  // namespace ngicp
