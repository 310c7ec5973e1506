// cloud_index.cu — K1: snapshot a caller cloud into HBM and build the uniform-grid search index that
// replaces nanoflann's serial kd-tree build (reference include/nano_gicp/nanoflann.hpp:132-138 ->
// impl/nanoflann_impl.hpp:1199-1211, divideTree :867-917).
//
// Layout produced (all in HBM, SURVEY §8d "canonical device layout"):
//   pts[n]        float4 {x,y,z,1}            original order
//   sorted[n]     float4 {x,y,z,bits(orig)}   ordered by cell key (x fastest, then y, then z), ties by original index
//   cell_start[]  int32, ncells+1 entries     cell_start[c] = first sorted slot whose key >= c  (so any run of
//                                              x-adjacent cells is ONE contiguous slot range)
//   desc          GridDesc                    origin, cell edge, dims — computed on the device, no host round trip
// Algorithmic bytes: 16 n read + 16 n sorted write + 4 n permutation = 36 B/point.
#include <cstdlib>
#include <mutex>
#include <vector>
#include <cstdio>
#include <ctime>
#include <cstdint>
#include <cstring>
#include <cooperative_groups.h>
#include "internal.h"

namespace ngicp {

// The library's own stream-ordered memory pool (one per device, created on first use, never trimmed): per-scan buffers are
// pointer bumps, and the embedding application's DEFAULT pool keeps the release threshold it chose (round 1 raised it).
cudaError_t pool_alloc_async(void** p, size_t nbytes, cudaStream_t s) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  static bool tried[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool = nullptr;
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lock(mu);
    if (!tried[dev]) {
      tried[dev] = true;
      cudaMemPoolProps props;
      memset(&props, 0, sizeof props);
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      if (cudaMemPoolCreate(&pools[dev], &props) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &thr);
      } else {
        cudaGetLastError();
        pools[dev] = nullptr;       // (driver without pool support: the default pool, as before)
      }
    }
    pool = pools[dev];
  }
  return pool ? cudaMallocFromPoolAsync(p, nbytes, pool, s) : cudaMallocAsync(p, nbytes, s);
}

cudaError_t DevBuf::alloc(size_t nbytes, const StreamPtr& stream) {
  release();
  if (nbytes == 0) nbytes = 16;
  st = stream;
  cudaError_t e;
  static const bool trace = getenv("NGICP_HOST_TRACE") != nullptr;
  timespec t0, t1;
  if (trace) clock_gettime(CLOCK_MONOTONIC, &t0);
  if (st && st->owned) e = pool_alloc_async(&p, nbytes, st->s);
  else e = cudaMalloc(&p, nbytes);
  if (trace) {
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
    if (ms > 0.5) fprintf(stderr, "[ngicp] slow alloc %.2f ms for %zu bytes\n", ms, nbytes);
  }
  if (e != cudaSuccess) { p = nullptr; bytes = 0; return e; }
  bytes = nbytes;
  return cudaSuccess;
}
cudaError_t DevBuf::reserve(size_t nbytes, const StreamPtr& stream) {
  if (p && bytes >= nbytes) return cudaSuccess;
  // grow geometrically so that streams of slightly different cloud sizes do not reallocate every call
  size_t want = nbytes + nbytes / 4 + 256;
  return alloc(want, stream);
}
void DevBuf::release() {
  if (!p) return;
  if (st && st->owned) cudaFreeAsync(p, st->s);
  else cudaFree(p);
  p = nullptr;
  bytes = 0;
}

namespace {
struct TableEntry { void* p; size_t bytes; int device; cudaEvent_t ready; };
std::mutex g_table_mutex;
std::vector<TableEntry> g_table_free;
constexpr size_t TABLE_CACHE_MAX = 12;
}  // namespace

cudaError_t TableBuf::acquire(size_t nbytes, int dev, const StreamPtr& stream) {
  release();
  st = stream;
  device = dev;
  TableEntry got{nullptr, 0, 0, nullptr};
  {
    std::lock_guard<std::mutex> lock(g_table_mutex);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < g_table_free.size(); i++) {
      const TableEntry& e = g_table_free[i];
      if (e.device == dev && e.bytes >= nbytes && (best == (size_t)-1 || e.bytes < g_table_free[best].bytes)) best = i;
    }
    if (best != (size_t)-1) { got = g_table_free[best]; g_table_free.erase(g_table_free.begin() + best); }
  }
  if (got.p) {
    // everything enqueued on the previous owner's stream before the release must finish before this stream writes
    if (got.ready) { cudaStreamWaitEvent(st->s, got.ready, 0); cudaEventDestroy(got.ready); }
    p = got.p;
    bytes = got.bytes;
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(&p, nbytes);
  if (e != cudaSuccess) { p = nullptr; bytes = 0; return e; }
  bytes = nbytes;
  return cudaSuccess;
}

void TableBuf::release() {
  if (!p) return;
  TableEntry e{p, bytes, device, nullptr};
  if (st && st->s && cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(e.ready, st->s);
  bool keep = false;
  {
    std::lock_guard<std::mutex> lock(g_table_mutex);
    if (g_table_free.size() < TABLE_CACHE_MAX) { g_table_free.push_back(e); keep = true; }
  }
  if (!keep) {
    if (e.ready) { cudaEventSynchronize(e.ready); cudaEventDestroy(e.ready); }
    cudaFree(p);
  }
  p = nullptr;
  bytes = 0;
}

__global__ void desc_init_kernel(GridDesc* d) {
  for (int i = 0; i < 3; i++) { d->bb_min[i] = f2ord(FLT_MAX); d->bb_max[i] = f2ord(-FLT_MAX); }
  d->ncells = 0; d->n = 0; d->nfinite = 0; d->vcount = 0; d->voverflow = 0;
}

// raw records (x,y,z at the start of each stride) -> float4, fused with the bounding-box reduction
__global__ void __launch_bounds__(256) pack_bbox_kernel(const unsigned char* __restrict__ raw, size_t stride, int n,
                                                        float4* __restrict__ pts, GridDesc* __restrict__ d) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int finite = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* r = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    float x, y, z;
    if ((stride & 15) == 0) { const float4 v = *reinterpret_cast<const float4*>(r); x = v.x; y = v.y; z = v.z; }
    else { x = r[0]; y = r[1]; z = r[2]; }
    pts[i] = make_float4(x, y, z, 1.0f);
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      finite++;
      mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
      mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) finite += __shfl_xor_sync(FULL, finite, o);
  // one set of atomics per BLOCK (7 same-address atomics per warp serialise in L2: 109k of them for a 500k-point cloud)
  __shared__ float s_mn[8][3], s_mx[8][3];
  __shared__ int s_fin[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { s_mn[w][a] = mn[a]; s_mx[w][a] = mx[a]; }
    s_fin[w] = finite;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int ww = 1; ww < nw; ww++) {
#pragma unroll
      for (int a = 0; a < 3; a++) { mn[a] = fminf(mn[a], s_mn[ww][a]); mx[a] = fmaxf(mx[a], s_mx[ww][a]); }
      finite += s_fin[ww];
    }
    if (finite > 0) {
#pragma unroll
      for (int a = 0; a < 3; a++) { atomicMin(&d->bb_min[a], f2ord(mn[a])); atomicMax(&d->bb_max[a], f2ord(mx[a])); }
      atomicAdd(&d->nfinite, finite);
    }
  }
}

// derive the grid from the bounding box; grow the cell until the dense table fits
struct GridShape { float origin[3]; float cell, inv_cell, margin; int dim[3]; int ncells, max_dim; };
__device__ void grid_shape_compute(const float* lo_in, const float* hi_in, int nfinite, float cell, int cap, GridShape& gs);
__device__ void grid_setup_device(GridDesc* d, float cell, int cap, int n) {
  float lo[3], hi[3];
  for (int a = 0; a < 3; a++) { lo[a] = ord2f(d->bb_min[a]); hi[a] = ord2f(d->bb_max[a]); }
  GridShape gs;
  grid_shape_compute(lo, hi, d->nfinite, cell, cap, gs);
  for (int a = 0; a < 3; a++) { d->origin[a] = gs.origin[a]; d->dim[a] = gs.dim[a]; }
  d->cell = gs.cell;
  d->inv_cell = gs.inv_cell;
  d->ncells = gs.ncells;
  d->n = n;
  d->max_dim = gs.max_dim;
  d->margin = gs.margin;
  d->occ_sq = 0ull;
}
__device__ void grid_shape_compute(const float* lo_in, const float* hi_in, int nfinite, float cell, int cap, GridShape& gs) {
  float lo[3], hi[3];
  for (int a = 0; a < 3; a++) { lo[a] = lo_in[a]; hi[a] = hi_in[a]; }
  if (nfinite == 0) { for (int a = 0; a < 3; a++) { lo[a] = 0.f; hi[a] = 0.f; } }
  int dim[3];
  // Grow the cell edge until the dense table fits.  The extent can be anything a float holds (DLO only strips NaN/Inf:
  // one stray 1e20 return is a legal input), so the loop runs until it fits — 1.25^400 spans the whole float range —
  // and when even that fails (extent overflowing to infinity) the grid degenerates to ONE cell: slow but correct, and
  // dim[] can never exceed the table the following kernels write into.
  bool fits = false;
  for (int it = 0; it < 400 && !fits; it++) {
    double prod = 1.0;
    const float inv = 1.0f / cell;
    for (int a = 0; a < 3; a++) {
      float ext = (hi[a] - lo[a]) * inv;
      if (!(ext < 2.0e9f)) ext = 2.0e9f;
      dim[a] = (int)floorf(ext) + 1;
      prod *= (double)dim[a];
    }
    fits = prod <= (double)cap;
    if (!fits) {
      // jump most of the way at once (a volume ratio r needs a factor cbrt(r) on compact clouds, at most r on
      // degenerate ones), then the +25 % steps finish; never past the float range
      const float jump = (float)fmin(cbrt(prod / (double)cap), 1.0e6);
      cell = fminf(cell * fmaxf(1.25f, jump), 1.0e37f);
    }
  }
  if (!fits) {
    cell = 3.0e38f;
    for (int a = 0; a < 3; a++) dim[a] = 1;
  }
  for (int a = 0; a < 3; a++) { gs.origin[a] = lo[a]; gs.dim[a] = dim[a]; }
  gs.cell = cell;
  gs.inv_cell = 1.0f / cell;
  gs.ncells = dim[0] * dim[1] * dim[2];
  gs.max_dim = max(dim[0], max(dim[1], dim[2]));
  gs.margin = cell * (0.01f + 1e-6f * (float)gs.max_dim);
}

__global__ void grid_setup_kernel(GridDesc* d, float cell, int cap, int n) { grid_setup_device(d, cell, cap, n); }

// Automatic cell edge.  A trial grid (edge c0) has been histogrammed; occ = sum n_c^2 / n is the mean number of
// points sharing a cell with a point.  LiDAR clouds are surface-like (n_c ~ c^2), so the edge that gives the
// wanted occupancy is c0 * sqrt(target / occ).
__global__ void grid_autocell_kernel(GridDesc* d, float c0, float target_occ, int cap, int n) {
  float cell = c0;
  if (n > 0 && d->occ_sq > 0ull) {
    const float occ = (float)((double)d->occ_sq / (double)n);
    float f = sqrtf(target_occ / occ);
    f = fminf(fmaxf(f, 0.125f), 2.0f);
    cell = c0 * f;
  }
  grid_setup_device(d, cell, cap, n);
}

// trial histogram only (no keys)
__global__ void __launch_bounds__(256) count_points_kernel(const float4* __restrict__ pts, int n, const GridDesc* __restrict__ d, int* __restrict__ cell_count) {
  const float ox = d->origin[0], oy = d->origin[1], oz = d->origin[2], inv = d->inv_cell;
  const int dx = d->dim[0], dy = d->dim[1], dz = d->dim[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    const int cx = cell_coord(p.x, ox, inv, dx), cy = cell_coord(p.y, oy, inv, dy), cz = cell_coord(p.z, oz, inv, dz);
    atomicAdd(&cell_count[(cz * dy + cy) * dx + cx], 1);
  }
}

// sum over points of the population of their trial cell
__global__ void __launch_bounds__(256) occupancy_kernel(const float4* __restrict__ pts, int n, GridDesc* __restrict__ d, const int* __restrict__ cell_count) {
  const float ox = d->origin[0], oy = d->origin[1], oz = d->origin[2], inv = d->inv_cell;
  const int dx = d->dim[0], dy = d->dim[1], dz = d->dim[2];
  unsigned long long s = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    const int cx = cell_coord(p.x, ox, inv, dx), cy = cell_coord(p.y, oy, inv, dy), cz = cell_coord(p.z, oz, inv, dz);
    s += (unsigned long long)cell_count[(cz * dy + cy) * dx + cx];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&d->occ_sq, s);
}

__global__ void __launch_bounds__(256) zero_cells_kernel(int* __restrict__ cell_start, const GridDesc* __restrict__ d) {
  const int total = d->ncells + 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) cell_start[i] = 0;
}

// histogram for the counting sort: counts of cell c go to table[c + 1] (see build_index)
__global__ void __launch_bounds__(256) bin_points_kernel(const float4* __restrict__ pts, int n, const GridDesc* __restrict__ d,
                                                         unsigned* __restrict__ keys, int* __restrict__ table) {
  const float ox = d->origin[0], oy = d->origin[1], oz = d->origin[2], inv = d->inv_cell;
  const int dx = d->dim[0], dy = d->dim[1], dz = d->dim[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    const int cx = cell_coord(p.x, ox, inv, dx), cy = cell_coord(p.y, oy, inv, dy), cz = cell_coord(p.z, oz, inv, dz);
    const unsigned key = (unsigned)((cz * dy + cy) * dx + cx);
    keys[i] = key;
    atomicAdd(&table[key + 1], 1);
  }
}

// counting-sort scatter: table[c + 1] holds start(c) and is bumped to start(c + 1) by the atomics, which turns
// `table` into the final lower-bound table; the order inside a cell is whatever the atomics produced ...
__global__ void __launch_bounds__(256) scatter_points_kernel(const unsigned* __restrict__ keys, int n, int* __restrict__ table,
                                                             unsigned* __restrict__ slot_orig) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int pos = atomicAdd(&table[keys[i] + 1], 1);
    slot_orig[pos] = (unsigned)i;
  }
}

// ... and this pass makes it deterministic: every point is moved to (cell start + number of points of its cell
// with a smaller original index), i.e. cells are ordered by original index exactly like a stable sort would.
// Fused with the gather of the coordinates.  Cells with more than ORDER_FIX_MAX points keep the scatter order
// (only the visiting order of equidistant neighbours could depend on it).
constexpr int ORDER_FIX_MAX = 4096;
__global__ void __launch_bounds__(256) order_gather_kernel(const float4* __restrict__ pts, const unsigned* __restrict__ keys,
                                                           const unsigned* __restrict__ slot_orig, int n, const int* __restrict__ table,
                                                           float4* __restrict__ sorted) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const unsigned o = slot_orig[p];
    const unsigned key = keys[o];
    const int a = table[key], b = table[key + 1];
    int dst = p;
    if (b - a <= ORDER_FIX_MAX) {
      int rank = 0;
      for (int j = a; j < b; ++j) rank += (slot_orig[j] < o) ? 1 : 0;
      dst = a + rank;
    }
    const float4 v = pts[o];
    sorted[dst] = make_float4(v.x, v.y, v.z, __uint_as_float(o));
  }
}

static inline int grid_for(int n, int threads = 256, int max_blocks = 148 * 8) {
  int g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  return g > max_blocks ? max_blocks : g;
}

cudaError_t host_source_consumed(const void* src, Scratch& sc, cudaStream_t st) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, src) != cudaSuccess) { cudaGetLastError(); return cudaSuccess; }   // plain pageable memory
  if (attr.type != cudaMemoryTypeHost) return cudaSuccess;
  cudaError_t e;
  if (!sc.copy_done && (e = cudaEventCreateWithFlags(&sc.copy_done, cudaEventDisableTiming)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(sc.copy_done, st)) != cudaSuccess) return e;
  return cudaEventSynchronize(sc.copy_done);
}

cudaError_t upload_cloud(DevCloud& c, const void* pts, size_t n, size_t stride_bytes, Scratch& sc, const StreamPtr& st) {
  cudaError_t e;
  c.n = (int)n;
  c.indexed = false;
  if ((e = c.pts.alloc(sizeof(float4) * (n ? n : 1), st)) != cudaSuccess) return e;
  if ((e = c.desc.alloc(sizeof(GridDesc), st)) != cudaSuccess) return e;
  desc_init_kernel<<<1, 1, 0, st->s>>>(c.desc.as<GridDesc>());
  note_launches(1);
  if (n == 0) return cudaGetLastError();
  const size_t raw_bytes = (n - 1) * stride_bytes + 12;  // last record may be shorter than the stride
  // records already in device memory (16-byte aligned) are read in place; host records go through a staging copy
  const unsigned char* raw = nullptr;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, pts) == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) &&
      (reinterpret_cast<uintptr_t>(pts) & 15) == 0 && (stride_bytes & 15) == 0 && stride_bytes >= 16) {
    raw = static_cast<const unsigned char*>(pts);
  } else {
    cudaGetLastError();  // clear a possible "invalid value" from probing a plain host pointer
    if ((e = sc.staging.reserve(raw_bytes + 16, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(sc.staging.p, pts, raw_bytes, cudaMemcpyDefault, st->s)) != cudaSuccess) return e;
    if ((e = host_source_consumed(pts, sc, st->s)) != cudaSuccess) return e;
    raw = sc.staging.as<unsigned char>();
  }
  pack_bbox_kernel<<<grid_for((int)n), 256, 0, st->s>>>(raw, stride_bytes, (int)n, c.pts.as<float4>(), c.desc.as<GridDesc>());
  note_launches(1);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// K1 in ONE persistent cooperative launch: snapshot + bounding box -> (trial histogram -> occupancy -> cell edge) ->
// zero the table -> histogram -> exclusive scan -> counting-sort scatter -> deterministic in-cell order + gather, separated by
// grid barriers instead of the 14 kernel boundaries of upload_cloud + build_index below (a 22k-point scan spends 80 us
// there on launch latency alone; its actual work is a few microseconds).  Every block derives the grid shape itself from
// the per-block bounding boxes (same arithmetic as grid_setup_device), block 0 writes the descriptor.  The table scan is
// slice-per-block: sums, barrier, every block adds up the slices before its own and scans its slice.
// Same results as the multi-kernel path (the in-cell order is made deterministic by the same last pass).
// ---------------------------------------------------------------------------------------------------------------
constexpr int IF_THREADS = 512;
constexpr int IF_WARPS = IF_THREADS / 32;
struct IndexFusedArgs {
  const unsigned char* raw; size_t stride; int n;
  float4* pts; float4* sorted; int* table; GridDesc* d;
  unsigned* keys; unsigned* slot_orig;
  float* part;               // [G][8]: per-block bbox + finite count
  unsigned long long* occ;   // [G]: per-block occupancy sums
  int* slice_sum;            // [G]
  unsigned* bar;             // grid barrier words, zero between launches
  float cell_req, target_occ; int table_cap, trial_cap;
  // {cell edge, point count, builds since the last trial histogram} of the previous cloud indexed through this scratch
  // (device memory, written by block 0 at the end): a cloud of about the same size re-uses that cell edge and skips the
  // trial histogram (three barriers and two passes) — consecutive scans / submaps of a stream look alike, and the search
  // is exact for ANY cell edge (it only costs speed); every 16th build measures again
  float* hint;
  unsigned char* bricks; long long brick_cap;     // brick occupancy flags to fill (nullptr: none), see KNN_BRICK
};

__device__ __forceinline__ void if_grid_barrier(unsigned* bar, unsigned& phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = (phase + 1u) * gridDim.x;
    unsigned v;
    do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
    __threadfence();
  }
  phase++;
  __syncthreads();
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, `total` = block sum
__device__ __forceinline__ int if_block_exclusive_scan(int v, int* scan /* IF_WARPS + 1 */, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) scan[w] = inc;
  __syncthreads();
  if (w == 0) {
    const int ws = lane < IF_WARPS ? scan[lane] : 0;
    int wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
    if (lane < IF_WARPS) scan[lane] = wi - ws;
    if (lane == 31) scan[IF_WARPS] = wi;
  }
  __syncthreads();
  const int r = scan[w] + inc - v;
  total = scan[IF_WARPS];
  __syncthreads();
  return r;
}

// BAR: grid barrier of the cooperative launch (GridBar) or hardware barrier of a thread-block cluster (ClusterBar: one
// cluster of IF_CLUSTER blocks builds one small cloud — an ordinary launch, which many streams can interleave)
struct GridBar {
  unsigned* bar; unsigned phase;
  __device__ __forceinline__ void sync() { if_grid_barrier(bar, phase); }
};
struct ClusterBar {
  __device__ __forceinline__ void sync() { __threadfence(); cooperative_groups::this_cluster().sync(); }
};
constexpr int IF_CLUSTER = 8;
template <class BAR>
__device__ __forceinline__ void index_fused_body(const IndexFusedArgs& a, const int G, const int blk, BAR& bar) {
  __shared__ float s_red[IF_WARPS][8];
  __shared__ int s_scan[IF_WARPS + 1];
  __shared__ GridShape s_gs;
  __shared__ float s_lo[3], s_hi[3];
  __shared__ int s_nfinite, s_carry, s_have_cell, s_used_hint;
  __shared__ unsigned long long s_occ[IF_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gtid = blk * IF_THREADS + tid, gstride = G * IF_THREADS;
  const int n = a.n;

  // ---- snapshot + bounding box of the finite points (pack_bbox_kernel) ----
  {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int finite = 0;
    for (int i = gtid; i < n; i += gstride) {
      const float* r = reinterpret_cast<const float*>(a.raw + (size_t)i * a.stride);
      float x, y, z;
      if ((a.stride & 15) == 0) { const float4 v = *reinterpret_cast<const float4*>(r); x = v.x; y = v.y; z = v.z; }
      else { x = r[0]; y = r[1]; z = r[2]; }
      a.pts[i] = make_float4(x, y, z, 1.0f);
      if (isfinite(x) && isfinite(y) && isfinite(z)) {
        finite++;
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(FULL, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL, mx[c], o));
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) finite += __shfl_xor_sync(FULL, finite, o);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) { s_red[w][c] = mn[c]; s_red[w][3 + c] = mx[c]; }
      s_red[w][6] = __int_as_float(finite);
    }
    __syncthreads();
    if (tid == 0) {
      for (int ww = 1; ww < IF_WARPS; ww++) {
#pragma unroll
        for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], s_red[ww][c]); mx[c] = fmaxf(mx[c], s_red[ww][3 + c]); }
        finite += __float_as_int(s_red[ww][6]);
      }
      float* p = a.part + (size_t)blk * 8;
#pragma unroll
      for (int c = 0; c < 3; c++) { p[c] = mn[c]; p[3 + c] = mx[c]; }
      p[6] = __int_as_float(finite);
    }
  }
  bar.sync();

  // ---- every block: the cloud's bounding box ----
  if (w == 0) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int finite = 0;
    for (int b = lane; b < G; b += 32) {
      const float* p = a.part + (size_t)b * 8;
#pragma unroll
      for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], __ldcg(p + c)); mx[c] = fmaxf(mx[c], __ldcg(p + 3 + c)); }
      finite += __float_as_int(__ldcg(p + 6));
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(FULL, mn[c], o));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL, mx[c], o));
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) finite += __shfl_xor_sync(FULL, finite, o);
    if (lane == 0) {
      // (bb_min / bb_max go through the order-preserving encoding in the multi-kernel path; decode(encode(x)) == x)
      for (int c = 0; c < 3; c++) { s_lo[c] = mn[c]; s_hi[c] = mx[c]; }
      s_nfinite = finite;
      float cell0 = a.cell_req;
      s_used_hint = 0;
      if (!(cell0 > 0.f) && a.hint != nullptr) {
        const float hc = __ldcg(a.hint), hn = __ldcg(a.hint + 1), cnt = __ldcg(a.hint + 2);
        if (hc > 0.f && cnt < 15.5f && (float)n >= 0.8f * hn && (float)n <= 1.25f * hn) { cell0 = hc; s_used_hint = 1; }
      }
      GridShape gs;
      grid_shape_compute(s_lo, s_hi, finite, cell0 > 0.f ? cell0 : 1.0f, cell0 > 0.f ? a.table_cap : a.trial_cap, gs);
      s_gs = gs;
      s_have_cell = cell0 > 0.f ? 1 : 0;
    }
  }
  __syncthreads();

  if (!s_have_cell) {
    // ---- automatic cell edge: histogram of a 1 m trial grid, point-weighted occupancy (count_points / occupancy /
    //      grid_autocell kernels) ----
    {
      const int total = s_gs.ncells + 2;
      for (int i = gtid; i < total; i += gstride) a.table[i] = 0;
    }
    bar.sync();
    const float ox = s_gs.origin[0], oy = s_gs.origin[1], oz = s_gs.origin[2], inv = s_gs.inv_cell;
    const int dx = s_gs.dim[0], dy = s_gs.dim[1], dz = s_gs.dim[2];
    for (int i = gtid; i < n; i += gstride) {
      const float4 p = a.pts[i];
      const int cx = cell_coord(p.x, ox, inv, dx), cy = cell_coord(p.y, oy, inv, dy), cz = cell_coord(p.z, oz, inv, dz);
      atomicAdd(&a.table[(cz * dy + cy) * dx + cx], 1);
    }
    bar.sync();
    unsigned long long so = 0ull;
    for (int i = gtid; i < n; i += gstride) {
      const float4 p = a.pts[i];
      const int cx = cell_coord(p.x, ox, inv, dx), cy = cell_coord(p.y, oy, inv, dy), cz = cell_coord(p.z, oz, inv, dz);
      so += (unsigned long long)__ldcg(&a.table[(cz * dy + cy) * dx + cx]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) so += __shfl_xor_sync(FULL, so, o);
    if (lane == 0) s_occ[w] = so;
    __syncthreads();
    if (tid == 0) {
      for (int ww = 1; ww < IF_WARPS; ww++) so += s_occ[ww];
      a.occ[blk] = so;
    }
    bar.sync();
    if (w == 0) {
      unsigned long long t = 0ull;
      for (int b = lane; b < G; b += 32) t += __ldcg(a.occ + b);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
      if (lane == 0) {
        float cell = 1.0f;
        if (n > 0 && t > 0ull) {
          const float occ = (float)((double)t / (double)n);
          float f = sqrtf(a.target_occ / occ);
          f = fminf(fmaxf(f, 0.125f), 2.0f);
          cell = 1.0f * f;
        }
        GridShape gs;
        grid_shape_compute(s_lo, s_hi, s_nfinite, cell, a.table_cap, gs);
        s_gs = gs;
      }
    }
    __syncthreads();
  }
  if (blk == 0 && tid == 0) {
    GridDesc* d = a.d;
    for (int c = 0; c < 3; c++) {
      d->bb_min[c] = f2ord(s_lo[c]); d->bb_max[c] = f2ord(s_hi[c]);
      d->origin[c] = s_gs.origin[c]; d->dim[c] = s_gs.dim[c];
    }
    d->cell = s_gs.cell; d->inv_cell = s_gs.inv_cell; d->ncells = s_gs.ncells; d->n = n;
    d->max_dim = s_gs.max_dim; d->margin = s_gs.margin; d->occ_sq = 0ull;
    d->nfinite = s_nfinite; d->vcount = 0; d->voverflow = 0;
  }
  const int ncells = s_gs.ncells;
  const bool write_hint = blk == 0 && tid == 0 && a.hint != nullptr && !(a.cell_req > 0.f);
  const float hint_cnt = write_hint ? (s_used_hint ? __ldcg(a.hint + 2) + 1.f : 0.f) : 0.f;
  const float hint_cell = s_gs.cell;

  // ---- zero the table, histogram of the cell keys (counts of cell c at table[c + 1]) ----
  {
    const int total = ncells + 2;
    int4* t4 = reinterpret_cast<int4*>(a.table + 1);              // &table[1] is 16-byte aligned
    const int n4 = (total - 1) >> 2;
    for (int i = gtid; i < n4; i += gstride) t4[i] = make_int4(0, 0, 0, 0);
    for (int i = 1 + (n4 << 2) + gtid; i < total; i += gstride) a.table[i] = 0;
    if (gtid == 0) a.table[0] = 0;
    if (a.bricks != nullptr) {
      const long long nb = (long long)((s_gs.dim[0] + KNN_BRICK - 1) / KNN_BRICK) * ((s_gs.dim[1] + KNN_BRICK - 1) / KNN_BRICK) *
                           ((s_gs.dim[2] + KNN_BRICK - 1) / KNN_BRICK);
      if (nb <= a.brick_cap) {
        unsigned* f4 = reinterpret_cast<unsigned*>(a.bricks);
        const long long nw = (nb + 3) / 4;
        for (long long i = gtid; i < nw; i += gstride) f4[i] = 0u;
      }
    }
  }
  bar.sync();
  {
    const float ox = s_gs.origin[0], oy = s_gs.origin[1], oz = s_gs.origin[2], inv = s_gs.inv_cell;
    const int dx = s_gs.dim[0], dy = s_gs.dim[1], dz = s_gs.dim[2];
    for (int i = gtid; i < n; i += gstride) {
      const float4 p = a.pts[i];
      const int cx = cell_coord(p.x, ox, inv, dx), cy = cell_coord(p.y, oy, inv, dy), cz = cell_coord(p.z, oz, inv, dz);
      const unsigned key = (unsigned)((cz * dy + cy) * dx + cx);
      a.keys[i] = key;
      atomicAdd(&a.table[key + 1], 1);
    }
  }
  bar.sync();

  // ---- exclusive scan of table[1 .. ncells] (data = table + 1, length ncells): slice per block ----
  int* data = a.table + 1;
  const int slice = (((ncells + G - 1) / G) + 3) & ~3;             // multiple of four: aligned int4 access
  const int s_lo_i = min(blk * slice, ncells), s_hi_i = min(s_lo_i + slice, ncells);
  {
    int sum = 0;
    const int n4 = (s_hi_i - s_lo_i) >> 2;
    const int4* d4 = reinterpret_cast<const int4*>(data + s_lo_i);
    for (int i = tid; i < n4; i += IF_THREADS) { const int4 v = __ldcg(d4 + i); sum += (v.x + v.y) + (v.z + v.w); }
    for (int i = s_lo_i + (n4 << 2) + tid; i < s_hi_i; i += IF_THREADS) sum += __ldcg(data + i);
    int total;
    if_block_exclusive_scan(sum, s_scan, total);
    if (tid == 0) a.slice_sum[blk] = total;
  }
  bar.sync();
  {
    int before = 0;
    for (int b = tid; b < blk; b += IF_THREADS) before += __ldcg(a.slice_sum + b);
    int carry;
    if_block_exclusive_scan(before, s_scan, carry);
    // the slice, 4 consecutive entries per thread and step, running carry across the steps
    for (int base = s_lo_i; base < s_hi_i; base += IF_THREADS * 4) {
      const int i = base + tid * 4;
      int4 v = make_int4(0, 0, 0, 0);
      if (i + 4 <= s_hi_i) v = __ldcg(reinterpret_cast<const int4*>(data + i));
      else {
        if (i < s_hi_i) v.x = __ldcg(data + i);
        if (i + 1 < s_hi_i) v.y = __ldcg(data + i + 1);
        if (i + 2 < s_hi_i) v.z = __ldcg(data + i + 2);
      }
      const int sum = (v.x + v.y) + (v.z + v.w);
      int total;
      const int off = carry + if_block_exclusive_scan(sum, s_scan, total);
      const int4 o = make_int4(off, off + v.x, off + v.x + v.y, off + v.x + v.y + v.z);
      if (i + 4 <= s_hi_i) *reinterpret_cast<int4*>(data + i) = o;
      else {
        if (i < s_hi_i) data[i] = o.x;
        if (i + 1 < s_hi_i) data[i + 1] = o.y;
        if (i + 2 < s_hi_i) data[i + 2] = o.z;
      }
      carry += total;
    }
  }
  bar.sync();

  // ---- counting-sort scatter (table[c + 1]: start(c) -> start(c + 1)), then deterministic in-cell order + gather ----
  for (int i = gtid; i < n; i += gstride) {
    const int pos = atomicAdd(&a.table[__ldcg(a.keys + i) + 1], 1);
    a.slot_orig[pos] = (unsigned)i;
  }
  bar.sync();
  const int nbx = (s_gs.dim[0] + KNN_BRICK - 1) / KNN_BRICK, nby = (s_gs.dim[1] + KNN_BRICK - 1) / KNN_BRICK,
            nbz = (s_gs.dim[2] + KNN_BRICK - 1) / KNN_BRICK;
  const bool mark = a.bricks != nullptr && (long long)nbx * nby * nbz <= a.brick_cap;
  for (int p = gtid; p < n; p += gstride) {
    const unsigned o = __ldcg(a.slot_orig + p);
    const unsigned key = __ldcg(a.keys + o);
    if (mark) {       // the point's box: plain stores of the same value, no atomics needed
      const int cx = (int)(key % (unsigned)s_gs.dim[0]), cyz = (int)(key / (unsigned)s_gs.dim[0]);
      const int cy = cyz % s_gs.dim[1], cz = cyz / s_gs.dim[1];
      a.bricks[((long long)(cz / KNN_BRICK) * nby + cy / KNN_BRICK) * nbx + cx / KNN_BRICK] = 1;
    }
    const int ca = __ldcg(a.table + key), cb = __ldcg(a.table + key + 1);
    int dst = p;
    if (cb - ca <= ORDER_FIX_MAX) {
      int rank = 0;
      for (int j = ca; j < cb; ++j) rank += (__ldcg(a.slot_orig + j) < o) ? 1 : 0;
      dst = ca + rank;
    }
    const float4 v = a.pts[o];
    a.sorted[dst] = make_float4(v.x, v.y, v.z, __uint_as_float(o));
  }
  // (every block read the hint before the first table pass, i.e. at least one barrier ago: nobody reads it any more)
  if (write_hint) { a.hint[0] = hint_cell; a.hint[1] = (float)n; a.hint[2] = hint_cnt; }
}

__global__ void __launch_bounds__(IF_THREADS, 2) index_fused_kernel(IndexFusedArgs a) {
  GridBar bar;
  bar.bar = a.bar; bar.phase = 0;
  index_fused_body(a, (int)gridDim.x, (int)blockIdx.x, bar);
  // leave the barrier words zeroed for the next launch: the last block to depart does it (nobody polls any more)
  if (threadIdx.x == 0 && atomicAdd(a.bar + 2, 1u) == gridDim.x - 1u) { a.bar[0] = 0u; a.bar[2] = 0u; }
}
// one cluster = one cloud (grid = IF_CLUSTER blocks)
__global__ void __launch_bounds__(IF_THREADS, 2) index_cluster_kernel(IndexFusedArgs a) {
  ClusterBar bar;
  index_fused_body(a, IF_CLUSTER, (int)cooperative_groups::this_cluster().block_rank(), bar);
  cooperative_groups::this_cluster().sync();     // no block of the cluster exits while another may still need it resident
}

static float auto_target_occupancy() {
  static float v = -1.f;
  if (v < 0.f) {
    const char* e = getenv("NGICP_TARGET_OCC");
    v = e ? (float)atof(e) : 10.0f;
    if (!(v > 0.f)) v = 10.0f;
  }
  return v;
}

static int bits_for(int cap) {
  int b = 1;
  while (b < 31 && (1ll << b) < (long long)cap) b++;
  return b;
}

cudaError_t build_index(DevCloud& c, float cell_req, int table_cap, Scratch& sc, const StreamPtr& st, int device) {
  cudaError_t e;
  const int n = c.n;
  if (table_cap < 64) table_cap = 64;
  c.table_cap = table_cap;
  // table buffer: 3 ints of padding so that &table[1] (where the scan runs) is 16-byte aligned
  if ((e = c.cell_start.acquire(sizeof(int) * ((size_t)table_cap + 8), device, st)) != cudaSuccess) return e;
  if ((e = c.sorted.alloc(sizeof(float4) * (n ? n : 1), st)) != cudaSuccess) return e;
  int* table = c.cell_start.as<int>() + 3;
  GridDesc* d = c.desc.as<GridDesc>();
  if (cell_req > 0.f || n == 0) {
    grid_setup_kernel<<<1, 1, 0, st->s>>>(d, cell_req > 0.f ? cell_req : 1.0f, table_cap, n);
    note_launches(1);
  } else {
    // two-stage: histogram a coarse trial grid, derive the cell edge from the point-weighted occupancy
    const float c0 = 1.0f;
    const int trial_cap = table_cap < (1 << 22) ? table_cap : (1 << 22);
    grid_setup_kernel<<<1, 1, 0, st->s>>>(d, c0, trial_cap, n);
    zero_cells_kernel<<<148 * 4, 256, 0, st->s>>>(table, d);
    count_points_kernel<<<grid_for(n), 256, 0, st->s>>>(c.pts.as<float4>(), n, d, table);
    occupancy_kernel<<<grid_for(n), 256, 0, st->s>>>(c.pts.as<float4>(), n, d, table);
    grid_autocell_kernel<<<1, 1, 0, st->s>>>(d, c0, auto_target_occupancy(), table_cap, n);
    note_launches(5);
  }
  zero_cells_kernel<<<148 * 4, 256, 0, st->s>>>(table, d);
  note_launches(1);
  if ((e = sc.tile_sums.reserve(sizeof(int) * scan_scratch_ints(table_cap + 1), st)) != cudaSuccess) return e;
  if (n > 0) {
    const size_t nb = sizeof(unsigned) * (size_t)n;
    if ((e = sc.keys_a.reserve(nb, st)) != cudaSuccess) return e;
    if ((e = sc.vals_a.reserve(nb, st)) != cudaSuccess) return e;
    // counting sort: histogram at table[c+1] -> exclusive scan (table[c+1] = start(c)) -> atomic scatter (table becomes
    // the lower-bound table) -> deterministic in-cell order + gather
    bin_points_kernel<<<grid_for(n), 256, 0, st->s>>>(c.pts.as<float4>(), n, d, sc.keys_a.as<unsigned>(), table);
    exclusive_scan_inplace(table + 1, &d->ncells, 0, table_cap, sc.tile_sums.as<int>(), st->s);
    scatter_points_kernel<<<grid_for(n), 256, 0, st->s>>>(sc.keys_a.as<unsigned>(), n, table, sc.vals_a.as<unsigned>());
    order_gather_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st->s>>>(c.pts.as<float4>(), sc.keys_a.as<unsigned>(), sc.vals_a.as<unsigned>(), n, table, c.sorted.as<float4>());
    note_launches(3);
  }
  c.indexed = true;
  return cudaGetLastError();
}

// upload_cloud + build_index in one cooperative launch (see index_fused_kernel); *done = false when the path does not
// apply (empty cloud, no cooperative launch, NGICP_INDEX_FUSED=0) and the caller must take the two functions above
cudaError_t upload_and_index_fused(DevCloud& c, const void* pts, size_t n, size_t stride_bytes, float cell_req, int table_cap, Scratch& sc,
                                   const StreamPtr& st, int device, bool* done) {
  *done = false;
  static const bool fused_on = !(getenv("NGICP_INDEX_FUSED") && atoi(getenv("NGICP_INDEX_FUSED")) == 0);
  if (!fused_on || n == 0 || sc.index_path == 1) return cudaSuccess;
  static std::mutex mu;
  static int max_blocks[64] = {};
  const int di = device & 63;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!max_blocks[di]) {
      int sms = 0, coop = 0, per_sm = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, index_fused_kernel, IF_THREADS, 0);
      max_blocks[di] = (coop && sms > 0 && per_sm > 0) ? sms * (per_sm > 2 ? 2 : per_sm) : -1;
      cudaGetLastError();
    }
  }
  if (max_blocks[di] <= 0) return cudaSuccess;
  cudaError_t e;
  if (table_cap < 64) table_cap = 64;
  c.n = (int)n;
  c.indexed = false;
  c.table_cap = table_cap;
  if ((e = c.pts.alloc(sizeof(float4) * n, st)) != cudaSuccess) return e;
  if ((e = c.desc.alloc(sizeof(GridDesc), st)) != cudaSuccess) return e;
  if ((e = c.cell_start.acquire(sizeof(int) * ((size_t)table_cap + 8), device, st)) != cudaSuccess) return e;
  if ((e = c.sorted.alloc(sizeof(float4) * n, st)) != cudaSuccess) return e;
  const size_t raw_bytes = (n - 1) * stride_bytes + 12;
  const unsigned char* raw = nullptr;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, pts) == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) &&
      (reinterpret_cast<uintptr_t>(pts) & 15) == 0 && (stride_bytes & 15) == 0 && stride_bytes >= 16) {
    raw = static_cast<const unsigned char*>(pts);
  } else {
    cudaGetLastError();
    if ((e = sc.staging.reserve(raw_bytes + 16, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(sc.staging.p, pts, raw_bytes, cudaMemcpyDefault, st->s)) != cudaSuccess) return e;
    if ((e = host_source_consumed(pts, sc, st->s)) != cudaSuccess) return e;
    raw = sc.staging.as<unsigned char>();
  }
  // index_path 3: one thread-block cluster per cloud (ordinary launch: many handles / streams interleave; a cooperative
  // launch must wait until all its blocks fit at once and holds the other streams up meanwhile).  Eight SMs have to zero
  // and scan the whole cell table, so this is for small voxelised clouds (22k points, 1 m cells: 0.10 ms against 0.044 ms
  // for the cooperative launch on an otherwise idle GPU; a raw 53k-point scan with its finer grid takes 0.73 ms): clouds
  // above 32k points take the cooperative launch
  const bool cluster = sc.index_path == 3 && n <= 32768;
  int blocks = (int)((n + IF_THREADS - 1) / IF_THREADS);
  if (blocks < 4) blocks = 4;
  if (blocks > max_blocks[di]) blocks = max_blocks[di];
  if (cluster) blocks = IF_CLUSTER;
  const size_t nb = sizeof(unsigned) * n;
  if ((e = sc.keys_a.reserve(nb, st)) != cudaSuccess) return e;
  if ((e = sc.vals_a.reserve(nb, st)) != cudaSuccess) return e;
  if ((e = sc.tile_sums.reserve(sizeof(int) * (size_t)(blocks * 12 + 64), st)) != cudaSuccess) return e;
  if (!sc.index_bar.p) {
    if ((e = sc.index_bar.reserve(128, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(sc.index_bar.p, 0, 128, st->s)) != cudaSuccess) return e;      // barrier words + cell-edge hint
  }
  IndexFusedArgs a;
  a.raw = raw; a.stride = stride_bytes; a.n = (int)n;
  a.pts = c.pts.as<float4>(); a.sorted = c.sorted.as<float4>(); a.table = c.cell_start.as<int>() + 3; a.d = c.desc.as<GridDesc>();
  a.keys = sc.keys_a.as<unsigned>(); a.slot_orig = sc.vals_a.as<unsigned>();
  int* ws = sc.tile_sums.as<int>();
  a.part = reinterpret_cast<float*>(ws);                                    // [blocks][8]
  a.occ = reinterpret_cast<unsigned long long*>(ws + (size_t)blocks * 8);   // [blocks] (8-byte aligned: blocks * 8 ints)
  a.slice_sum = ws + (size_t)blocks * 10;                                   // [blocks]
  a.bar = sc.index_bar.as<unsigned>();
  static const bool hint_on = !(getenv("NGICP_CELL_HINT") && atoi(getenv("NGICP_CELL_HINT")) == 0);
  a.hint = hint_on ? reinterpret_cast<float*>(sc.index_bar.as<unsigned>() + 16) : nullptr;
  a.bricks = nullptr; a.brick_cap = 0;
  if (n >= (size_t)KNN_BRICK_MIN_POINTS) {
    a.brick_cap = brick_flag_bytes(table_cap);
    if ((e = c.bricks.alloc((size_t)a.brick_cap, st)) != cudaSuccess) return e;
    a.bricks = c.bricks.as<unsigned char>();
  }
  a.cell_req = cell_req; a.target_occ = auto_target_occupancy(); a.table_cap = table_cap;
  a.trial_cap = table_cap < (1 << 22) ? table_cap : (1 << 22);
  void* kargs[] = {(void*)&a};
  if (cluster) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(IF_CLUSTER);
    cfg.blockDim = dim3(IF_THREADS);
    cfg.stream = st->s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = IF_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if ((e = cudaLaunchKernelEx(&cfg, index_cluster_kernel, a)) != cudaSuccess) return e;
  } else {
    if ((e = cudaLaunchCooperativeKernel((const void*)index_fused_kernel, dim3(blocks), dim3(IF_THREADS), kargs, 0, st->s)) != cudaSuccess) return e;
  }
  note_launches(1);
  c.indexed = true;
  *done = true;
  return cudaGetLastError();
}

void index_prime_kernels() {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, index_fused_kernel);
  cudaFuncGetAttributes(&fa, index_cluster_kernel);
  cudaGetLastError();
}

}  // namespace ngicp
