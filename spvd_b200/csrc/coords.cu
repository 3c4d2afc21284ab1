// Coordinate work of the SPVD hot path: fp32 coordinate scaling (V1), sorted-unique voxel sets
// (torch.unique(dim=0), models/sparse_utils.py:66) and the stride-2 coarse coordinate set (K2).
//
// Everything is integer/byte work bounded by HBM/L2 bandwidth and launch latency; sizes stay on the
// device (n_dev) so no entry point synchronises.  The sort is a hand-written LSD radix sort on the
// packed 64-bit keys: one up-front pass builds all eight digit histograms, passes whose digit is
// uniform (typically 4 of 8: the high byte of each 16-bit field) are skipped on the device without
// a host round trip (kernels read SortState::skip / parity and exit).
#include "common.cuh"

namespace spvd {

typedef unsigned long long u64;

constexpr int kSortThreads = 512;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
constexpr int kSortWarps = kSortThreads / 32;

struct SortState {
  unsigned hist[8][256];
  int skip[8];
  int parity[9];
  int n;
  int cmax[3];
  int pad;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct SortWs {
  u64* buf[2];
  unsigned* tile_hist;
  int* tile_count;
  SortState* st;
  int n_tiles;
};

static size_t sort_ws_bytes(int n_cap) {
  int n_tiles = cdiv(n_cap > 0 ? n_cap : 1, kSortTile);
  size_t b = 0;
  b += align_up((size_t)n_cap * 8, 256) * 2;
  b += align_up((size_t)256 * n_tiles * 4, 256);
  b += align_up((size_t)n_tiles * 4, 256);
  b += align_up(sizeof(SortState), 256);
  return b + 256;
}

static SortWs carve_ws(void* ws, int n_cap) {
  SortWs w;
  w.n_tiles = cdiv(n_cap > 0 ? n_cap : 1, kSortTile);
  char* p = (char*)align_up((size_t)ws, 256);
  w.buf[0] = (u64*)p;
  p += align_up((size_t)n_cap * 8, 256);
  w.buf[1] = (u64*)p;
  p += align_up((size_t)n_cap * 8, 256);
  w.tile_hist = (unsigned*)p;
  p += align_up((size_t)256 * w.n_tiles * 4, 256);
  w.tile_count = (int*)p;
  p += align_up((size_t)w.n_tiles * 4, 256);
  w.st = (SortState*)p;
  return w;
}

// ---------------------------------------------------------------------------------------------
// V1: fp32 scaling with the rounding sequence pinned (no FMA contraction, no fast math)
// ---------------------------------------------------------------------------------------------
__global__ void k_point_voxel_coords(const float4* __restrict__ pc, int n, float init_res, float after_res,
                                     float inv_after, int scale_mode, float4* __restrict__ fcoords,
                                     int4* __restrict__ icoords) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pc[i];
  float4 f;
  f.x = p.x;
  if (scale_mode == SPVD_SCALE_MUL) {
    f.y = __fmul_rn(__fmul_rn(p.y, init_res), inv_after);
    f.z = __fmul_rn(__fmul_rn(p.z, init_res), inv_after);
    f.w = __fmul_rn(__fmul_rn(p.w, init_res), inv_after);
  } else {
    f.y = __fdiv_rn(__fmul_rn(p.y, init_res), after_res);
    f.z = __fdiv_rn(__fmul_rn(p.z, init_res), after_res);
    f.w = __fdiv_rn(__fmul_rn(p.w, init_res), after_res);
  }
  fcoords[i] = f;
  if (icoords) icoords[i] = make_int4((int)floorf(f.x), (int)floorf(f.y), (int)floorf(f.z), (int)floorf(f.w));
}

// ---------------------------------------------------------------------------------------------
// key packing
// ---------------------------------------------------------------------------------------------
__global__ void k_state_init(SortState* st, int n_cap, const int* n_dev) {
  int t = threadIdx.x;
  unsigned* h = &st->hist[0][0];
  for (int i = t; i < 8 * 256; i += blockDim.x) h[i] = 0;
  if (t == 0) {
    st->n = resolve_n(n_dev, n_cap);
    st->cmax[0] = st->cmax[1] = st->cmax[2] = INT_MIN;
  }
}

__global__ void k_pack_keys(const int4* __restrict__ coords, SortState* st, u64* __restrict__ keys, int* status) {
  int n = st->n;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = coords[i];
  if (coord_in_range(c.x, c.y, c.z, c.w)) {
    keys[i] = pack_key(c.x, c.y, c.z, c.w);
  } else {
    keys[i] = kKeySentinel;
    if (status) atomicOr(status, SPVD_STATUS_COORD_RANGE);
  }
}

// batch-wide per-axis maximum of the fine coordinates (torchsparse: coords.max(0), frame at
// experiments/ConditionalGeneration.ipynb:451)
__global__ void k_coord_max(const int4* __restrict__ coords, SortState* st) {
  int n = st->n;
  int mx = INT_MIN, my = INT_MIN, mz = INT_MIN;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    mx = max(mx, c.y);
    my = max(my, c.z);
    mz = max(mz, c.w);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    my = max(my, __shfl_xor_sync(0xffffffffu, my, o));
    mz = max(mz, __shfl_xor_sync(0xffffffffu, mz, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&st->cmax[0], mx);
    atomicMax(&st->cmax[1], my);
    atomicMax(&st->cmax[2], mz);
  }
}

// parent = floor(c/2); spconv mode drops parents above (max_in - 1) // 2
__global__ void k_parent_keys(const int4* __restrict__ coords, SortState* st, int mode, u64* __restrict__ keys,
                              int* status) {
  int n = st->n;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = coords[i];
  int px = c.y >> 1, py = c.z >> 1, pz = c.w >> 1;
  bool keep = true;
  if (mode == SPVD_DOWNSAMPLE_SPCONV)
    keep = px <= ((st->cmax[0] - 1) >> 1) && py <= ((st->cmax[1] - 1) >> 1) && pz <= ((st->cmax[2] - 1) >> 1);
  if (!coord_in_range(c.x, px, py, pz)) {
    keep = false;
    if (status) atomicOr(status, SPVD_STATUS_COORD_RANGE);
  }
  keys[i] = keep ? pack_key(c.x, px, py, pz) : kKeySentinel;
}

// ---------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, keys only
// ---------------------------------------------------------------------------------------------
__global__ void k_digit_hist(const u64* __restrict__ keys, SortState* st) {
  __shared__ unsigned h[8 * 256];
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  int n = st->n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    u64 k = keys[i];
#pragma unroll
    for (int p = 0; p < 8; ++p) atomicAdd(&h[p * 256 + (int)((k >> (8 * p)) & 255)], 1u);
  }
  __syncthreads();
  unsigned* g = &st->hist[0][0];
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x)
    if (h[i]) atomicAdd(g + i, h[i]);
}

__global__ void k_sort_plan(SortState* st, int first_pass = 0) {
  __shared__ int uniform[8];
  if (threadIdx.x < 8) uniform[threadIdx.x] = threadIdx.x < first_pass ? 1 : 0;   // digits the input is already sorted by
  __syncthreads();
  unsigned n = (unsigned)st->n;
  for (int p = 0; p < 8; ++p)
    if (n == 0 || st->hist[p][threadIdx.x] == n) uniform[p] = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    int par = 0;
    for (int p = 0; p < 8; ++p) {
      st->skip[p] = uniform[p];
      st->parity[p] = par;
      if (!uniform[p]) par ^= 1;
    }
    st->parity[8] = par;
  }
}

__global__ void __launch_bounds__(kSortThreads) k_tile_hist(const u64* __restrict__ b0, const u64* __restrict__ b1,
                                                            const SortState* __restrict__ st, int pass, int n_tiles,
                                                            unsigned* __restrict__ tile_hist) {
  if (st->skip[pass]) return;
  __shared__ unsigned h[256];
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  __syncthreads();
  const u64* src = st->parity[pass] ? b1 : b0;
  int n = st->n, shift = 8 * pass;
  int base = blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    int i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(int)((src[i] >> shift) & 255)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256) tile_hist[threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) k_scatter(u64* __restrict__ b0, u64* __restrict__ b1,
                                                          const SortState* __restrict__ st, int pass, int n_tiles,
                                                          const unsigned* __restrict__ tile_hist) {
  if (st->skip[pass]) return;
  __shared__ unsigned base[256];
  __shared__ unsigned scan[256];
  __shared__ unsigned wcount[kSortWarps][256];
  const u64* src = st->parity[pass] ? b1 : b0;
  u64* dst = st->parity[pass] ? b0 : b1;
  const int n = st->n, shift = 8 * pass, tile = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid < 256) {
    unsigned total = 0, before = 0;
    const unsigned* row = tile_hist + tid * n_tiles;
    for (int t = 0; t < n_tiles; ++t) {
      unsigned c = row[t];
      total += c;
      if (t < tile) before += c;
    }
    scan[tid] = total;
    base[tid] = before;
  }
  __syncthreads();
  if (tid < 32) {  // exclusive scan of the 256 digit totals by one warp (8 per lane)
    unsigned v[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = scan[tid * 8 + j];
      s += v[j];
    }
    unsigned incl = s;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    unsigned run = incl - s;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      scan[tid * 8 + j] = run;
      run += v[j];
    }
  }
  __syncthreads();
  if (tid < 256) base[tid] += scan[tid];
  for (int r = 0; r < kSortItems; ++r) {
    for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();
    int i = tile * kSortTile + r * kSortThreads + tid;
    bool valid = i < n;
    u64 key = valid ? src[i] : 0ull;
    int digit = valid ? (int)((key >> shift) & 255) : 256;
    unsigned peers = __match_any_sync(0xffffffffu, digit);
    int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) wcount[warp][digit] = __popc(peers);
    __syncthreads();
    if (tid < 256) {
      unsigned run = base[tid];
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        unsigned c = wcount[w][tid];
        wcount[w][tid] = run;
        run += c;
      }
      base[tid] = run;
    }
    __syncthreads();
    if (valid) dst[wcount[warp][digit] + rank] = key;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// unique over the sorted keys (the sentinel, which sorts last, is not a voxel)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int head_flags(const u64* __restrict__ src, int n, int first, u64* k) {
  int cnt = 0;
  u64 prev = (first > 0 && first <= n) ? src[first - 1] : kKeySentinel;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    int i = first + j;
    k[j] = i < n ? src[i] : kKeySentinel;
    bool head = k[j] != kKeySentinel && (i == 0 || k[j] != prev);
    prev = k[j];
    cnt += head ? 1 : 0;
  }
  return cnt;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int wsum[kSortWarps];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int wbase = 0, tot = 0;
  for (int w = 0; w < kSortWarps; ++w) {
    int s = wsum[w];
    if (w < warp) wbase += s;
    tot += s;
  }
  *total = tot;
  return wbase + incl - v;
}

__global__ void __launch_bounds__(kSortThreads) k_unique_count(const u64* __restrict__ b0, const u64* __restrict__ b1,
                                                               const SortState* __restrict__ st,
                                                               int* __restrict__ tile_count) {
  const u64* src = st->parity[8] ? b1 : b0;
  u64 k[kSortItems];
  int cnt = head_flags(src, st->n, blockIdx.x * kSortTile + threadIdx.x * kSortItems, k);
  int total;
  block_exclusive_scan(cnt, &total);
  if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kSortThreads) k_unique_write(const u64* __restrict__ b0, const u64* __restrict__ b1,
                                                               const SortState* __restrict__ st,
                                                               const int* __restrict__ tile_count, int n_tiles,
                                                               int4* __restrict__ out_coords,
                                                               int* __restrict__ out_count) {
  const u64* src = st->parity[8] ? b1 : b0;
  __shared__ int s_base;
  if (threadIdx.x < 32) {
    int s = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += 32) s += tile_count[t];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) s_base = s;
  }
  const int n = st->n;
  const int first = blockIdx.x * kSortTile + threadIdx.x * kSortItems;
  u64 k[kSortItems];
  int cnt = head_flags(src, n, first, k);
  int total;
  int pos = block_exclusive_scan(cnt, &total);  // contains a __syncthreads: s_base is visible below
  pos += s_base;
  u64 prev = (first > 0 && first <= n) ? src[first - 1] : kKeySentinel;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    bool head = k[j] != kKeySentinel && (first + j == 0 || k[j] != prev);
    prev = k[j];
    if (head) out_coords[pos++] = unpack_key(k[j]);
  }
  if ((int)blockIdx.x == n_tiles - 1 && threadIdx.x == 0) *out_count = s_base + total;
}

static int sorted_unique(const SortWs& w, int n_cap, int4* out_coords, int* out_count, cudaStream_t s) {
  int grid = cdiv(n_cap, 1024);
  if (grid > 296) grid = 296;
  if (grid < 1) grid = 1;
  k_digit_hist<<<grid, 256, 0, s>>>(w.buf[0], w.st);
  k_sort_plan<<<1, 256, 0, s>>>(w.st, 0);
  for (int p = 0; p < 8; ++p) {
    k_tile_hist<<<w.n_tiles, kSortThreads, 0, s>>>(w.buf[0], w.buf[1], w.st, p, w.n_tiles, w.tile_hist);
    k_scatter<<<w.n_tiles, kSortThreads, 0, s>>>(w.buf[0], w.buf[1], w.st, p, w.n_tiles, w.tile_hist);
  }
  k_unique_count<<<w.n_tiles, kSortThreads, 0, s>>>(w.buf[0], w.buf[1], w.st, w.tile_count);
  k_unique_write<<<w.n_tiles, kSortThreads, 0, s>>>(w.buf[0], w.buf[1], w.st, w.tile_count, w.n_tiles, out_coords,
                                                    out_count);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

// ---------------------------------------------------------------------------------------------
// Mask-sorted view of a kernel map (what torchsparse calls the "sorted implicit GEMM" order): output rows sorted by the
// bit set of their active kernel offsets, so that the 128 rows of a tile share (most of) their offsets and an
// output-stationary convolution tile issues few zero rows.  Key = (mask << 32) | row: the input is already sorted by
// the low 32 bits and every pass of the LSD sort is stable, so only the four mask digits are sorted.
// ---------------------------------------------------------------------------------------------
__global__ void k_mask_keys(const int* __restrict__ map_t, int ld, int kvol, const SortState* __restrict__ st,
                            u64* __restrict__ keys) {
  const int n = st->n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned m = 0;
  for (int k = 0; k < kvol; ++k) m |= (__ldg(map_t + (size_t)k * ld + i) >= 0 ? 1u : 0u) << k;
  keys[i] = ((u64)m << 32) | (u64)(unsigned)i;
}

__global__ void __launch_bounds__(128) k_sorted_map(const u64* __restrict__ b0, const u64* __restrict__ b1,
                                                    const SortState* __restrict__ st, const int* __restrict__ map_t, int ld,
                                                    int kvol, int* __restrict__ map_sorted, unsigned* __restrict__ mask_sorted,
                                                    int* __restrict__ perm) {
  __shared__ unsigned wm[4];
  const u64* src = st->parity[8] ? b1 : b0;
  const int n = st->n, j = blockIdx.x * 128 + threadIdx.x;
  int row = -1;
  unsigned m = 0;
  if (j < n) {
    const u64 key = src[j];
    row = (int)(unsigned)(key & 0xffffffffull);
    m = (unsigned)(key >> 32);
  }
  perm[j] = row;
  for (int k = 0; k < kvol; ++k) map_sorted[(size_t)k * ld + j] = row >= 0 ? __ldg(map_t + (size_t)k * ld + row) : -1;
  m = __reduce_or_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) mask_sorted[blockIdx.x] = wm[0] | wm[1] | wm[2] | wm[3];
}

// ---------------------------------------------------------------------------------------------
// Segmented path: when the rows are batch-major and no cloud has more than `cap` rows (the caller's
// hint), each cloud is sorted and de-duplicated by ONE CTA entirely in shared memory (bitonic sort of
// the packed keys) -- 3 launches per level instead of ~20 for the global radix sort.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_batch(const int4* __restrict__ coords, int n, int b) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (coords[mid].x < b) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// The same answer from a whole warp: 32 probes per round (a 33-ary search), ~5 dependent loads instead of ~17 for 65k
// rows -- the planner kernels below are latency chains on the step's critical path, and two scalar searches per thread
// were most of k_seg_sort's 27 us.  All 32 lanes must call; all return the result.
__device__ __forceinline__ int lower_bound_batch_warp(const int4* __restrict__ coords, int n, int b) {
  const int lane = threadIdx.x & 31;
  int lo = 0, hi = n;
  while (hi > lo) {
    const int step = (hi - lo + 32) / 33;                  // >= 1
    const int p = lo + step * (lane + 1) - 1;              // probe positions, increasing with the lane
    const bool less = p < hi && coords[p].x < b;           // a prefix of the lanes (rows are sorted by batch)
    const int c = __popc(__ballot_sync(0xffffffffu, less));
    const int first_not_less = lo + step * (c + 1) - 1;    // position of probe c (answer <= it when it is inside the range)
    lo += step * c;                                        // answer > position of probe c - 1
    if (c < 32 && first_not_less < hi) hi = first_not_less;
  }
  return lo;
}

__global__ void k_coord_max2(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev,
                             int* __restrict__ cmax) {
  int n = resolve_n(n_dev, n_cap);
  int mx = INT_MIN, my = INT_MIN, mz = INT_MIN;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    mx = max(mx, c.y);
    my = max(my, c.z);
    mz = max(mz, c.w);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    my = max(my, __shfl_xor_sync(0xffffffffu, my, o));
    mz = max(mz, __shfl_xor_sync(0xffffffffu, mz, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&cmax[0], mx);
    atomicMax(&cmax[1], my);
    atomicMax(&cmax[2], mz);
  }
}

// parent_mode: 0 = keys of the rows themselves (stride-1 voxels of the points), 1 = floor(c/2), 2 = floor(c/2)
// with the spconv bound (cmax = batch-wide maxima of the fine level)
__global__ void __launch_bounds__(1024) k_seg_sort(const int4* __restrict__ coords, int n_cap,
                                                   const int* __restrict__ n_dev, int cap, int parent_mode,
                                                   const int* __restrict__ cmax, u64* __restrict__ temp,
                                                   int* __restrict__ cloud_count, int* status) {
  extern __shared__ u64 s_keys[];
  __shared__ int s_wsum[32];
  __shared__ int s_range[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = resolve_n(n_dev, n_cap);
  if (tid < 64) {                                          // warp 0: first row of cloud b, warp 1: of cloud b + 1
    const int r = lower_bound_batch_warp(coords, n, b + (tid >> 5));
    if ((tid & 31) == 0) s_range[tid >> 5] = r;
  }
  __syncthreads();
  const int lo = s_range[0], hi = s_range[1];
  int cnt = hi - lo;
  if (cnt > cap) {
    if (tid == 0 && status) atomicOr(status, SPVD_STATUS_SEGMENT_OVERFLOW);
    cnt = cap;
  }
  int bx = 0, by = 0, bz = 0;
  if (parent_mode == 2) {
    bx = (cmax[0] - 1) >> 1;
    by = (cmax[1] - 1) >> 1;
    bz = (cmax[2] - 1) >> 1;
  }
  for (int i = tid; i < cap; i += 1024) {
    u64 key = kKeySentinel;
    if (i < cnt) {
      int4 c = coords[lo + i];
      int x = c.y, y = c.z, z = c.w;
      bool keep = true;
      if (parent_mode) {
        x >>= 1; y >>= 1; z >>= 1;
        if (parent_mode == 2) keep = x <= bx && y <= by && z <= bz;
      }
      if (!coord_in_range(c.x, x, y, z)) {
        keep = false;
        if (status) atomicOr(status, SPVD_STATUS_COORD_RANGE);
      }
      if (keep) key = pack_key(c.x, x, y, z);
    }
    s_keys[i] = key;
  }
  __syncthreads();
  // bitonic sort of the first `n_sort` = next_pow2(cnt) keys (the rest is sentinel padding that already sits at the end).
  // Steps with distance j <= 32 keep every warp inside its own 64-key block: they need __syncwarp only.
  int n_sort = 64;
  while (n_sort < cnt) n_sort <<= 1;
  for (int k = 2; k <= n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n_sort >> 1); t += 1024) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // lower index of the pair
        const int l = i | j;
        const bool up = (i & k) == 0;
        const u64 a = s_keys[i], c2 = s_keys[l];
        if ((a > c2) == up) {
          s_keys[i] = c2;
          s_keys[l] = a;
        }
      }
      if (j > 32) __syncthreads();
      else __syncwarp();
    }
    if (k >= 64) __syncthreads();   // the next level starts with distance k > 32
  }
  __syncthreads();
  // unique: each thread owns a contiguous run of cap/1024 (>= 1) elements
  const int per = (cap + 1023) >> 10;
  const int first = tid * per;
  int heads = 0;
  for (int j = 0; j < per; ++j) {
    const int i = first + j;
    if (i < cap) {
      const u64 k = s_keys[i];
      heads += (k != kKeySentinel && (i == 0 || k != s_keys[i - 1])) ? 1 : 0;
    }
  }
  int incl = heads;
  const int lane = tid & 31, warp = tid >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  int wbase = 0, total = 0;
  for (int w = 0; w < 32; ++w) {
    const int v = s_wsum[w];
    if (w < warp) wbase += v;
    total += v;
  }
  int pos = lo + wbase + incl - heads;
  for (int j = 0; j < per; ++j) {
    const int i = first + j;
    if (i < cap) {
      const u64 k = s_keys[i];
      if (k != kKeySentinel && (i == 0 || k != s_keys[i - 1])) temp[pos++] = k;
    }
  }
  if (tid == 0) cloud_count[b] = total;
}

// gather the per-cloud unique runs into the level's coordinate array; also yields the level's per-cloud
// counts (= cloud_count) and offsets
__global__ void __launch_bounds__(256) k_seg_compact(const int4* __restrict__ src_coords, int n_cap,
                                                     const int* __restrict__ n_dev, const u64* __restrict__ temp,
                                                     const int* __restrict__ cloud_count, int nb,
                                                     int4* __restrict__ out_coords, int* __restrict__ out_count,
                                                     int* __restrict__ batch_offsets) {
  __shared__ int s_base, s_lo;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = resolve_n(n_dev, n_cap);
  if (tid < 32) {
    int sum = 0;
    for (int i = tid; i < b; i += 32) sum += cloud_count[i];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (tid == 0) s_base = sum;
  } else if (tid < 64) {
    const int r = lower_bound_batch_warp(src_coords, n, b);
    if (tid == 32) s_lo = r;
  }
  __syncthreads();
  const int base = s_base, cnt = cloud_count[b];
  const int lo = s_lo;
  for (int i = tid; i < cnt; i += 256) out_coords[base + i] = unpack_key(temp[lo + i]);
  if (tid == 0) {
    if (batch_offsets) batch_offsets[b] = base;
    if (b == nb - 1) {
      *out_count = base + cnt;
      if (batch_offsets) batch_offsets[nb] = base + cnt;
    }
  }
}

static int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// sorted-unique of batch-major rows, one CTA per cloud.  temp: n_cap u64; cloud_count: nb int32 (doubles as
// the new level's per-cloud counts); cmax: 3 int32 scratch (spconv bound)
static int seg_sorted_unique(const int4* coords, int n_cap, const int* n_dev, int nb, int max_cloud_rows, int parent_mode,
                             int* cmax, u64* temp, int* cloud_count, int4* out_coords, int* out_count,
                             int* batch_offsets, int* status, cudaStream_t s) {
  const int cap = next_pow2(max_cloud_rows);
  const size_t smem = (size_t)cap * 8;
  static bool configured_dev[64] = {};   // the attribute is per device: one flag per device, not per process
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(k_seg_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    configured = true;
  }
  if (parent_mode == 2) {
    SPVD_CUDA(cudaMemsetAsync(cmax, 0x80, 12, s));   // 0x80808080 = very negative
    int grid = cdiv(n_cap, 1024);
    if (grid > 148) grid = 148;
    k_coord_max2<<<grid, 256, 0, s>>>(coords, n_cap, n_dev, cmax);
  }
  k_seg_sort<<<nb, 1024, smem, s>>>(coords, n_cap, n_dev, cap, parent_mode, cmax, temp, cloud_count, status);
  k_seg_compact<<<nb, 256, 0, s>>>(coords, n_cap, n_dev, temp, cloud_count, nb, out_coords, out_count, batch_offsets);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

}  // namespace spvd

using namespace spvd;

// Segmented variants of spvd_unique_coords / spvd_downsample_coords for batch-major rows with at most
// max_cloud_rows (<= 16384) rows per cloud.  parent = 0: unique of the rows themselves; 1: stride-2 parents.
// Also writes the new level's per-cloud counts [n_batch] and offsets [n_batch+1].  workspace as for
// spvd_unique_workspace_bytes(n_cap).
extern "C" int spvd_segmented_unique(const int32_t* coords, int n_cap, const int32_t* n_dev, int n_batch,
                                     int max_cloud_rows, int parent, int downsample_mode, int32_t* out_coords,
                                     int32_t* out_count, int32_t* batch_counts, int32_t* batch_offsets, int32_t* status,
                                     void* workspace, size_t workspace_bytes, spvd_stream_t stream) {
  SPVD_CHECK_ARG(coords && out_coords && out_count && batch_counts && workspace && n_cap > 0 && n_batch > 0,
                 "spvd_segmented_unique: bad arguments");
  SPVD_CHECK_ARG(max_cloud_rows > 0 && max_cloud_rows <= 16384, "spvd_segmented_unique: max_cloud_rows %d not in 1..16384",
                 max_cloud_rows);
  if (workspace_bytes < sort_ws_bytes(n_cap)) {
    set_error("spvd_segmented_unique: workspace %zu < %zu", workspace_bytes, sort_ws_bytes(n_cap));
    return SPVD_ERR_WORKSPACE;
  }
  SortWs w = carve_ws(workspace, n_cap);
  int mode = !parent ? 0 : (downsample_mode == SPVD_DOWNSAMPLE_SPCONV ? 2 : 1);
  return seg_sorted_unique((const int4*)coords, n_cap, n_dev, n_batch, max_cloud_rows, mode, &w.st->cmax[0], w.buf[0],
                           batch_counts, (int4*)out_coords, out_count, batch_offsets, status, (cudaStream_t)stream);
}

extern "C" int spvd_point_voxel_coords(const float* pc, int n, float init_res, float after_res, int scale_mode,
                                       float* fcoords, int32_t* icoords, spvd_stream_t stream) {
  SPVD_CHECK_ARG(n >= 0 && pc && fcoords, "spvd_point_voxel_coords: bad arguments");
  SPVD_CHECK_ARG(scale_mode == SPVD_SCALE_MUL || scale_mode == SPVD_SCALE_DIV, "bad scale_mode %d", scale_mode);
  if (n == 0) return SPVD_OK;
  volatile float one = 1.0f;
  float inv = one / after_res;  // fp32 reciprocal, as torch computes it for a CUDA division by a scalar
  k_point_voxel_coords<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)pc, n, init_res, after_res, inv, scale_mode, (float4*)fcoords, (int4*)icoords);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" size_t spvd_unique_workspace_bytes(int n_cap) { return sort_ws_bytes(n_cap); }

extern "C" int spvd_unique_coords(const int32_t* icoords, int n_cap, const int32_t* n_dev, int32_t* out_coords,
                                  int32_t* out_count, int32_t* status, void* workspace, size_t workspace_bytes,
                                  spvd_stream_t stream) {
  SPVD_CHECK_ARG(n_cap > 0 && icoords && out_coords && out_count && workspace, "spvd_unique_coords: bad arguments");
  if (workspace_bytes < sort_ws_bytes(n_cap)) {
    set_error("spvd_unique_coords: workspace %zu < %zu", workspace_bytes, sort_ws_bytes(n_cap));
    return SPVD_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  SortWs w = carve_ws(workspace, n_cap);
  k_state_init<<<1, 256, 0, s>>>(w.st, n_cap, n_dev);
  k_pack_keys<<<cdiv(n_cap, 256), 256, 0, s>>>((const int4*)icoords, w.st, w.buf[0], status);
  return sorted_unique(w, n_cap, (int4*)out_coords, out_count, s);
}

extern "C" size_t spvd_downsample_workspace_bytes(int n_cap) { return sort_ws_bytes(n_cap); }

extern "C" int spvd_downsample_coords(const int32_t* coords, int n_cap, const int32_t* n_dev, int mode,
                                      int32_t* out_coords, int32_t* out_count, int32_t* status, void* workspace,
                                      size_t workspace_bytes, spvd_stream_t stream) {
  SPVD_CHECK_ARG(n_cap > 0 && coords && out_coords && out_count && workspace, "spvd_downsample_coords: bad arguments");
  SPVD_CHECK_ARG(mode == SPVD_DOWNSAMPLE_SPCONV || mode == SPVD_DOWNSAMPLE_FLOOR, "bad downsample mode %d", mode);
  if (workspace_bytes < sort_ws_bytes(n_cap)) {
    set_error("spvd_downsample_coords: workspace %zu < %zu", workspace_bytes, sort_ws_bytes(n_cap));
    return SPVD_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  SortWs w = carve_ws(workspace, n_cap);
  k_state_init<<<1, 256, 0, s>>>(w.st, n_cap, n_dev);
  if (mode == SPVD_DOWNSAMPLE_SPCONV) {
    int grid = cdiv(n_cap, 1024);
    if (grid > 148) grid = 148;
    k_coord_max<<<grid, 256, 0, s>>>((const int4*)coords, w.st);
  }
  k_parent_keys<<<cdiv(n_cap, 256), 256, 0, s>>>((const int4*)coords, w.st, mode, w.buf[0], status);
  return sorted_unique(w, n_cap, (int4*)out_coords, out_count, s);
}

extern "C" size_t spvd_kmap_sort_workspace_bytes(int n_cap) { return sort_ws_bytes(n_cap); }

extern "C" int spvd_kmap_sort(const int32_t* map_t, int ld, int n_cap, const int32_t* n_dev, int kvol, int32_t* map_sorted,
                              uint32_t* mask_sorted, int32_t* perm, void* workspace, size_t workspace_bytes,
                              spvd_stream_t stream) {
  SPVD_CHECK_ARG(map_t && map_sorted && mask_sorted && perm && workspace && n_cap > 0 && kvol >= 1 && kvol <= 32,
                 "spvd_kmap_sort: bad arguments");
  SPVD_CHECK_ARG(ld % 128 == 0 && ld >= n_cap, "spvd_kmap_sort: ld %d must be a multiple of 128 >= n_cap", ld);
  SPVD_CHECK_ARG(workspace_bytes >= sort_ws_bytes(n_cap), "spvd_kmap_sort: workspace too small (%zu < %zu)", workspace_bytes,
                 sort_ws_bytes(n_cap));
  cudaStream_t s = (cudaStream_t)stream;
  SortWs w = carve_ws(workspace, n_cap);
  k_state_init<<<1, 256, 0, s>>>(w.st, n_cap, n_dev);
  k_mask_keys<<<cdiv(n_cap, 256), 256, 0, s>>>(map_t, ld, kvol, w.st, w.buf[0]);
  int grid = cdiv(n_cap, 1024);
  if (grid > 296) grid = 296;
  k_digit_hist<<<grid, 256, 0, s>>>(w.buf[0], w.st);
  k_sort_plan<<<1, 256, 0, s>>>(w.st, 4);
  for (int p = 4; p < 4 + (kvol + 7) / 8; ++p) {
    k_tile_hist<<<w.n_tiles, kSortThreads, 0, s>>>(w.buf[0], w.buf[1], w.st, p, w.n_tiles, w.tile_hist);
    k_scatter<<<w.n_tiles, kSortThreads, 0, s>>>(w.buf[0], w.buf[1], w.st, p, w.n_tiles, w.tile_hist);
  }
  k_sorted_map<<<ld / 128, 128, 0, s>>>(w.buf[0], w.buf[1], w.st, map_t, ld, kvol, map_sorted, mask_sorted, perm);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
