// Coordinate hash table (H1) and everything that probes it: generic neighbourhood queries
// (torchsparse.backend.GPUHashTable as used by models/sparse_utils.py:36-55), the submanifold
// kernel map (K1), the stride-2 maps (K2/K3) and the point->voxel / trilinear queries (V2/V3).
//
// Open addressing, linear probing, 64-bit packed keys, load factor <= 0.5 (the reference sizes its
// tables 2*N).  The tables of one forward are a few MB and live in L2; each probe is one 8-byte
// load in the common case.  All kernels are one thread per query row with offset-major, coalesced
// stores.
#include "common.cuh"

namespace spvd {

typedef unsigned long long u64;

__device__ __forceinline__ int4 load_coord(const int* __restrict__ c, int i, int order) {
  int4 v = reinterpret_cast<const int4*>(c)[i];
  if (order == SPVD_ORDER_XYZB) return make_int4(v.w, v.x, v.y, v.z);
  return v;
}

__global__ void k_hash_insert(const int* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, int order,
                              u64* keys, int* vals, unsigned cap, int* status) {
  int n = resolve_n(n_dev, n_cap);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = load_coord(coords, i, order);
  if (!coord_in_range(c.x, c.y, c.z, c.w)) {
    if (status) atomicOr(status, SPVD_STATUS_COORD_RANGE);
    return;
  }
  u64 key = pack_key(c.x, c.y, c.z, c.w);
  unsigned slot = hash_slot(key, cap);
  for (unsigned probe = 0; probe < cap; ++probe) {
    u64 old = atomicCAS(keys + slot, 0ull, key);
    if (old == 0ull || old == key) {
      vals[slot] = i + 1;
      return;
    }
    if (++slot == cap) slot = 0;
  }
  if (status) atomicOr(status, SPVD_STATUS_HASH_FULL);
}

__device__ __forceinline__ int find_coord(const u64* __restrict__ keys, const int* __restrict__ vals, unsigned cap,
                                          int b, int x, int y, int z) {
  if (!coord_in_range(b, x, y, z)) return 0;
  return hash_find(keys, vals, cap, pack_key(b, x, y, z));
}

// offsets: kernel_size 3 -> {-1,0,1}^3 with x fastest; 2 -> {0,1}^3 with z fastest; 1 -> {0}
__device__ __forceinline__ void kernel_offset(int ks, int k, int& dx, int& dy, int& dz) {
  if (ks == 3) {
    dx = k % 3 - 1;
    dy = (k / 3) % 3 - 1;
    dz = k / 9 - 1;
  } else if (ks == 2) {
    dx = (k >> 2) & 1;
    dy = (k >> 1) & 1;
    dz = k & 1;
  } else {
    dx = dy = dz = 0;
  }
}

__global__ void k_hash_query(const int* __restrict__ query, int m, int order, int ks, int kvol,
                             const u64* __restrict__ keys, const int* __restrict__ vals, unsigned cap,
                             int* __restrict__ out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m * kvol) return;
  int i = (int)(t / kvol), k = (int)(t % kvol);
  int4 c = load_coord(query, i, order);
  int dx, dy, dz;
  kernel_offset(ks, k, dx, dy, dz);
  out[t] = find_coord(keys, vals, cap, c.x, c.y + dx, c.z + dy, c.w + dz);
}

// K1: one CTA = one 128-row output tile; three threads per voxel, each with the 9 probes of one dz plane in flight (the
// kernel is a latency chain of L2-resident probes: three times the threads, a third of the chain per thread)
__global__ void __launch_bounds__(384) k_kmap_subm3(const int4* __restrict__ coords, int n_cap,
                                                    const int* __restrict__ n_dev, const u64* __restrict__ keys,
                                                    const int* __restrict__ vals, unsigned cap,
                                                    int* __restrict__ map_t, int ld, unsigned* __restrict__ tile_mask) {
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0;
  __syncthreads();
  int n = resolve_n(n_dev, n_cap);
  const int row = threadIdx.x & 127, k0 = 9 * (threadIdx.x >> 7);
  int o = blockIdx.x * 128 + row;
  bool valid = o < n;
  int4 c = valid ? coords[o] : make_int4(0, 0, 0, 0);
  unsigned wmask = 0;
  int r[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int k = k0 + j;
    r[j] = -1;
    if (valid) {
      if (k == 13) {
        r[j] = o;
      } else {
        int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
        r[j] = find_coord(keys, vals, cap, c.x, c.y + dx, c.z + dy, c.w + dz) - 1;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    if (o < ld) map_t[(size_t)(k0 + j) * ld + o] = r[j];
    if (__ballot_sync(0xffffffffu, r[j] >= 0)) wmask |= 1u << (k0 + j);
  }
  if ((threadIdx.x & 31) == 0 && wmask) atomicOr(&s_mask, wmask);
  __syncthreads();
  if (threadIdx.x == 0 && tile_mask) tile_mask[blockIdx.x] = s_mask;
}

// K2/K3: fine voxel i has parent floor(c/2) (a row of the coarse table or absent) and parity offset
// koff = 4*px + 2*py + pz.
__global__ void k_kmap_down2(const int4* __restrict__ fine, int n_cap, const int* __restrict__ n_dev,
                             const u64* __restrict__ keys, const int* __restrict__ vals, unsigned cap,
                             int* __restrict__ map_down_t, int ld_coarse, unsigned* tile_mask_down,
                             int* __restrict__ map_up_t, int ld_fine, unsigned* tile_mask_up,
                             int* __restrict__ parent_out, int* __restrict__ koff_out) {
  int n = resolve_n(n_dev, n_cap);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld_fine && i >= n_cap) return;
  int prow = -1, koff = 0;
  if (i < n) {
    int4 c = fine[i];
    int px = c.y >> 1, py = c.z >> 1, pz = c.w >> 1;
    koff = ((c.y & 1) << 2) | ((c.z & 1) << 1) | (c.w & 1);
    prow = find_coord(keys, vals, cap, c.x, px, py, pz) - 1;
    if (prow >= 0) {
      if (map_down_t && prow < ld_coarse) {
        map_down_t[(size_t)koff * ld_coarse + prow] = i;
        if (tile_mask_down) atomicOr(tile_mask_down + (prow >> 7), 1u << koff);
      }
      if (tile_mask_up) atomicOr(tile_mask_up + (i >> 7), 1u << koff);
    }
  }
  if (i < n_cap) {
    if (parent_out) parent_out[i] = prow;
    if (koff_out) koff_out[i] = koff;
  }
  if (map_up_t && i < ld_fine) {
#pragma unroll
    for (int k = 0; k < 8; ++k) map_up_t[(size_t)k * ld_fine + i] = (k == koff && prow >= 0) ? prow : -1;
  }
}

// V2: voxel of a point at tensor stride s = (int(b), floor(xyz / s))  (models/sparse_utils.py:81-87)
__global__ void k_point_voxel_query(const float4* __restrict__ pc, int n, float s, const u64* __restrict__ keys,
                                    const int* __restrict__ vals, unsigned cap, int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pc[i];
  int x = (int)floorf(__fdiv_rn(p.y, s)), y = (int)floorf(__fdiv_rn(p.z, s)), z = (int)floorf(__fdiv_rn(p.w, s));
  idx[i] = find_coord(keys, vals, cap, (int)p.x, x, y, z) - 1;
}

// V3: 8 corner rows (z fastest) + trilinear weights (torchsparse calc_ti_weights, scale=1): zero where
// the corner is absent, renormalised by (sum + 1e-8)
__global__ void k_devox_query(const float4* __restrict__ pc, int n, float s, const u64* __restrict__ keys,
                              const int* __restrict__ vals, unsigned cap, int* __restrict__ idx8,
                              float* __restrict__ w8) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pc[i];
  float x = __fdiv_rn(p.y, s), y = __fdiv_rn(p.z, s), z = __fdiv_rn(p.w, s);
  float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  int b = (int)p.x, ix = (int)xf, iy = (int)yf, iz = (int)zf;
  float ax[2] = {__fsub_rn(__fadd_rn(xf, 1.f), x), __fsub_rn(x, xf)};
  float ay[2] = {__fsub_rn(__fadd_rn(yf, 1.f), y), __fsub_rn(y, yf)};
  float az[2] = {__fsub_rn(__fadd_rn(zf, 1.f), z), __fsub_rn(z, zf)};
  int id[8];
  float w[8], sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int dx = (k >> 2) & 1, dy = (k >> 1) & 1, dz = k & 1;
    id[k] = find_coord(keys, vals, cap, b, ix + dx, iy + dy, iz + dz) - 1;
    w[k] = id[k] >= 0 ? __fmul_rn(__fmul_rn(ax[dx], ay[dy]), az[dz]) : 0.f;
    sum = __fadd_rn(sum, w[k]);
  }
  float den = __fadd_rn(sum, 1e-8f);
  int4* io = reinterpret_cast<int4*>(idx8 + (size_t)i * 8);
  float4* wo = reinterpret_cast<float4*>(w8 + (size_t)i * 8);
  io[0] = make_int4(id[0], id[1], id[2], id[3]);
  io[1] = make_int4(id[4], id[5], id[6], id[7]);
  wo[0] = make_float4(__fdiv_rn(w[0], den), __fdiv_rn(w[1], den), __fdiv_rn(w[2], den), __fdiv_rn(w[3], den));
  wo[1] = make_float4(__fdiv_rn(w[4], den), __fdiv_rn(w[5], den), __fdiv_rn(w[6], den), __fdiv_rn(w[7], den));
}

// ---- pair-major view of a kernel map ------------------------------------------------------------
// For sparse levels (few active offsets per voxel) the convolution runs pair-major: per offset k the
// valid (input row, output row) pairs are compacted into tiles of 128 pairs.  counts -> tile prefix ->
// ordered compaction (deterministic: pairs of an offset stay sorted by output row).
constexpr int kPairSeg = 4096;   // map entries per CTA

__global__ void __launch_bounds__(1024) k_pairs_count(const int* __restrict__ map_t, int ld, int n_cap,
                                                      const int* __restrict__ n_dev, int nseg,
                                                      int* __restrict__ counts /* [kvol][nseg] */) {
  __shared__ int s_sum[32];
  const int k = blockIdx.y, seg = blockIdx.x, n = resolve_n(n_dev, n_cap);
  const int* row = map_t + (size_t)k * ld;
  int c = 0;
  const int end = min(n, (seg + 1) * kPairSeg);
  for (int o = seg * kPairSeg + threadIdx.x; o < end; o += 1024) c += row[o] >= 0;
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    c = s_sum[threadIdx.x];
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (threadIdx.x == 0) counts[k * nseg + seg] = c;
  }
}

// per offset: exclusive prefix of its segment counts (in place) and its first tile; then the tile -> offset table
__global__ void k_pairs_plan(int* __restrict__ counts, int kvol, int nseg, int tile_cap, int* __restrict__ tile_base,
                             int* __restrict__ tile_k, int* __restrict__ n_tiles, int* __restrict__ n_pairs) {
  __shared__ int total[32];
  __shared__ int base[33];
  if ((int)threadIdx.x < kvol) {
    int run = 0;
    int* c = counts + threadIdx.x * nseg;
    for (int sgm = 0; sgm < nseg; ++sgm) {
      int v = c[sgm];
      c[sgm] = run;
      run += v;
    }
    total[threadIdx.x] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, pairs = 0;
    for (int k = 0; k < kvol; ++k) {
      base[k] = run;
      tile_base[k] = run;
      run += (total[k] + 127) >> 7;
      pairs += total[k];
    }
    base[kvol] = run;
    *n_tiles = run < tile_cap ? run : tile_cap;
    *n_pairs = pairs;
  }
  __syncthreads();
  for (int k = 0; k < kvol; ++k)
    for (int t = base[k] + threadIdx.x; t < base[k + 1] && t < tile_cap; t += blockDim.x) tile_k[t] = k;
}

__global__ void __launch_bounds__(1024) k_pairs_fill(const int* __restrict__ map_t, int ld, int n_cap,
                                                     const int* __restrict__ n_dev, int nseg,
                                                     const int* __restrict__ seg_base, const int* __restrict__ tile_base,
                                                     int tile_cap, int* __restrict__ pair_in, int* __restrict__ pair_out) {
  __shared__ int s_warp[32];
  __shared__ int s_run;
  const int k = blockIdx.y, seg = blockIdx.x, n = resolve_n(n_dev, n_cap);
  const int* row = map_t + (size_t)k * ld;
  const long long cap = (long long)tile_cap * 128;
  const long long start = (long long)tile_base[k] * 128 + seg_base[k * nseg + seg];
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int end = min(n, (seg + 1) * kPairSeg);
  for (int o0 = seg * kPairSeg; o0 < end; o0 += 1024) {
    const int o = o0 + threadIdx.x;
    const int v = o < end ? row[o] : -1;
    const unsigned bal = __ballot_sync(0xffffffffu, v >= 0);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = s_warp[w];
      if (w < warp) wbase += c;
      total += c;
    }
    const int run = s_run;
    if (v >= 0) {
      const long long pos = start + run + wbase + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) {
        pair_in[pos] = v;
        pair_out[pos] = o;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_run = run + total;
    __syncthreads();
  }
  // the CTA of the last segment that holds rows pads the tail of the offset's last tile
  if (seg == (n > 0 ? (n - 1) / kPairSeg : 0)) {
    const int cnt = seg_base[k * nseg + seg] + s_run;   // pairs of this offset in total
    const int padded = (cnt + 127) & ~127;
    for (int j = cnt + threadIdx.x; j < padded; j += 1024) {
      const long long pos = (long long)tile_base[k] * 128 + j;
      if (pos < cap) {
        pair_in[pos] = -1;
        pair_out[pos] = -1;
      }
    }
  }
}

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_kmap_pairs(const int32_t* map_t, int ld, int n_cap, const int32_t* n_dev, int kvol, int tile_cap,
                               int32_t* pair_in, int32_t* pair_out, int32_t* tile_k, int32_t* n_tiles, int32_t* n_pairs,
                               int32_t* scratch, spvd_stream_t stream) {
  SPVD_CHECK_ARG(map_t && pair_in && pair_out && tile_k && n_tiles && n_pairs && scratch && kvol >= 1 && kvol <= 32 &&
                     tile_cap > 0,
                 "spvd_kmap_pairs: bad arguments");
  // scratch: 32 + kvol * ceil(n_cap / 4096) int32
  cudaStream_t s = (cudaStream_t)stream;
  const int nseg = cdiv(n_cap > 0 ? n_cap : 1, kPairSeg);
  int* tile_base = scratch;        // [32]
  int* counts = scratch + 32;      // [kvol][nseg]
  k_pairs_count<<<dim3(nseg, kvol), 1024, 0, s>>>(map_t, ld, n_cap, n_dev, nseg, counts);
  k_pairs_plan<<<1, 256, 0, s>>>(counts, kvol, nseg, tile_cap, tile_base, tile_k, n_tiles, n_pairs);
  k_pairs_fill<<<dim3(nseg, kvol), 1024, 0, s>>>(map_t, ld, n_cap, n_dev, nseg, counts, tile_base, tile_cap, pair_in,
                                                 pair_out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_hash_build(const int32_t* coords, int n_cap, const int32_t* n_dev, int coord_order,
                               int64_t* keys, int32_t* vals, int capacity, int32_t* status, spvd_stream_t stream) {
  SPVD_CHECK_ARG(coords && keys && vals && n_cap >= 0, "spvd_hash_build: bad arguments");
  SPVD_CHECK_ARG(capacity >= 2 * n_cap && capacity > 0, "spvd_hash_build: capacity %d < 2*%d", capacity, n_cap);
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(keys, 0, (size_t)capacity * 8, s));
  SPVD_CUDA(cudaMemsetAsync(vals, 0, (size_t)capacity * 4, s));
  if (n_cap == 0) return SPVD_OK;
  k_hash_insert<<<cdiv(n_cap, 256), 256, 0, s>>>(coords, n_cap, n_dev, coord_order, (u64*)keys, vals,
                                                 (unsigned)capacity, status);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_hash_query(const int32_t* query, int m, int coord_order, int kernel_size, const int64_t* keys,
                               const int32_t* vals, int capacity, int32_t* out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(query && keys && vals && out && m >= 0 && capacity > 0, "spvd_hash_query: bad arguments");
  SPVD_CHECK_ARG(kernel_size >= 1 && kernel_size <= 3, "spvd_hash_query: kernel_size %d not in 1..3", kernel_size);
  if (m == 0) return SPVD_OK;
  int kvol = kernel_size * kernel_size * kernel_size;
  long long total = (long long)m * kvol;
  k_hash_query<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(query, m, coord_order, kernel_size, kvol,
                                                                   (const u64*)keys, vals, (unsigned)capacity, out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_kmap_subm3(const int32_t* coords, int n_cap, const int32_t* n_dev, const int64_t* keys,
                               const int32_t* vals, int capacity, int32_t* map_t, int ld, uint32_t* tile_mask,
                               spvd_stream_t stream) {
  SPVD_CHECK_ARG(coords && keys && vals && map_t && capacity > 0, "spvd_kmap_subm3: bad arguments");
  SPVD_CHECK_ARG(ld >= n_cap && ld % 128 == 0, "spvd_kmap_subm3: ld %d must be a multiple of 128 >= %d", ld, n_cap);
  if (ld == 0) return SPVD_OK;
  k_kmap_subm3<<<ld / 128, 384, 0, (cudaStream_t)stream>>>((const int4*)coords, n_cap, n_dev, (const u64*)keys, vals,
                                                           (unsigned)capacity, map_t, ld, tile_mask);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_kmap_down2(const int32_t* fine_coords, int n_fine_cap, const int32_t* n_fine_dev,
                               const int64_t* keys, const int32_t* vals, int capacity, int32_t* map_down_t,
                               int ld_coarse, uint32_t* tile_mask_down, int32_t* map_up_t, int ld_fine,
                               uint32_t* tile_mask_up, int32_t* parent, int32_t* koff, spvd_stream_t stream) {
  SPVD_CHECK_ARG(fine_coords && keys && vals && capacity > 0 && n_fine_cap >= 0, "spvd_kmap_down2: bad arguments");
  SPVD_CHECK_ARG(!map_down_t || (ld_coarse > 0 && ld_coarse % 128 == 0), "spvd_kmap_down2: bad ld_coarse %d", ld_coarse);
  SPVD_CHECK_ARG(!map_up_t || (ld_fine >= n_fine_cap && ld_fine % 128 == 0), "spvd_kmap_down2: bad ld_fine %d", ld_fine);
  cudaStream_t s = (cudaStream_t)stream;
  if (map_down_t) SPVD_CUDA(cudaMemsetAsync(map_down_t, 0xFF, (size_t)8 * ld_coarse * 4, s));
  if (tile_mask_down) SPVD_CUDA(cudaMemsetAsync(tile_mask_down, 0, (size_t)(ld_coarse / 128) * 4, s));
  if (tile_mask_up) SPVD_CUDA(cudaMemsetAsync(tile_mask_up, 0, (size_t)(ld_fine / 128) * 4, s));
  int span = map_up_t ? (ld_fine > n_fine_cap ? ld_fine : n_fine_cap) : n_fine_cap;
  if (span == 0) return SPVD_OK;
  k_kmap_down2<<<cdiv(span, 256), 256, 0, s>>>((const int4*)fine_coords, n_fine_cap, n_fine_dev, (const u64*)keys,
                                               vals, (unsigned)capacity, map_down_t, ld_coarse, tile_mask_down,
                                               map_up_t, map_up_t ? ld_fine : 0, tile_mask_up, parent, koff);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_point_voxel_query(const float* pc, int n, int stride, const int64_t* keys, const int32_t* vals,
                                      int capacity, int32_t* idx, spvd_stream_t stream) {
  SPVD_CHECK_ARG(pc && keys && vals && idx && n >= 0 && stride >= 1 && capacity > 0, "spvd_point_voxel_query: bad arguments");
  if (n == 0) return SPVD_OK;
  k_point_voxel_query<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)pc, n, (float)stride,
                                                                      (const u64*)keys, vals, (unsigned)capacity, idx);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_devox_query(const float* pc, int n, int stride, const int64_t* keys, const int32_t* vals,
                                int capacity, int32_t* idx8, float* w8, spvd_stream_t stream) {
  SPVD_CHECK_ARG(pc && keys && vals && idx8 && w8 && n >= 0 && stride >= 1 && capacity > 0, "spvd_devox_query: bad arguments");
  if (n == 0) return SPVD_OK;
  k_devox_query<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>((const float4*)pc, n, (float)stride, (const u64*)keys,
                                                                vals, (unsigned)capacity, idx8, w8);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
