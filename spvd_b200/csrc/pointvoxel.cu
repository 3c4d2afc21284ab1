// Point <-> voxel feature transfer: torch_scatter.scatter_mean (models/sparse_utils.py:69,94) and
// torchsparse spdevoxelize (models/sparse_utils.py:119,128), forward and backward.
//
// HBM/L2-bound gather/scatter.  Layout: features are row-major [rows, C] fp32; a thread owns one
// (row, 4-channel) slot when C % 4 == 0 (128-bit loads/stores, vector RED for the scatters) and one
// (row, channel) element otherwise (C = 3 at the network input).
#include "common.cuh"

namespace spvd {

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// ---- scatter_mean ---------------------------------------------------------------------------
__global__ void k_scatter_add4(const float4* __restrict__ src, const int* __restrict__ idx, int n, int c4,
                               float* __restrict__ out, float* __restrict__ counts) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c4) return;
  int p = (int)(t / c4), q = (int)(t % c4);
  int v = idx[p];
  if (v < 0) return;
  red_add_v4(out + ((size_t)v * c4 + q) * 4, src[t]);
  if (q == 0) atomicAdd(counts + v, 1.0f);
}

__global__ void k_scatter_add1(const float* __restrict__ src, const int* __restrict__ idx, int n, int c,
                               float* __restrict__ out, float* __restrict__ counts) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  int p = (int)(t / c), q = (int)(t % c);
  int v = idx[p];
  if (v < 0) return;
  atomicAdd(out + (size_t)v * c + q, src[t]);
  if (q == 0) atomicAdd(counts + v, 1.0f);
}

__global__ void k_divide_rows(float* __restrict__ out, const float* __restrict__ counts, int nv, int c) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nv * c) return;
  float cnt = fmaxf(counts[t / c], 1.0f);
  out[t] = __fdiv_rn(out[t], cnt);
}

// ---- scatter_mean with the 1/count of every point's voxel precomputed at planning time: one zero-fill + one RED pass,
// no count accumulation and no divide pass over the voxel rows ------------------------------------------------------
__global__ void k_voxel_point_count(const int* __restrict__ idx, int n, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = i < n ? idx[i] : -1;
  if (v < 0) return;
  const unsigned peers = __match_any_sync(__activemask(), v);       // one atomic per distinct voxel of the warp
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + v, __popc(peers));
}

__global__ void k_point_mean_weight(const int* __restrict__ idx, const int* __restrict__ counts, int n, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = idx[i];
  w[i] = v >= 0 ? __fdiv_rn(1.0f, (float)counts[v]) : 0.f;
}

__global__ void k_scatter_wadd4(const float4* __restrict__ src, const int* __restrict__ idx, const float* __restrict__ w, int n,
                                int c4, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c4) return;
  const int p = (int)(t / c4), q = (int)(t % c4);
  const int v = idx[p];
  if (v < 0) return;
  const float wp = w[p];
  const float4 x = src[t];
  red_add_v4(out + ((size_t)v * c4 + q) * 4, make_float4(x.x * wp, x.y * wp, x.z * wp, x.w * wp));
}

// ---- scatter_mean as a segmented reduction: the planner groups the points by voxel (disjoint ranges of an order array,
// handed out warp by warp from one cursor) and cuts every voxel into units of <= kUnit points; the feature pass sums each
// unit's rows with plain loads and adds its share of the mean with ONE vector RED per 4 channels.  At stride 16 a voxel
// holds ~35 points (hundreds early in sampling) and the row-0 sink of unmatched points thousands: per-point same-address
// RED serialises there (145 us for C = 256 in ncu). -------------------------------------------------------------------
constexpr int kUnit = 32;

// counts [nv] (from k_voxel_point_count), meta = {next free order slot, next free unit}
__global__ void k_group_alloc(const int* __restrict__ counts, const int* __restrict__ nv_dev, int nv_cap, int* __restrict__ meta,
                              int* __restrict__ start, int* __restrict__ cursor, int2* __restrict__ units) {
  const int nv = resolve_n(nv_dev, nv_cap);
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = v < nv ? counts[v] : 0;
  const int nu = (c + kUnit - 1) / kUnit;
  const int lane = threadIdx.x & 31;
  int incl = c, uincl = nu;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o), tu = __shfl_up_sync(0xffffffffu, uincl, o);
    if (lane >= o) incl += t, uincl += tu;
  }
  int base = 0, ubase = 0;
  if (lane == 31 && incl > 0) base = atomicAdd(meta, incl), ubase = atomicAdd(meta + 1, uincl);
  base = __shfl_sync(0xffffffffu, base, 31);
  ubase = __shfl_sync(0xffffffffu, ubase, 31);
  if (v < nv) {
    start[v] = cursor[v] = base + incl - c;
    const int u0 = ubase + uincl - nu;
    for (int u = 0; u < nu; ++u) units[u0 + u] = make_int2(v, u);
  }
}

__global__ void k_group_fill(const int* __restrict__ idx, int n, int* __restrict__ cursor, int* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = i < n ? idx[i] : -1;
  if (v < 0) return;
  const unsigned peers = __match_any_sync(__activemask(), v);
  const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(cursor + v, __popc(peers));
  base = __shfl_sync(peers, base, leader);
  order[base + __popc(peers & ((1u << lane) - 1u))] = i;
}

// one thread per (unit, 4 channels); the output was zero-filled
__global__ void k_segment_mean4(const float4* __restrict__ src, const int* __restrict__ start, const int* __restrict__ counts,
                                const int* __restrict__ order, const int2* __restrict__ units, const int* __restrict__ meta,
                                int c4, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int u = (int)(t / c4), q = (int)(t % c4);
  if (u >= meta[1]) return;
  const int2 rec = units[u];
  const int v = rec.x, cnt = counts[v];
  const int lo = start[v] + rec.y * kUnit, hi = min(start[v] + cnt, lo + kUnit);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = lo;
  for (; j + 4 <= hi; j += 4) {
    const int o0 = order[j], o1 = order[j + 1], o2 = order[j + 2], o3 = order[j + 3];
    const float4 f0 = src[(size_t)o0 * c4 + q], f1 = src[(size_t)o1 * c4 + q];
    const float4 f2 = src[(size_t)o2 * c4 + q], f3 = src[(size_t)o3 * c4 + q];
    acc.x += (f0.x + f1.x) + (f2.x + f3.x);
    acc.y += (f0.y + f1.y) + (f2.y + f3.y);
    acc.z += (f0.z + f1.z) + (f2.z + f3.z);
    acc.w += (f0.w + f1.w) + (f2.w + f3.w);
  }
  for (; j < hi; ++j) {
    const float4 f = src[(size_t)order[j] * c4 + q];
    acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
  }
  const float inv = __fdiv_rn(1.0f, (float)cnt);
  red_add_v4(out + ((size_t)v * c4 + q) * 4, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
}

__global__ void k_scatter_mean_bwd(const float* __restrict__ gout, const int* __restrict__ idx,
                                   const float* __restrict__ counts, int n, int c, float* __restrict__ gsrc) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  int p = (int)(t / c), q = (int)(t % c);
  int v = idx[p];
  gsrc[t] = v >= 0 ? __fdiv_rn(gout[(size_t)v * c + q], fmaxf(counts[v], 1.0f)) : 0.f;
}

// ---- devoxelize -----------------------------------------------------------------------------
// kSlots float4 channel slots per thread (q, q + c4/kSlots, ...): the 8 corner indices / weights are loaded once and
// 8 * kSlots gathers are in flight per thread -- with one slot the kernel is bound by (dependent-load latency) x (waves)
template <int kSlots>
__global__ void k_devox_fwd4(const float4* __restrict__ feats, const int* __restrict__ idx8,
                             const float* __restrict__ w8, int n, int c4, float4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int cq = c4 / kSlots;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * cq) return;
  int p = (int)(t / cq), q = (int)(t % cq);
  const int4* ip = reinterpret_cast<const int4*>(idx8 + (size_t)p * 8);
  const float4* wp = reinterpret_cast<const float4*>(w8 + (size_t)p * 8);
  int4 i0 = ip[0], i1 = ip[1];
  float4 w0 = wp[0], w1 = wp[1];
  int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
  float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  float4 acc[kSlots];
#pragma unroll
  for (int h = 0; h < kSlots; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (id[j] >= 0) {
#pragma unroll
      for (int h = 0; h < kSlots; ++h) {
        float4 f = feats[(size_t)id[j] * c4 + q + h * cq];
        acc[h].x = fmaf(w[j], f.x, acc[h].x);
        acc[h].y = fmaf(w[j], f.y, acc[h].y);
        acc[h].z = fmaf(w[j], f.z, acc[h].z);
        acc[h].w = fmaf(w[j], f.w, acc[h].w);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < kSlots; ++h) out[(size_t)p * c4 + q + h * cq] = acc[h];
}

__global__ void k_devox_fwd1(const float* __restrict__ feats, const int* __restrict__ idx8,
                             const float* __restrict__ w8, int n, int c, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  int p = (int)(t / c), q = (int)(t % c);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int id = idx8[(size_t)p * 8 + j];
    if (id >= 0) acc = fmaf(w8[(size_t)p * 8 + j], feats[(size_t)id * c + q], acc);
  }
  out[t] = acc;
}

__global__ void k_devox_bwd4(const float4* __restrict__ gout, const int* __restrict__ idx8,
                             const float* __restrict__ w8, int n, int c4, float* __restrict__ gfeats) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c4) return;
  int p = (int)(t / c4), q = (int)(t % c4);
  float4 g = gout[t];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int id = idx8[(size_t)p * 8 + j];
    if (id >= 0) {
      float w = w8[(size_t)p * 8 + j];
      red_add_v4(gfeats + ((size_t)id * c4 + q) * 4, make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
    }
  }
}

__global__ void k_devox_bwd1(const float* __restrict__ gout, const int* __restrict__ idx8,
                             const float* __restrict__ w8, int n, int c, float* __restrict__ gfeats) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  int p = (int)(t / c), q = (int)(t % c);
  float g = gout[t];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int id = idx8[(size_t)p * 8 + j];
    if (id >= 0) atomicAdd(gfeats + (size_t)id * c + q, w8[(size_t)p * 8 + j] * g);
  }
}

static inline bool aligned16(const void* p) { return ((size_t)p & 15) == 0; }

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_scatter_mean_fwd(const float* src, const int32_t* idx, int n, int c, float* out, float* counts,
                                     int nv, spvd_stream_t stream) {
  SPVD_CHECK_ARG(src && idx && out && counts && n >= 0 && c > 0 && nv >= 0, "spvd_scatter_mean_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (nv == 0) return SPVD_OK;
  SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)nv * c * 4, s));
  SPVD_CUDA(cudaMemsetAsync(counts, 0, (size_t)nv * 4, s));
  if (n > 0) {
    if (c % 4 == 0 && aligned16(src) && aligned16(out))
      k_scatter_add4<<<cdiv((long long)n * (c / 4), 256), 256, 0, s>>>((const float4*)src, idx, n, c / 4, out, counts);
    else
      k_scatter_add1<<<cdiv((long long)n * c, 256), 256, 0, s>>>(src, idx, n, c, out, counts);
  }
  SPVD_CUDA(launch_pdl(k_divide_rows, dim3(cdiv((long long)nv * c, 256)), dim3(256), 0, s, out, counts, nv, c));
  return SPVD_OK;
}

extern "C" int spvd_point_mean_weights(const int32_t* idx, int n, int nv_cap, int32_t* counts, float* w, spvd_stream_t stream) {
  SPVD_CHECK_ARG(idx && counts && w && n >= 0 && nv_cap >= 0, "spvd_point_mean_weights: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(counts, 0, (size_t)nv_cap * 4, s));
  if (n == 0) return SPVD_OK;
  k_voxel_point_count<<<cdiv(n, 256), 256, 0, s>>>(idx, n, counts);
  k_point_mean_weight<<<cdiv(n, 256), 256, 0, s>>>(idx, counts, n, w);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_group_points_by_voxel(const int32_t* idx, int n, int nv_cap, const int32_t* nv_dev, int32_t* counts,
                                          int32_t* start, int32_t* cursor, int32_t* order, int32_t* units, spvd_stream_t stream) {
  SPVD_CHECK_ARG(idx && counts && start && cursor && order && units && n >= 0 && nv_cap >= 0, "spvd_group_points_by_voxel: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(counts, 0, (size_t)(nv_cap + 2) * 4, s));   // counts [nv_cap] | next free slot | units
  if (n > 0) k_voxel_point_count<<<cdiv(n, 256), 256, 0, s>>>(idx, n, counts);
  if (nv_cap > 0) k_group_alloc<<<cdiv(nv_cap, 256), 256, 0, s>>>(counts, nv_dev, nv_cap, counts + nv_cap, start, cursor, (int2*)units);
  if (n > 0) k_group_fill<<<cdiv(n, 256), 256, 0, s>>>(idx, n, cursor, order);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_segment_mean_fwd(const float* src, const int32_t* start, const int32_t* counts, const int32_t* order,
                                     const int32_t* units, int n, int nv_cap, int c, float* out, int nv, spvd_stream_t stream) {
  SPVD_CHECK_ARG(src && start && counts && order && units && out && c > 0 && c % 4 == 0 && nv >= 0 && nv <= nv_cap && n >= 0,
                 "spvd_segment_mean_fwd: bad arguments (c %% 4 == 0)");
  SPVD_CHECK_ARG(aligned16(src) && aligned16(out), "spvd_segment_mean_fwd: pointers must be 16-byte aligned");
  if (nv == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)nv * c * 4, s));
  const long long max_units = (long long)nv + n / 32;                  // ceil(cnt / 32) summed over the voxels
  SPVD_CUDA(launch_pdl(k_segment_mean4, dim3(cdiv(max_units * (c / 4), 256)), dim3(256), 0, s, (const float4*)src, start, counts,
                       order, (const int2*)units, counts + nv_cap, c / 4, out));
  return SPVD_OK;
}

extern "C" int spvd_scatter_wmean_fwd(const float* src, const int32_t* idx, const float* w, int n, int c, float* out, int nv,
                                      spvd_stream_t stream) {
  SPVD_CHECK_ARG(src && idx && w && out && n >= 0 && c > 0 && c % 4 == 0 && nv >= 0, "spvd_scatter_wmean_fwd: bad arguments (c %% 4 == 0)");
  SPVD_CHECK_ARG(aligned16(src) && aligned16(out), "spvd_scatter_wmean_fwd: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (nv == 0) return SPVD_OK;
  SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)nv * c * 4, s));
  if (n > 0)
    SPVD_CUDA(launch_pdl(k_scatter_wadd4, dim3(cdiv((long long)n * (c / 4), 256)), dim3(256), 0, s, (const float4*)src, idx, w, n,
                         c / 4, out));
  return SPVD_OK;
}

extern "C" int spvd_scatter_mean_bwd(const float* gout, const int32_t* idx, const float* counts, int n, int c,
                                     float* gsrc, spvd_stream_t stream) {
  SPVD_CHECK_ARG(gout && idx && counts && gsrc && n >= 0 && c > 0, "spvd_scatter_mean_bwd: bad arguments");
  if (n == 0) return SPVD_OK;
  k_scatter_mean_bwd<<<cdiv((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(gout, idx, counts, n, c, gsrc);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_devoxelize_fwd(const float* feats, const int32_t* idx8, const float* w8, int n, int c, float* out,
                                   spvd_stream_t stream) {
  SPVD_CHECK_ARG(feats && idx8 && w8 && out && n >= 0 && c > 0, "spvd_devoxelize_fwd: bad arguments");
  if (n == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (c % 4 == 0 && aligned16(feats) && aligned16(out))
  {
    const int c4 = c / 4;
    if (c4 % 4 == 0 && c4 >= 32)
      SPVD_CUDA(launch_pdl(k_devox_fwd4<4>, dim3(cdiv((long long)n * (c4 / 4), 256)), dim3(256), 0, s, (const float4*)feats, idx8,
                           w8, n, c4, (float4*)out));
    else if (c4 % 2 == 0)
      SPVD_CUDA(launch_pdl(k_devox_fwd4<2>, dim3(cdiv((long long)n * (c4 / 2), 256)), dim3(256), 0, s, (const float4*)feats, idx8,
                           w8, n, c4, (float4*)out));
    else
      SPVD_CUDA(launch_pdl(k_devox_fwd4<1>, dim3(cdiv((long long)n * c4, 256)), dim3(256), 0, s, (const float4*)feats, idx8,
                           w8, n, c4, (float4*)out));
  }
  else
    SPVD_CUDA(launch_pdl(k_devox_fwd1, dim3(cdiv((long long)n * c, 256)), dim3(256), 0, s, feats, idx8, w8, n, c, out));
  return SPVD_OK;
}

extern "C" int spvd_devoxelize_bwd(const float* gout, const int32_t* idx8, const float* w8, int n, int c,
                                   float* gfeats, int nv, spvd_stream_t stream) {
  SPVD_CHECK_ARG(gout && idx8 && w8 && gfeats && n >= 0 && c > 0 && nv >= 0, "spvd_devoxelize_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (nv == 0) return SPVD_OK;
  SPVD_CUDA(cudaMemsetAsync(gfeats, 0, (size_t)nv * c * 4, s));
  if (n == 0) return SPVD_OK;
  if (c % 4 == 0 && aligned16(gout) && aligned16(gfeats))
    k_devox_bwd4<<<cdiv((long long)n * (c / 4), 256), 256, 0, s>>>((const float4*)gout, idx8, w8, n, c / 4, gfeats);
  else
    k_devox_bwd1<<<cdiv((long long)n * c, 256), 256, 0, s>>>(gout, idx8, w8, n, c, gfeats);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
