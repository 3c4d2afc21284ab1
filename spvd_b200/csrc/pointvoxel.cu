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

__global__ void k_scatter_mean_bwd(const float* __restrict__ gout, const int* __restrict__ idx,
                                   const float* __restrict__ counts, int n, int c, float* __restrict__ gsrc) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c) return;
  int p = (int)(t / c), q = (int)(t % c);
  int v = idx[p];
  gsrc[t] = v >= 0 ? __fdiv_rn(gout[(size_t)v * c + q], fmaxf(counts[v], 1.0f)) : 0.f;
}

// ---- devoxelize -----------------------------------------------------------------------------
__global__ void k_devox_fwd4(const float4* __restrict__ feats, const int* __restrict__ idx8,
                             const float* __restrict__ w8, int n, int c4, float4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c4) return;
  int p = (int)(t / c4), q = (int)(t % c4);
  const int4* ip = reinterpret_cast<const int4*>(idx8 + (size_t)p * 8);
  const float4* wp = reinterpret_cast<const float4*>(w8 + (size_t)p * 8);
  int4 i0 = ip[0], i1 = ip[1];
  float4 w0 = wp[0], w1 = wp[1];
  int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
  float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (id[j] >= 0) {
      float4 f = feats[(size_t)id[j] * c4 + q];
      acc.x = fmaf(w[j], f.x, acc.x);
      acc.y = fmaf(w[j], f.y, acc.y);
      acc.z = fmaf(w[j], f.z, acc.z);
      acc.w = fmaf(w[j], f.w, acc.w);
    }
  }
  out[t] = acc;
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
    SPVD_CUDA(launch_pdl(k_devox_fwd4, dim3(cdiv((long long)n * (c / 4), 256)), dim3(256), 0, s, (const float4*)feats, idx8,
                         w8, n, c / 4, (float4*)out));
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
