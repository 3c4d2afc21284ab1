// Small fused elementwise kernels around the sparse convolutions (inference path): BatchNorm folded
// to a per-channel affine + SiLU (spnn.BatchNorm + spnn.SiLU, models/ddpm_unet_attn.py:24-28), the
// FiLM time modulation (models/modules/graph.py:19-41) and per-cloud voxel counts (the offsets that
// models/modules/transformer.py:12-15 gets from to_dense_batch).  HBM-bound, 128-bit accesses.
#include "common.cuh"

namespace spvd {

__device__ __forceinline__ float silu(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ float rna_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// y = act(x * scale[c] + shift[c]);  act bit 0: SiLU, bit 1: round the result to TF32 (rna), bit 2: store y as f16
// kPer float4 per thread (a block's elements 256 apart, all loads issued before the first store): large tensors stream
// faster with two loads in flight per thread, small ones want as many threads as they can get
template <int kPer>
__global__ void __launch_bounds__(256) k_affine_act4(const float4* __restrict__ x, const float4* __restrict__ scale,
                                                     const float4* __restrict__ shift, long long total4, int c4, int act,
                                                     float4* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const long long t0 = (long long)blockIdx.x * (256 * kPer) + threadIdx.x;
  float4 v[kPer];
#pragma unroll
  for (int e = 0; e < kPer; ++e)
    if (t0 + e * 256 < total4) v[e] = x[t0 + e * 256];
#pragma unroll
  for (int e = 0; e < kPer; ++e) {
    const long long t = t0 + e * 256;
    if (t >= total4) break;
    const int q = (int)(t % c4);
    const float4 s = __ldg(scale + q), b = __ldg(shift + q);
    float4 r = v[e];
    r.x = fmaf(r.x, s.x, b.x);
    r.y = fmaf(r.y, s.y, b.y);
    r.z = fmaf(r.z, s.z, b.z);
    r.w = fmaf(r.w, s.w, b.w);
    if (act & 1) {
      r.x = silu(r.x);
      r.y = silu(r.y);
      r.z = silu(r.z);
      r.w = silu(r.w);
    }
    if (act & 4) {
      store_half4(y, t, r);
      continue;
    }
    if (act & 2) {
      r.x = rna_tf32(r.x);
      r.y = rna_tf32(r.y);
      r.z = rna_tf32(r.z);
      r.w = rna_tf32(r.w);
    }
    y[t] = r;
  }
}

__global__ void k_affine_act1(const float* __restrict__ x, const float* __restrict__ scale,
                              const float* __restrict__ shift, long long total, int c, int act,
                              float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int q = (int)(t % c);
  float v = fmaf(x[t], scale[q], shift[q]);
  if (act & 1) v = silu(v);
  if (act & 2) v = rna_tf32(v);
  y[t] = v;
}

// y[i] = (1 + tt[b_i, c]) * x[i, c] + tt[b_i, C + c],  b_i = coords[i].batch
__global__ void k_film4(const float4* __restrict__ x, const float* __restrict__ tt, const int4* __restrict__ coords,
                        long long total4, int c4, float4* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total4) return;
  int row = (int)(t / c4), q = (int)(t % c4);
  int b = coords[row].x;
  const float4* tb = reinterpret_cast<const float4*>(tt + (size_t)b * c4 * 8);
  float4 v = x[t], sc = tb[q], sh = tb[c4 + q];     // (the FiLM table is an activation: plain loads under PDL)
  v.x = fmaf(1.f + sc.x, v.x, sh.x);
  v.y = fmaf(1.f + sc.y, v.y, sh.y);
  v.z = fmaf(1.f + sc.z, v.z, sh.z);
  v.w = fmaf(1.f + sc.w, v.w, sh.w);
  y[t] = v;
}

// rows are sorted by (batch, x, y, z): counts[b] = lower_bound(b+1) - lower_bound(b), one thread per cloud
__device__ __forceinline__ int batch_lower_bound(const int4* __restrict__ coords, int n, int b) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (coords[mid].x < b) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void k_batch_counts(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, int nb,
                               int* __restrict__ counts) {
  int n = resolve_n(n_dev, n_cap);
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  counts[b] = batch_lower_bound(coords, n, b + 1) - batch_lower_bound(coords, n, b);
}

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_bn_silu_fwd(const float* x, const float* scale, const float* shift, long long rows, int c, int act,
                                float* y, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && scale && shift && y && rows >= 0 && c > 0, "spvd_bn_silu_fwd: bad arguments");
  if (rows == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  bool v4 = c % 4 == 0 && (((size_t)x | (size_t)y | (size_t)scale | (size_t)shift) & 15) == 0;
  SPVD_CHECK_ARG(v4 || !(act & 4), "spvd_bn_silu_fwd: f16 output needs c %% 4 == 0 and 16-byte aligned pointers");
  if (v4) {
    if (rows * (c / 4) >= (1 << 20))
      SPVD_CUDA(launch_pdl(k_affine_act4<2>, dim3(cdiv(rows * (c / 4), 512)), dim3(256), 0, s, (const float4*)x,
                           (const float4*)scale, (const float4*)shift, rows * (c / 4), c / 4, act, (float4*)y));
    else
      SPVD_CUDA(launch_pdl(k_affine_act4<1>, dim3(cdiv(rows * (c / 4), 256)), dim3(256), 0, s, (const float4*)x,
                           (const float4*)scale, (const float4*)shift, rows * (c / 4), c / 4, act, (float4*)y));
  } else
    SPVD_CUDA(launch_pdl(k_affine_act1, dim3(cdiv(rows * c, 256)), dim3(256), 0, s, x, scale, shift, rows * c, c, act, y));
  return SPVD_OK;
}

extern "C" int spvd_film_fwd(const float* x, const float* tt, const int32_t* coords, long long rows, int c, float* y,
                             spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && tt && coords && y && rows >= 0 && c > 0 && c % 4 == 0, "spvd_film_fwd: bad arguments (c %% 4 == 0)");
  SPVD_CHECK_ARG((((size_t)x | (size_t)y | (size_t)tt) & 15) == 0, "spvd_film_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return SPVD_OK;
  SPVD_CUDA(launch_pdl(k_film4, dim3(cdiv(rows * (c / 4), 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, tt,
                       (const int4*)coords, rows * (c / 4), c / 4, (float4*)y));
  return SPVD_OK;
}

extern "C" int spvd_batch_counts(const int32_t* coords, int n_cap, const int32_t* n_dev, int nb, int32_t* counts,
                                 spvd_stream_t stream) {
  SPVD_CHECK_ARG(coords && counts && n_cap >= 0 && nb > 0, "spvd_batch_counts: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  k_batch_counts<<<cdiv(nb, 128), 128, 0, s>>>((const int4*)coords, n_cap, n_dev, nb, counts);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
