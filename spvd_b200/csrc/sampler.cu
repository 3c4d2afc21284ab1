// Per-step work of the sampling loop outside the network: the DDPM update
// (utils/schedulers.py:78-89) fused with the re-sparsification of the new points
// (SparseSchedulerGPU.torch2sparse, utils/schedulers.py:249-257: per-cloud min shift, floor(c / pres) -> int32
// coordinates with the batch column prepended).  Two launches, no host sync; the reference does this with a CPU
// randn + H2D copy, ~15 eager torch kernels and a torch.unique sync per step.
#include "common.cuh"

namespace spvd {

__device__ __forceinline__ int float_order_key(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float float_from_key(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

// x_new = c1 * (x - c2 * eps) + sigma * z   (eps == NULL: x_new = x), and the per-cloud minima of x_new
__global__ void __launch_bounds__(256) k_ddpm_update_min(const float* __restrict__ x, const float* __restrict__ eps,
                                                         const float* __restrict__ z, float c1, float c2, float sigma,
                                                         int n_per_cloud, int f, float* __restrict__ x_new,
                                                         int* __restrict__ cloud_min) {
  const int b = blockIdx.y;
  float m[3] = {INFINITY, INFINITY, INFINITY};
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_per_cloud; p += gridDim.x * blockDim.x) {
    const size_t row = ((size_t)b * n_per_cloud + p) * f;
    for (int a = 0; a < f; ++a) {
      float v = x[row + a];
      if (eps) {
        v = __fmul_rn(__fsub_rn(v, __fmul_rn(c2, eps[row + a])), c1);
        if (z) v = __fadd_rn(v, __fmul_rn(sigma, z[row + a]));
      }
      if (x_new) x_new[row + a] = v;
      if (a < 3) m[a] = fminf(m[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float v = m[a];
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v != INFINITY) atomicMin(cloud_min + b * 3 + a, float_order_key(v));
  }
}

__global__ void k_quantize_points(const float* __restrict__ x, const int* __restrict__ cloud_min, int n_per_cloud,
                                  int f, float pres, long long total, int4* __restrict__ coords,
                                  float4* __restrict__ pc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / n_per_cloud);
  const float* p = x + (size_t)i * f;
  int q[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    q[a] = (int)floorf(__fdiv_rn(__fsub_rn(p[a], float_from_key(cloud_min[b * 3 + a])), pres));
  if (coords) coords[i] = make_int4(b, q[0], q[1], q[2]);
  if (pc) pc[i] = make_float4((float)b, (float)q[0], (float)q[1], (float)q[2]);
}

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_ddpm_step(const float* x_t, const float* eps, const float* z, float c1, float c2, float sigma,
                              int n_clouds, int n_per_cloud, int f, float pres, float* x_new, int32_t* coords, float* pc,
                              int32_t* scratch, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x_t && scratch && n_clouds > 0 && n_per_cloud > 0 && f >= 3 && pres > 0.f, "spvd_ddpm_step: bad arguments");
  SPVD_CHECK_ARG(!eps || x_new, "spvd_ddpm_step: the update needs an output buffer");
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(scratch, 0x7F, (size_t)n_clouds * 3 * 4, s));   // 0x7F7F7F7F: larger than any finite key
  dim3 grid(cdiv(n_per_cloud, 256) < 8 ? cdiv(n_per_cloud, 256) : 8, n_clouds);
  k_ddpm_update_min<<<grid, 256, 0, s>>>(x_t, eps, z, c1, c2, sigma, n_per_cloud, f, x_new, scratch);
  const long long total = (long long)n_clouds * n_per_cloud;
  k_quantize_points<<<cdiv(total, 256), 256, 0, s>>>(eps ? x_new : x_t, scratch, n_per_cloud, f, pres, total, (int4*)coords,
                                                     (float4*)pc);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
