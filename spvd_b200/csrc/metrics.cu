// Evaluation distances between point clouds: Chamfer (nearest neighbour in squared L2, both
// directions; metrics/chamfer_dist/chamfer.cu:15-229) and the approximate earth mover's distance (auction-style soft
// matching over ten temperature levels + squared-distance cost; metrics/PyTorchEMD/cuda/emd_kernel.cu:26-330).
//
// Both are O(N*M) per cloud pair on 3-float points: the other cloud is staged in shared memory 1024 points at a time as
// float4 (one 128-bit broadcast load per point instead of three scalar ones) and every thread owns one point.  Unlike the
// reference, which walks the batch with 32 CTAs, the grid covers (point tiles x batch), so a B = 32 x 2048 problem fills
// the GPU; the EMD levels are separate launches (three passes per level over device-resident state).
#include "common.cuh"

namespace spvd {

constexpr int kMT = 256;      // threads per CTA = points per CTA
constexpr int kTile = 1024;   // points of the other cloud per shared-memory tile

__device__ __forceinline__ void stage_points(const float* __restrict__ xyz, const float* __restrict__ extra, int base, int count,
                                             float4* s) {
  for (int l = threadIdx.x; l < count; l += blockDim.x) {
    const float* p = xyz + (size_t)(base + l) * 3;
    s[l] = make_float4(p[0], p[1], p[2], extra ? extra[base + l] : 0.f);
  }
}

// dist[b, i] = min_j |a_i - b_j|^2, idx = first argmin
__global__ void __launch_bounds__(kMT) k_chamfer_nn(const float* __restrict__ a, const float* __restrict__ bpts, int n, int m,
                                                    float* __restrict__ dist, int* __restrict__ idx) {
  __shared__ float4 s[kTile];
  const int b = blockIdx.y, i = blockIdx.x * kMT + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (i < n) {
    const float* p = a + ((size_t)b * n + i) * 3;
    x = p[0], y = p[1], z = p[2];
  }
  float best = INFINITY;
  int arg = 0;
  for (int l0 = 0; l0 < m; l0 += kTile) {
    const int cnt = min(kTile, m - l0);
    __syncthreads();
    stage_points(bpts + (size_t)b * m * 3, nullptr, l0, cnt, s);
    __syncthreads();
    for (int l = 0; l < cnt; ++l) {
      const float4 q = s[l];
      const float dx = q.x - x, dy = q.y - y, dz = q.z - z;
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < best) best = d, arg = l0 + l;
    }
  }
  if (i < n) {
    dist[(size_t)b * n + i] = best;
    idx[(size_t)b * n + i] = arg;
  }
}

// d(dist_i)/d a_i = 2 (a_i - b_idx), d/d b_idx = -that
__global__ void k_chamfer_bwd(const float* __restrict__ a, const float* __restrict__ bpts, const int* __restrict__ idx,
                              const float* __restrict__ g, int n, int m, float* __restrict__ ga, float* __restrict__ gb) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = idx[(size_t)b * n + i];
  const float* p = a + ((size_t)b * n + i) * 3;
  const float* q = bpts + ((size_t)b * m + j) * 3;
  const float w = 2.f * g[(size_t)b * n + i];
  for (int c = 0; c < 3; ++c) {
    const float v = w * (p[c] - q[c]);
    atomicAdd(ga + ((size_t)b * n + i) * 3 + c, v);
    atomicAdd(gb + ((size_t)b * m + j) * 3 + c, -v);
  }
}

// ---- approxmatch, one temperature level = three passes ------------------------------------------------------------
// pass 1: ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level d_kl) remainR[l])
__global__ void __launch_bounds__(kMT) k_emd_ratio_l(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n, int m,
                                                     float level, const float* __restrict__ remain_l,
                                                     const float* __restrict__ remain_r, float* __restrict__ ratio_l) {
  __shared__ float4 s[kTile];
  const int b = blockIdx.y, k = blockIdx.x * kMT + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (k < n) {
    const float* p = xyz1 + ((size_t)b * n + k) * 3;
    x = p[0], y = p[1], z = p[2];
  }
  float sum = 1e-9f;
  for (int l0 = 0; l0 < m; l0 += kTile) {
    const int cnt = min(kTile, m - l0);
    __syncthreads();
    stage_points(xyz2 + (size_t)b * m * 3, remain_r + (size_t)b * m, l0, cnt, s);
    __syncthreads();
    for (int l = 0; l < cnt; ++l) {
      const float4 q = s[l];
      const float dx = q.x - x, dy = q.y - y, dz = q.z - z;
      sum += __expf(level * (dx * dx + dy * dy + dz * dz)) * q.w;
    }
  }
  if (k < n) ratio_l[(size_t)b * n + k] = remain_l[(size_t)b * n + k] / sum;
}

// pass 2: what every right point receives, its consumption ratio, and its remainder
__global__ void __launch_bounds__(kMT) k_emd_ratio_r(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n, int m,
                                                     float level, const float* __restrict__ ratio_l, float* __restrict__ remain_r,
                                                     float* __restrict__ ratio_r) {
  __shared__ float4 s[kTile];
  const int b = blockIdx.y, l = blockIdx.x * kMT + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (l < m) {
    const float* p = xyz2 + ((size_t)b * m + l) * 3;
    x = p[0], y = p[1], z = p[2];
  }
  float sum = 0.f;
  for (int k0 = 0; k0 < n; k0 += kTile) {
    const int cnt = min(kTile, n - k0);
    __syncthreads();
    stage_points(xyz1 + (size_t)b * n * 3, ratio_l + (size_t)b * n, k0, cnt, s);
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      const float4 q = s[k];
      const float dx = x - q.x, dy = y - q.y, dz = z - q.z;
      sum += __expf(level * (dx * dx + dy * dy + dz * dz)) * q.w;
    }
  }
  if (l < m) {
    const float r = remain_r[(size_t)b * m + l];
    sum *= r;
    const float consumption = fminf(r / (sum + 1e-9f), 1.0f);
    ratio_r[(size_t)b * m + l] = consumption * r;
    remain_r[(size_t)b * m + l] = fmaxf(0.0f, r - sum);
  }
}

// pass 3: match[b, l, k] += exp(level d_kl) ratioL[k] ratioR[l]; remainL[k] -= its row sum
__global__ void __launch_bounds__(kMT) k_emd_match(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n, int m,
                                                   float level, const float* __restrict__ ratio_l, const float* __restrict__ ratio_r,
                                                   float* __restrict__ remain_l, float* __restrict__ match) {
  __shared__ float4 s[kTile];
  const int b = blockIdx.y, k = blockIdx.x * kMT + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f, rl = 0.f;
  if (k < n) {
    const float* p = xyz1 + ((size_t)b * n + k) * 3;
    x = p[0], y = p[1], z = p[2];
    rl = ratio_l[(size_t)b * n + k];
  }
  float sum = 0.f;
  float* mrow = match + (size_t)b * n * m + k;
  for (int l0 = 0; l0 < m; l0 += kTile) {
    const int cnt = min(kTile, m - l0);
    __syncthreads();
    stage_points(xyz2 + (size_t)b * m * 3, ratio_r + (size_t)b * m, l0, cnt, s);
    __syncthreads();
    if (k < n)
      for (int l = 0; l < cnt; ++l) {
        const float4 q = s[l];
        const float dx = q.x - x, dy = q.y - y, dz = q.z - z;
        const float w = __expf(level * (dx * dx + dy * dy + dz * dz)) * rl * q.w;
        mrow[(size_t)(l0 + l) * n] += w;        // coalesced across the threads (consecutive k)
        sum += w;
      }
  }
  if (k < n) remain_l[(size_t)b * n + k] = fmaxf(0.0f, remain_l[(size_t)b * n + k] - sum);
}

__global__ void k_fill(float* p, long long count, float v) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) p[t] = v;
}

// cost[b] = sum_kl match[b, l, k] d_kl ; one CTA per (tile of k, b), block reduce + atomic
__global__ void __launch_bounds__(kMT) k_emd_cost(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                  const float* __restrict__ match, int n, int m, float* __restrict__ cost) {
  __shared__ float4 s[kTile];
  __shared__ float red[kMT / 32];
  const int b = blockIdx.y, k = blockIdx.x * kMT + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (k < n) {
    const float* p = xyz1 + ((size_t)b * n + k) * 3;
    x = p[0], y = p[1], z = p[2];
  }
  float sum = 0.f;
  const float* mrow = match + (size_t)b * n * m + k;
  for (int l0 = 0; l0 < m; l0 += kTile) {
    const int cnt = min(kTile, m - l0);
    __syncthreads();
    stage_points(xyz2 + (size_t)b * m * 3, nullptr, l0, cnt, s);
    __syncthreads();
    if (k < n)
      for (int l = 0; l < cnt; ++l) {
        const float4 q = s[l];
        const float dx = q.x - x, dy = q.y - y, dz = q.z - z;
        sum += (dx * dx + dy * dy + dz * dz) * mrow[(size_t)(l0 + l) * n];
      }
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kMT / 32; ++w) t += red[w];
    atomicAdd(cost + b, t);
  }
}

// grad1[b, k] = g_b * sum_l 2 match[l, k] (x1_k - x2_l); grad2[b, l] = g_b * sum_k 2 match[l, k] (x2_l - x1_k)
__global__ void __launch_bounds__(kMT) k_emd_grad1(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                   const float* __restrict__ match, const float* __restrict__ g, int n, int m,
                                                   float* __restrict__ grad1) {
  __shared__ float4 s[kTile];
  const int b = blockIdx.y, k = blockIdx.x * kMT + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (k < n) {
    const float* p = xyz1 + ((size_t)b * n + k) * 3;
    x = p[0], y = p[1], z = p[2];
  }
  float gx = 0.f, gy = 0.f, gz = 0.f;
  const float* mrow = match + (size_t)b * n * m + k;
  for (int l0 = 0; l0 < m; l0 += kTile) {
    const int cnt = min(kTile, m - l0);
    __syncthreads();
    stage_points(xyz2 + (size_t)b * m * 3, nullptr, l0, cnt, s);
    __syncthreads();
    if (k < n)
      for (int l = 0; l < cnt; ++l) {
        const float4 q = s[l];
        const float w = 2.f * mrow[(size_t)(l0 + l) * n];
        gx += (x - q.x) * w, gy += (y - q.y) * w, gz += (z - q.z) * w;
      }
  }
  if (k < n) {
    float* o = grad1 + ((size_t)b * n + k) * 3;
    o[0] = gx * g[b], o[1] = gy * g[b], o[2] = gz * g[b];
  }
}

// one warp per right point l: lanes stride over k (coalesced match row reads)
__global__ void __launch_bounds__(kMT) k_emd_grad2(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                   const float* __restrict__ match, const float* __restrict__ g, int n, int m,
                                                   float* __restrict__ grad2) {
  const int b = blockIdx.y, l = blockIdx.x * (kMT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (l >= m) return;
  const float* q = xyz2 + ((size_t)b * m + l) * 3;
  const float x = q[0], y = q[1], z = q[2];
  const float* mrow = match + (size_t)b * n * m + (size_t)l * n;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int k = lane; k < n; k += 32) {
    const float* p = xyz1 + ((size_t)b * n + k) * 3;
    const float w = 2.f * mrow[k];
    gx += (x - p[0]) * w, gy += (y - p[1]) * w, gz += (z - p[2]) * w;
  }
  for (int o = 16; o; o >>= 1) {
    gx += __shfl_xor_sync(0xffffffffu, gx, o);
    gy += __shfl_xor_sync(0xffffffffu, gy, o);
    gz += __shfl_xor_sync(0xffffffffu, gz, o);
  }
  if (lane == 0) {
    float* o = grad2 + ((size_t)b * m + l) * 3;
    o[0] = gx * g[b], o[1] = gy * g[b], o[2] = gz * g[b];
  }
}

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_chamfer_fwd(const float* xyz1, const float* xyz2, int n_batch, int n, int m, float* dist1, float* dist2,
                                int32_t* idx1, int32_t* idx2, spvd_stream_t stream) {
  SPVD_CHECK_ARG(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2 && n_batch >= 0 && n > 0 && m > 0, "spvd_chamfer_fwd: bad arguments");
  if (n_batch == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  k_chamfer_nn<<<dim3(cdiv(n, kMT), n_batch), kMT, 0, s>>>(xyz1, xyz2, n, m, dist1, idx1);
  k_chamfer_nn<<<dim3(cdiv(m, kMT), n_batch), kMT, 0, s>>>(xyz2, xyz1, m, n, dist2, idx2);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_chamfer_bwd(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2,
                                const float* gdist1, const float* gdist2, int n_batch, int n, int m, float* gxyz1, float* gxyz2,
                                spvd_stream_t stream) {
  SPVD_CHECK_ARG(xyz1 && xyz2 && idx1 && idx2 && gdist1 && gdist2 && gxyz1 && gxyz2 && n_batch >= 0 && n > 0 && m > 0,
                 "spvd_chamfer_bwd: bad arguments");
  if (n_batch == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(gxyz1, 0, (size_t)n_batch * n * 12, s));
  SPVD_CUDA(cudaMemsetAsync(gxyz2, 0, (size_t)n_batch * m * 12, s));
  k_chamfer_bwd<<<dim3(cdiv(n, 256), n_batch), 256, 0, s>>>(xyz1, xyz2, idx1, gdist1, n, m, gxyz1, gxyz2);
  k_chamfer_bwd<<<dim3(cdiv(m, 256), n_batch), 256, 0, s>>>(xyz2, xyz1, idx2, gdist2, m, n, gxyz2, gxyz1);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" size_t spvd_emd_workspace_floats(int n_batch, int n, int m) { return (size_t)n_batch * 2 * ((size_t)n + m); }

extern "C" int spvd_emd_approxmatch(const float* xyz1, const float* xyz2, int n_batch, int n, int m, float* match, float* workspace,
                                    spvd_stream_t stream) {
  SPVD_CHECK_ARG(xyz1 && xyz2 && match && workspace && n_batch >= 0 && n > 0 && m > 0, "spvd_emd_approxmatch: bad arguments");
  if (n_batch == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* remain_l = workspace;                          // [B, n]
  float* remain_r = remain_l + (size_t)n_batch * n;     // [B, m]
  float* ratio_l = remain_r + (size_t)n_batch * m;      // [B, n]
  float* ratio_r = ratio_l + (size_t)n_batch * n;       // [B, m]
  const float multi_l = n >= m ? 1.f : (float)(m / n), multi_r = n >= m ? (float)(n / m) : 1.f;   // integer ratios, as the reference
  SPVD_CUDA(cudaMemsetAsync(match, 0, (size_t)n_batch * n * m * 4, s));
  k_fill<<<cdiv((long long)n_batch * n, 256), 256, 0, s>>>(remain_l, (long long)n_batch * n, multi_l);
  k_fill<<<cdiv((long long)n_batch * m, 256), 256, 0, s>>>(remain_r, (long long)n_batch * m, multi_r);
  for (int j = 7; j >= -2; --j) {
    const float level = j == -2 ? 0.f : -powf(4.0f, (float)j);
    k_emd_ratio_l<<<dim3(cdiv(n, kMT), n_batch), kMT, 0, s>>>(xyz1, xyz2, n, m, level, remain_l, remain_r, ratio_l);
    k_emd_ratio_r<<<dim3(cdiv(m, kMT), n_batch), kMT, 0, s>>>(xyz1, xyz2, n, m, level, ratio_l, remain_r, ratio_r);
    k_emd_match<<<dim3(cdiv(n, kMT), n_batch), kMT, 0, s>>>(xyz1, xyz2, n, m, level, ratio_l, ratio_r, remain_l, match);
  }
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_emd_matchcost_fwd(const float* xyz1, const float* xyz2, const float* match, int n_batch, int n, int m,
                                      float* cost, spvd_stream_t stream) {
  SPVD_CHECK_ARG(xyz1 && xyz2 && match && cost && n_batch >= 0 && n > 0 && m > 0, "spvd_emd_matchcost_fwd: bad arguments");
  if (n_batch == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(cost, 0, (size_t)n_batch * 4, s));
  k_emd_cost<<<dim3(cdiv(n, kMT), n_batch), kMT, 0, s>>>(xyz1, xyz2, match, n, m, cost);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_emd_matchcost_bwd(const float* grad_cost, const float* xyz1, const float* xyz2, const float* match, int n_batch,
                                      int n, int m, float* grad1, float* grad2, spvd_stream_t stream) {
  SPVD_CHECK_ARG(grad_cost && xyz1 && xyz2 && match && grad1 && grad2 && n_batch >= 0 && n > 0 && m > 0,
                 "spvd_emd_matchcost_bwd: bad arguments");
  if (n_batch == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  k_emd_grad1<<<dim3(cdiv(n, kMT), n_batch), kMT, 0, s>>>(xyz1, xyz2, match, grad_cost, n, m, grad1);
  k_emd_grad2<<<dim3(cdiv(m, kMT / 32), n_batch), kMT, 0, s>>>(xyz1, xyz2, match, grad_cost, n, m, grad2);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
