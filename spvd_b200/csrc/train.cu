// Training-mode kernels of the denoiser's glue (row TR of the scope table): everything between the sparse convolutions of
// EmbResBlock / PointBlock / SparseAttention that the forward + backward pass needs, so that a training step runs no
// library (cuBLAS / cuDNN / ATen) kernel:
//   * BatchNorm1d with batch statistics (+ SiLU), forward and backward           models/ddpm_unet_attn.py:24-28
//   * FiLM backward (per-cloud segmented reduction of the scale / shift gradients) models/modules/graph.py:19-41
//   * LayerNorm forward (saving mean / rstd) and backward                         models/modules/transformer.py:102-125
//   * per-cloud multi-head attention backward                                     models/modules/transformer.py:40-67
//   * channel split (backward of torchsparse.cat), SiLU, column sums (bias gradients), embedding-row scatter,
//     sinusoidal timestep embedding, fused MSE loss + gradient                    models/spvd.py:232,380-392; models/utils.py:15-19
// All HBM-bound row-major [rows, C] passes: 128-bit accesses where C % 4 == 0, fp32 partial sums per CTA, fp64 atomics for
// the per-channel totals (65k rows of fp32 squares would otherwise lose the last digits of the variance).
#include <math.h>

#include "common.cuh"

namespace spvd {

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + __expf(-v)); }
__device__ __forceinline__ float silu_t(float v) { return v * sigmoid_f(v); }
__device__ __forceinline__ float dsilu_t(float v) {
  const float s = sigmoid_f(v);
  return s * (1.f + v * (1.f - s));
}

// ------------------------------------------------------------------------------------------------
// per-channel reduction skeleton over a row-major [rows, c] matrix: every thread owns one channel (thread t -> channel
// t % c of row lane t / c) and walks the rows of its CTA's chunk; the row lanes are then combined in shared memory and
// the CTA adds its partial sums to fp64 totals.  f(row, ch, &a, &b) yields the two summands.
// ------------------------------------------------------------------------------------------------
constexpr int kRedThreads = 256;
constexpr int kRedMaxCtas = 592;   // 4 per SM: the fp64 atomics of a launch stay at a few hundred per channel

static inline int red_rows_per_cta(long long rows) {
  long long per = (rows + kRedMaxCtas - 1) / kRedMaxCtas;
  if (per < 32) per = 32;
  return (int)per;
}

template <typename F>
__device__ __forceinline__ void col_reduce2(long long rows, int c, int rows_per_cta, double* __restrict__ tot, F f) {
  __shared__ float sa[kRedThreads], sb[kRedThreads];
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  const long long r_end = min(rows, r_begin + rows_per_cta);
  for (int c0 = 0; c0 < c; c0 += kRedThreads) {            // channel chunks of up to 256 (one pass for every SPVD width)
    const int cw = min(kRedThreads, c - c0);
    const int lanes = kRedThreads / cw;                    // row lanes
    const int ch = threadIdx.x % cw, lane = threadIdx.x / cw;
    float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f, a2 = 0.f, b2 = 0.f, a3 = 0.f, b3 = 0.f;
    if (lane < lanes) {
      long long r = r_begin + lane;
      for (; r + 3 * lanes < r_end; r += 4 * lanes) {      // four independent rows in flight per thread
        f(r, c0 + ch, &a0, &b0);
        f(r + lanes, c0 + ch, &a1, &b1);
        f(r + 2 * lanes, c0 + ch, &a2, &b2);
        f(r + 3 * lanes, c0 + ch, &a3, &b3);
      }
      for (; r < r_end; r += lanes) f(r, c0 + ch, &a0, &b0);
    }
    float a = (a0 + a1) + (a2 + a3), b = (b0 + b1) + (b2 + b3);
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x < cw) {
      for (int l = 1; l < lanes; ++l) {
        a += sa[threadIdx.x + l * cw];
        b += sb[threadIdx.x + l * cw];
      }
      atomicAdd(tot + c0 + ch, (double)a);
      atomicAdd(tot + c + c0 + ch, (double)b);
    }
    __syncthreads();
  }
}

// The same with four channels per thread (c % 4 == 0, 16-byte aligned rows): thread t -> channels 4 * (t % (c/4)) .. +3 of
// row lane t / (c/4), four rows in flight -> 64 bytes per thread instead of 16 outstanding (the scalar skeleton keeps
// ~16 KB per SM in flight and reaches ~1.2 TB/s; HBM wants a few times that).  f(row, ch, a, b) adds to float4 sums.
__device__ __forceinline__ void add4(float4& d, const float4& v) { d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w; }

template <typename F>
__device__ __forceinline__ void col_reduce4(long long rows, int c, int rows_per_cta, double* __restrict__ tot, F f) {
  __shared__ float4 sa[kRedThreads], sb[kRedThreads];
  const int c4 = c >> 2;
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  const long long r_end = min(rows, r_begin + rows_per_cta);
  for (int g0 = 0; g0 < c4; g0 += kRedThreads) {
    const int gw = min(kRedThreads, c4 - g0);
    const int lanes = kRedThreads / gw;
    const int gi = threadIdx.x % gw, lane = threadIdx.x / gw;
    const int ch = 4 * (g0 + gi);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a0 = z, b0 = z, a1 = z, b1 = z, a2 = z, b2 = z, a3 = z, b3 = z;
    if (lane < lanes) {
      long long r = r_begin + lane;
      for (; r + 3 * lanes < r_end; r += 4 * lanes) {
        f(r, ch, a0, b0);
        f(r + lanes, ch, a1, b1);
        f(r + 2 * lanes, ch, a2, b2);
        f(r + 3 * lanes, ch, a3, b3);
      }
      for (; r < r_end; r += lanes) f(r, ch, a0, b0);
    }
    add4(a0, a1); add4(a2, a3); add4(a0, a2);
    add4(b0, b1); add4(b2, b3); add4(b0, b2);
    sa[threadIdx.x] = a0;
    sb[threadIdx.x] = b0;
    __syncthreads();
    if (threadIdx.x < gw) {
      for (int l = 1; l < lanes; ++l) {
        add4(a0, sa[threadIdx.x + l * gw]);
        add4(b0, sb[threadIdx.x + l * gw]);
      }
      atomicAdd(tot + ch, (double)a0.x); atomicAdd(tot + ch + 1, (double)a0.y);
      atomicAdd(tot + ch + 2, (double)a0.z); atomicAdd(tot + ch + 3, (double)a0.w);
      atomicAdd(tot + c + ch, (double)b0.x); atomicAdd(tot + c + ch + 1, (double)b0.y);
      atomicAdd(tot + c + ch + 2, (double)b0.z); atomicAdd(tot + c + ch + 3, (double)b0.w);
    }
    __syncthreads();
  }
}

// rows per CTA for the float4 skeleton: at least two full 4-row rounds of every row lane
static inline int red_rows_per_cta4(long long rows, int c) {
  const int c4 = c / 4, gw = c4 < kRedThreads ? c4 : kRedThreads, lanes = kRedThreads / (gw > 0 ? gw : 1);
  long long per = (rows + kRedMaxCtas - 1) / kRedMaxCtas;
  if (per < 8ll * lanes) per = 8ll * lanes;
  return (int)per;
}

static inline bool vec4_ok(int c, const void* p0, const void* p1 = nullptr, const void* p2 = nullptr, const void* p3 = nullptr,
                           const void* p4 = nullptr, const void* p5 = nullptr) {
  return c % 4 == 0 && ((((size_t)p0 | (size_t)p1 | (size_t)p2 | (size_t)p3 | (size_t)p4 | (size_t)p5) & 15) == 0);
}

// ------------------------------------------------------------------------------------------------ BatchNorm (train)
__global__ void __launch_bounds__(kRedThreads) k_bn_stats(const float* __restrict__ x, long long rows, int c, int rpc,
                                                          double* __restrict__ tot) {
  col_reduce2(rows, c, rpc, tot, [&](long long r, int ch, float* a, float* b) {
    const float v = x[r * c + ch];
    *a += v;
    *b = fmaf(v, v, *b);
  });
}

__global__ void __launch_bounds__(kRedThreads) k_bn_stats_v4(const float* __restrict__ x, long long rows, int c, int rpc,
                                                             double* __restrict__ tot) {
  col_reduce4(rows, c, rpc, tot, [&](long long r, int ch, float4& a, float4& b) {
    const float4 v = *reinterpret_cast<const float4*>(x + r * c + ch);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
  });
}

// mean / biased variance -> scale, shift of the normalising affine; running statistics as nn.BatchNorm1d updates them
// (momentum, unbiased variance); the totals are cleared for the next user of the workspace
__global__ void k_bn_finalize(double* __restrict__ tot, long long rows, int c, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                              float* __restrict__ running_var, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                              float* __restrict__ scale, float* __restrict__ shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double n = (double)rows;
  const double mean = tot[ch] / n;
  double var = tot[c + ch] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  save_mean[ch] = (float)mean;
  save_invstd[ch] = invstd;
  scale[ch] = g * invstd;
  shift[ch] = b - (float)mean * g * invstd;
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
  tot[ch] = 0.0;
  tot[c + ch] = 0.0;
}

// totals of dz and dz * xhat, dz = gy * silu'(z) (z = x * scale + shift) when the forward applied SiLU
__global__ void __launch_bounds__(kRedThreads) k_bn_bwd_reduce(const float* __restrict__ x, const float* __restrict__ gy,
                                                               long long rows, int c, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int act, int rpc,
                                                               double* __restrict__ tot) {
  col_reduce2(rows, c, rpc, tot, [&](long long r, int ch, float* a, float* b) {
    const float xh = (x[r * c + ch] - mean[ch]) * invstd[ch];
    float dz = gy[r * c + ch];
    if (act & 1) dz *= dsilu_t(fmaf(xh, gamma ? gamma[ch] : 1.f, beta ? beta[ch] : 0.f));
    *a += dz;
    *b = fmaf(dz, xh, *b);
  });
}

__global__ void __launch_bounds__(kRedThreads) k_bn_bwd_reduce_v4(const float* __restrict__ x, const float* __restrict__ gy,
                                                                  long long rows, int c, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, int act, int rpc,
                                                                  double* __restrict__ tot) {
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  col_reduce4(rows, c, rpc, tot, [&](long long r, int ch, float4& a, float4& b) {
    const float4 xv = *reinterpret_cast<const float4*>(x + r * c + ch);
    float4 dz = *reinterpret_cast<const float4*>(gy + r * c + ch);
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean + ch)), is = __ldg(reinterpret_cast<const float4*>(invstd + ch));
    const float4 xh = make_float4((xv.x - m.x) * is.x, (xv.y - m.y) * is.y, (xv.z - m.z) * is.z, (xv.w - m.w) * is.w);
    if (act & 1) {
      const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + ch)) : one;
      const float4 bt = beta ? __ldg(reinterpret_cast<const float4*>(beta + ch)) : zero;
      dz.x *= dsilu_t(fmaf(xh.x, g.x, bt.x)); dz.y *= dsilu_t(fmaf(xh.y, g.y, bt.y));
      dz.z *= dsilu_t(fmaf(xh.z, g.z, bt.z)); dz.w *= dsilu_t(fmaf(xh.w, g.w, bt.w));
    }
    a.x += dz.x; a.y += dz.y; a.z += dz.z; a.w += dz.w;
    b.x = fmaf(dz.x, xh.x, b.x); b.y = fmaf(dz.y, xh.y, b.y); b.z = fmaf(dz.z, xh.z, b.z); b.w = fmaf(dz.w, xh.w, b.w);
  });
}

// the apply pass, four channels per thread
__global__ void __launch_bounds__(256) k_bn_bwd_apply_v4(const float* __restrict__ x, const float* __restrict__ gy, long long rows,
                                                         int c, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                                         const double* __restrict__ tot, float* __restrict__ gx,
                                                         float* __restrict__ ggamma, float* __restrict__ gbeta) {
  const int c4 = c >> 2;
  const long long total4 = rows * c4;
  const double inv_n = 1.0 / (double)rows;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
    const int ch = 4 * (int)(t % c4);
    const float4 xv = reinterpret_cast<const float4*>(x)[t];
    const float4 gv = reinterpret_cast<const float4*>(gy)[t];
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean + ch)), is = __ldg(reinterpret_cast<const float4*>(invstd + ch));
    const float4 g = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + ch)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 bt = beta ? __ldg(reinterpret_cast<const float4*>(beta + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
#define SPVD_BN_APPLY1(f, i)                                                                 \
  {                                                                                          \
    const float xh = (xv.f - m.f) * is.f;                                                    \
    float dz = gv.f;                                                                         \
    if (act & 1) dz *= dsilu_t(fmaf(xh, g.f, bt.f));                                         \
    const float m1 = (float)(tot[ch + i] * inv_n), m2 = (float)(tot[c + ch + i] * inv_n);    \
    o.f = g.f * is.f * (dz - m1 - xh * m2);                                                  \
  }
    SPVD_BN_APPLY1(x, 0) SPVD_BN_APPLY1(y, 1) SPVD_BN_APPLY1(z, 2) SPVD_BN_APPLY1(w, 3)
#undef SPVD_BN_APPLY1
    reinterpret_cast<float4*>(gx)[t] = o;
  }
  if (blockIdx.x == 0)
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      if (gbeta) gbeta[ch] = (float)tot[ch];
      if (ggamma) ggamma[ch] = (float)tot[c + ch];
    }
}

// gx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat));  CTA 0 also writes the parameter gradients
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ gy, long long rows,
                                                      int c, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                                      const double* __restrict__ tot, float* __restrict__ gx,
                                                      float* __restrict__ ggamma, float* __restrict__ gbeta) {
  const long long total = rows * c;
  const double inv_n = 1.0 / (double)rows;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(t % c);
    const float g = gamma ? gamma[ch] : 1.f;
    const float xh = (x[t] - mean[ch]) * invstd[ch];
    float dz = gy[t];
    if (act & 1) dz *= dsilu_t(fmaf(xh, g, beta ? beta[ch] : 0.f));
    const float m1 = (float)(tot[ch] * inv_n), m2 = (float)(tot[c + ch] * inv_n);
    gx[t] = g * invstd[ch] * (dz - m1 - xh * m2);
  }
  if (blockIdx.x == 0)
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      if (gbeta) gbeta[ch] = (float)tot[ch];
      if (ggamma) ggamma[ch] = (float)tot[c + ch];
    }
}

__global__ void k_clear_f64(double* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------ column sums
__global__ void __launch_bounds__(kRedThreads) k_colsum(const float* __restrict__ g, long long rows, int c, int rpc,
                                                        double* __restrict__ tot) {
  col_reduce2(rows, c, rpc, tot, [&](long long r, int ch, float* a, float* b) {
    *a += g[r * c + ch];
    (void)b;
  });
}
__global__ void __launch_bounds__(kRedThreads) k_colsum_v4(const float* __restrict__ g, long long rows, int c, int rpc,
                                                           double* __restrict__ tot) {
  col_reduce4(rows, c, rpc, tot, [&](long long r, int ch, float4& a, float4& b) {
    add4(a, *reinterpret_cast<const float4*>(g + r * c + ch));
    (void)b;
  });
}
__global__ void k_f64_to_f32(double* __restrict__ tot, int n, int clear_n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)tot[i];
  if (i < clear_n) tot[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------ FiLM backward
// gx = (1 + scale[b]) * gy;   gtt[b, c] += sum_rows gy * x,  gtt[b, C + c] += sum_rows gy   (rows of cloud b)
__global__ void __launch_bounds__(256) k_film_bwd(const float* __restrict__ x, const float* __restrict__ gy,
                                                  const float* __restrict__ tt, int ldt, const int* __restrict__ offsets, int c,
                                                  int rows_per_cta, float* __restrict__ gx, float* __restrict__ gtt, int ldg) {
  __shared__ float sa[256], sb[256];
  const int b = blockIdx.y;
  const int r0 = offsets[b] + blockIdx.x * rows_per_cta;
  const int r1 = min(offsets[b + 1], r0 + rows_per_cta);
  if (r0 >= r1) return;
  const float* trow = tt + (size_t)b * ldt;
  for (int c0 = 0; c0 < c; c0 += 256) {
    const int cw = min(256, c - c0), lanes = 256 / cw;
    const int ch = threadIdx.x % cw, lane = threadIdx.x / cw;
    float a = 0.f, s = 0.f;
    if (lane < lanes) {
      const float sc = 1.f + trow[c0 + ch];
      for (int r = r0 + lane; r < r1; r += lanes) {
        const size_t i = (size_t)r * c + c0 + ch;
        const float g = gy[i];
        gx[i] = sc * g;
        a = fmaf(g, x[i], a);
        s += g;
      }
    }
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < cw) {
      for (int l = 1; l < lanes; ++l) {
        a += sa[threadIdx.x + l * cw];
        s += sb[threadIdx.x + l * cw];
      }
      atomicAdd(gtt + (size_t)b * ldg + c0 + ch, a);
      atomicAdd(gtt + (size_t)b * ldg + c + c0 + ch, s);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// one warp per row; writes y and (optionally) the row statistics the backward needs
__global__ void k_layernorm_train(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  int rows, int c, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                                  float* __restrict__ rstd_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * c;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s += xr[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / c;
  float v = 0.f;
  for (int i = lane; i < c; i += 32) {
    const float d = xr[i] - mean;
    v = fmaf(d, d, v);
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / c + eps);
  for (int i = lane; i < c; i += 32) y[(size_t)row * c + i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// gx = rstd * (g' - mean(g') - xhat * mean(g' * xhat)),  g' = gy * gamma
__global__ void k_layernorm_bwd_x(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ gamma,
                                  const float* __restrict__ mean, const float* __restrict__ rstd, int rows, int c,
                                  float* __restrict__ gx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float mu = mean[row], rs = rstd[row];
  const float* xr = x + (size_t)row * c;
  const float* gr = gy + (size_t)row * c;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < c; i += 32) {
    const float gp = gr[i] * gamma[i];
    s1 += gp;
    s2 = fmaf(gp, (xr[i] - mu) * rs, s2);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  s1 /= c;
  s2 /= c;
  for (int i = lane; i < c; i += 32) {
    const float xh = (xr[i] - mu) * rs;
    gx[(size_t)row * c + i] = rs * (gr[i] * gamma[i] - s1 - xh * s2);
  }
}

__global__ void __launch_bounds__(kRedThreads) k_layernorm_bwd_params(const float* __restrict__ x, const float* __restrict__ gy,
                                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                      long long rows, int c, int rpc, double* __restrict__ tot) {
  col_reduce2(rows, c, rpc, tot, [&](long long r, int ch, float* a, float* b) {
    const float g = gy[r * c + ch];
    *a = fmaf(g, (x[r * c + ch] - mean[r]) * rstd[r], *a);    // d gamma
    *b += g;                                                  // d beta
  });
}

// ------------------------------------------------------------------------------------------------ attention (train)
// Forward with the log-sum-exp of every (row, head) saved; same layout and tiling as k_attention (runtime.cu):
// qkv [rows, 3C] = q | k | v with channel head*D + d, rows of a cloud contiguous, 4 lanes per query.
template <int D>
__global__ void __launch_bounds__(128) k_attention_train(const float* __restrict__ qkv, const int* __restrict__ offsets, int C,
                                                         int heads, float scale, float* __restrict__ out, float* __restrict__ lse) {
  constexpr int KB = D <= 32 ? 128 : 64;
  constexpr int QB = 32, S = 4;
  __shared__ float4 Ks[KB][D / 4];
  __shared__ float4 Vs[KB][D / 4];
  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = offsets[b], L = offsets[b + 1] - r0;
  if ((int)blockIdx.x * QB >= L) return;
  const int slice = threadIdx.x & (S - 1);
  const int qi = blockIdx.x * QB + (threadIdx.x >> 2);
  const bool valid = qi < L;
  float q[D], acc[D];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + (size_t)(r0 + (valid ? qi : 0)) * 3 * C + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 t = qp[i];
      q[4 * i] = t.x * scale; q[4 * i + 1] = t.y * scale; q[4 * i + 2] = t.z * scale; q[4 * i + 3] = t.w * scale;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < L; k0 += KB) {
    __syncthreads();
    const int nk = min(KB, L - k0);
    for (int e = threadIdx.x; e < nk * (D / 4); e += 128) {
      const int j = e / (D / 4), i = e % (D / 4);
      const float4* kp = reinterpret_cast<const float4*>(qkv + (size_t)(r0 + k0 + j) * 3 * C + C + h * D);
      Ks[j][i] = kp[i];
      Vs[j][i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(kp) + C)[i];
    }
    __syncthreads();
    for (int j = slice; j < nk; j += S) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        const float4 k4 = Ks[j][i];
        d = fmaf(q[4 * i], k4.x, d); d = fmaf(q[4 * i + 1], k4.y, d); d = fmaf(q[4 * i + 2], k4.z, d); d = fmaf(q[4 * i + 3], k4.w, d);
      }
      const float mn = fmaxf(m, d);
      const float alpha = __expf(m - mn), p = __expf(d - mn);
      m = mn;
      l = l * alpha + p;
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        const float4 v4 = Vs[j][i];
        acc[4 * i] = fmaf(p, v4.x, acc[4 * i] * alpha);
        acc[4 * i + 1] = fmaf(p, v4.y, acc[4 * i + 1] * alpha);
        acc[4 * i + 2] = fmaf(p, v4.z, acc[4 * i + 2] * alpha);
        acc[4 * i + 3] = fmaf(p, v4.w, acc[4 * i + 3] * alpha);
      }
    }
  }
#pragma unroll
  for (int o = 1; o < S; o <<= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m, o), l_o = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, m_o);
    const float wa = m == -INFINITY ? 0.f : __expf(m - mn), wb = m_o == -INFINITY ? 0.f : __expf(m_o - mn);
    l = l * wa + l_o * wb;
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = acc[i] * wa + __shfl_xor_sync(0xffffffffu, acc[i], o) * wb;
    m = mn;
  }
  if (valid) {
    const float inv = 1.f / l;
    float4* op = reinterpret_cast<float4*>(out + (size_t)(r0 + qi) * C + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      if ((i & (S - 1)) != slice) continue;
      op[i] = make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv);
    }
    if (slice == 0 && lse) lse[(size_t)(r0 + qi) * heads + h] = m + __logf(l);
  }
}

// dQ and delta = dO . O per (row, head): 4 lanes per query, keys / values staged in shared memory
template <int D>
__global__ void __launch_bounds__(128) k_attention_bwd_q(const float* __restrict__ qkv, const float* __restrict__ out,
                                                         const float* __restrict__ gout, const float* __restrict__ lse,
                                                         const int* __restrict__ offsets, int C, int heads, float scale,
                                                         float* __restrict__ gqkv, float* __restrict__ delta) {
  constexpr int KB = D <= 32 ? 128 : 64;
  constexpr int QB = 32, S = 4;
  __shared__ float4 Ks[KB][D / 4];
  __shared__ float4 Vs[KB][D / 4];
  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = offsets[b], L = offsets[b + 1] - r0;
  if ((int)blockIdx.x * QB >= L) return;
  const int slice = threadIdx.x & (S - 1);
  const int qi = blockIdx.x * QB + (threadIdx.x >> 2);
  const bool valid = qi < L;
  const size_t row = (size_t)(r0 + (valid ? qi : 0));
  float q[D], go[D], dq[D];
  float dl = 0.f;
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + row * 3 * C + h * D);
    const float4* gp = reinterpret_cast<const float4*>(gout + row * C + h * D);
    const float4* op = reinterpret_cast<const float4*>(out + row * C + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 t = qp[i], g = gp[i], o = op[i];
      q[4 * i] = t.x * scale; q[4 * i + 1] = t.y * scale; q[4 * i + 2] = t.z * scale; q[4 * i + 3] = t.w * scale;
      go[4 * i] = g.x; go[4 * i + 1] = g.y; go[4 * i + 2] = g.z; go[4 * i + 3] = g.w;
      dl += g.x * o.x + g.y * o.y + g.z * o.z + g.w * o.w;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) dq[i] = 0.f;
  }
  const float ls = valid ? lse[row * heads + h] : 0.f;
  for (int k0 = 0; k0 < L; k0 += KB) {
    __syncthreads();
    const int nk = min(KB, L - k0);
    for (int e = threadIdx.x; e < nk * (D / 4); e += 128) {
      const int j = e / (D / 4), i = e % (D / 4);
      const float4* kp = reinterpret_cast<const float4*>(qkv + (size_t)(r0 + k0 + j) * 3 * C + C + h * D);
      Ks[j][i] = kp[i];
      Vs[j][i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(kp) + C)[i];
    }
    __syncthreads();
    for (int j = slice; j < nk; j += S) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        const float4 k4 = Ks[j][i], v4 = Vs[j][i];
        s = fmaf(q[4 * i], k4.x, s); s = fmaf(q[4 * i + 1], k4.y, s); s = fmaf(q[4 * i + 2], k4.z, s); s = fmaf(q[4 * i + 3], k4.w, s);
        dp = fmaf(go[4 * i], v4.x, dp); dp = fmaf(go[4 * i + 1], v4.y, dp); dp = fmaf(go[4 * i + 2], v4.z, dp); dp = fmaf(go[4 * i + 3], v4.w, dp);
      }
      const float ds = __expf(s - ls) * (dp - dl);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        const float4 k4 = Ks[j][i];
        dq[4 * i] = fmaf(ds, k4.x, dq[4 * i]); dq[4 * i + 1] = fmaf(ds, k4.y, dq[4 * i + 1]);
        dq[4 * i + 2] = fmaf(ds, k4.z, dq[4 * i + 2]); dq[4 * i + 3] = fmaf(ds, k4.w, dq[4 * i + 3]);
      }
    }
  }
#pragma unroll
  for (int o = 1; o < S; o <<= 1)
#pragma unroll
    for (int i = 0; i < D; ++i) dq[i] += __shfl_xor_sync(0xffffffffu, dq[i], o);
  if (valid) {
    float4* gq = reinterpret_cast<float4*>(gqkv + row * 3 * C + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      if ((i & (S - 1)) != slice) continue;
      gq[i] = make_float4(dq[4 * i] * scale, dq[4 * i + 1] * scale, dq[4 * i + 2] * scale, dq[4 * i + 3] * scale);
    }
    if (slice == 0) delta[row * heads + h] = dl;
  }
}

// dK and dV: 4 lanes per key, queries (scaled q, dO, lse, delta) staged in shared memory
template <int D>
__global__ void __launch_bounds__(128) k_attention_bwd_kv(const float* __restrict__ qkv, const float* __restrict__ gout,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          const int* __restrict__ offsets, int C, int heads, float scale,
                                                          float* __restrict__ gqkv) {
  constexpr int QT = D <= 32 ? 128 : 64;   // queries per shared-memory tile
  constexpr int KBLK = 32, S = 4;
  __shared__ float4 Qs[QT][D / 4];
  __shared__ float4 Gs[QT][D / 4];
  __shared__ float Ls[QT], Ds[QT];
  const int b = blockIdx.z, h = blockIdx.y;
  const int r0 = offsets[b], L = offsets[b + 1] - r0;
  if ((int)blockIdx.x * KBLK >= L) return;
  const int slice = threadIdx.x & (S - 1);
  const int kj = blockIdx.x * KBLK + (threadIdx.x >> 2);
  const bool valid = kj < L;
  const size_t row = (size_t)(r0 + (valid ? kj : 0));
  float k[D], v[D], dk[D], dv[D];
  {
    const float4* kp = reinterpret_cast<const float4*>(qkv + row * 3 * C + C + h * D);
    const float4* vp = reinterpret_cast<const float4*>(qkv + row * 3 * C + 2 * C + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 a = kp[i], c4 = vp[i];
      k[4 * i] = a.x; k[4 * i + 1] = a.y; k[4 * i + 2] = a.z; k[4 * i + 3] = a.w;
      v[4 * i] = c4.x; v[4 * i + 1] = c4.y; v[4 * i + 2] = c4.z; v[4 * i + 3] = c4.w;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) dk[i] = dv[i] = 0.f;
  }
  for (int q0 = 0; q0 < L; q0 += QT) {
    __syncthreads();
    const int nq = min(QT, L - q0);
    for (int e = threadIdx.x; e < nq * (D / 4); e += 128) {
      const int j = e / (D / 4), i = e % (D / 4);
      float4 t = reinterpret_cast<const float4*>(qkv + (size_t)(r0 + q0 + j) * 3 * C + h * D)[i];
      t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
      Qs[j][i] = t;
      Gs[j][i] = reinterpret_cast<const float4*>(gout + (size_t)(r0 + q0 + j) * C + h * D)[i];
    }
    for (int j = threadIdx.x; j < nq; j += 128) {
      Ls[j] = lse[(size_t)(r0 + q0 + j) * heads + h];
      Ds[j] = delta[(size_t)(r0 + q0 + j) * heads + h];
    }
    __syncthreads();
    for (int j = slice; j < nq; j += S) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        const float4 q4 = Qs[j][i], g4 = Gs[j][i];
        s = fmaf(q4.x, k[4 * i], s); s = fmaf(q4.y, k[4 * i + 1], s); s = fmaf(q4.z, k[4 * i + 2], s); s = fmaf(q4.w, k[4 * i + 3], s);
        dp = fmaf(g4.x, v[4 * i], dp); dp = fmaf(g4.y, v[4 * i + 1], dp); dp = fmaf(g4.z, v[4 * i + 2], dp); dp = fmaf(g4.w, v[4 * i + 3], dp);
      }
      const float p = __expf(s - Ls[j]);
      const float ds = p * (dp - Ds[j]);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        const float4 q4 = Qs[j][i], g4 = Gs[j][i];
        dv[4 * i] = fmaf(p, g4.x, dv[4 * i]); dv[4 * i + 1] = fmaf(p, g4.y, dv[4 * i + 1]);
        dv[4 * i + 2] = fmaf(p, g4.z, dv[4 * i + 2]); dv[4 * i + 3] = fmaf(p, g4.w, dv[4 * i + 3]);
        dk[4 * i] = fmaf(ds, q4.x, dk[4 * i]); dk[4 * i + 1] = fmaf(ds, q4.y, dk[4 * i + 1]);     // q4 carries the scale
        dk[4 * i + 2] = fmaf(ds, q4.z, dk[4 * i + 2]); dk[4 * i + 3] = fmaf(ds, q4.w, dk[4 * i + 3]);
      }
    }
  }
#pragma unroll
  for (int o = 1; o < S; o <<= 1)
#pragma unroll
    for (int i = 0; i < D; ++i) {
      dk[i] += __shfl_xor_sync(0xffffffffu, dk[i], o);
      dv[i] += __shfl_xor_sync(0xffffffffu, dv[i], o);
    }
  if (valid) {
    float4* gk = reinterpret_cast<float4*>(gqkv + row * 3 * C + C + h * D);
    float4* gv = reinterpret_cast<float4*>(gqkv + row * 3 * C + 2 * C + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      if ((i & (S - 1)) != slice) continue;
      gk[i] = make_float4(dk[4 * i], dk[4 * i + 1], dk[4 * i + 2], dk[4 * i + 3]);
      gv[i] = make_float4(dv[4 * i], dv[4 * i + 1], dv[4 * i + 2], dv[4 * i + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ small elementwise
__global__ void k_split(const float* __restrict__ g, long long rows, int ca, int cb, float* __restrict__ ga, float* __restrict__ gb) {
  const long long total = rows * (ca + cb);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / (ca + cb);
    const int q = (int)(t % (ca + cb));
    if (q < ca) ga[r * ca + q] = g[t];
    else gb[r * cb + (q - ca)] = g[t];
  }
}
__global__ void k_cat(const float* __restrict__ a, const float* __restrict__ b, long long rows, int ca, int cb, float* __restrict__ y) {
  const long long total = rows * (ca + cb);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / (ca + cb);
    const int q = (int)(t % (ca + cb));
    y[t] = q < ca ? a[r * ca + q] : b[r * cb + (q - ca)];
  }
}
__global__ void k_silu_fwd(const float* __restrict__ x, long long n, float* __restrict__ y) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) y[t] = silu_t(x[t]);
}
__global__ void k_silu_bwd(const float* __restrict__ x, const float* __restrict__ gy, long long n, float* __restrict__ gx) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    gx[t] = gy[t] * dsilu_t(x[t]);
}
// y = a + b (gradient accumulation / residual sums on the training path)
__global__ void k_add(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ y) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) y[t] = a[t] + b[t];
}
// out[b, :] = table[idx[b], :] (+ base[b, :]);  backward: gtable[idx[b], :] += g[b, :]
__global__ void k_embedding_fwd(const float* __restrict__ table, const long long* __restrict__ idx, const float* __restrict__ base,
                                int e, float* __restrict__ out) {
  const int b = blockIdx.x;
  const float* row = table + (size_t)idx[b] * e;
  for (int i = threadIdx.x; i < e; i += blockDim.x) out[(size_t)b * e + i] = row[i] + (base ? base[(size_t)b * e + i] : 0.f);
}
__global__ void k_embedding_bwd(const float* __restrict__ g, const long long* __restrict__ idx, int e, float* __restrict__ gtable) {
  const int b = blockIdx.x;
  float* row = gtable + (size_t)idx[b] * e;
  for (int i = threadIdx.x; i < e; i += blockDim.x) atomicAdd(row + i, g[(size_t)b * e + i]);
}
// models/utils.py:15-19: exponent = -ln(1e4) * linspace(0, 1, d/2); emb = [sin(t * e^exponent), cos(t * e^exponent)]
__global__ void k_timestep_embedding(const long long* __restrict__ t, int d, float* __restrict__ out) {
  const int b = blockIdx.x, half = d / 2;
  const float tv = (float)t[b];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float lin = half > 1 ? (float)i / (float)(half - 1) : 0.f;
    const float e = tv * expf(-logf(10000.f) * lin);
    out[(size_t)b * d + i] = sinf(e);
    out[(size_t)b * d + half + i] = cosf(e);
  }
}
// loss = mean((p - t)^2), gp = 2 (p - t) / n * gscale
__global__ void __launch_bounds__(256) k_mse(const float* __restrict__ p, const float* __restrict__ t, long long n, float gscale,
                                             float* __restrict__ gp, double* __restrict__ tot) {
  __shared__ float sm[8];
  float a = 0.f;
  const float k = 2.f * gscale / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = p[i] - t[i];
    a = fmaf(d, d, a);
    if (gp) gp[i] = k * d;
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) a += sm[w];
    atomicAdd(tot, (double)a);
  }
}
__global__ void k_mse_finalize(double* __restrict__ tot, long long n, float* __restrict__ loss) {
  loss[0] = (float)(tot[0] / (double)n);
  tot[0] = 0.0;
}

static inline int grid_for(long long n, int per = 256, int cap = 148 * 16) {
  long long g = (n + per - 1) / per;
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_bn_silu_fwd(const float*, const float*, const float*, long long, int, int, float*, spvd_stream_t);

extern "C" int spvd_bn_train_fwd(const float* x, long long rows, int c, const float* gamma, const float* beta, float eps,
                                 float momentum, float* running_mean, float* running_var, float* save_mean,
                                 float* save_invstd, float* scale, float* shift, int act, float* y, double* workspace,
                                 spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && save_mean && save_invstd && scale && shift && y && workspace && rows > 0 && c > 0,
                 "spvd_bn_train_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int rpc = red_rows_per_cta(rows);
  if (vec4_ok(c, x)) {
    const int rp4 = red_rows_per_cta4(rows, c);
    k_bn_stats_v4<<<cdiv(rows, rp4), kRedThreads, 0, s>>>(x, rows, c, rp4, workspace);
  }
  else k_bn_stats<<<cdiv(rows, rpc), kRedThreads, 0, s>>>(x, rows, c, rpc, workspace);
  k_bn_finalize<<<cdiv(c, 128), 128, 0, s>>>(workspace, rows, c, gamma, beta, eps, momentum, running_mean, running_var, save_mean,
                                             save_invstd, scale, shift);
  SPVD_LAUNCH_CHECK();
  return spvd_bn_silu_fwd(x, scale, shift, rows, c, act & 3, y, stream);   // bit 1: round y to TF32 (operand of a tensor-core conv)
}

extern "C" int spvd_bn_train_bwd(const float* x, const float* gy, long long rows, int c, const float* gamma, const float* beta,
                                 const float* save_mean, const float* save_invstd, int act, float* gx, float* ggamma,
                                 float* gbeta, double* workspace, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && gy && save_mean && save_invstd && gx && workspace && rows > 0 && c > 0, "spvd_bn_train_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int rpc = red_rows_per_cta(rows);
  if (vec4_ok(c, x, gy, save_mean, save_invstd, gamma, beta) && ((size_t)gx & 15) == 0) {
    const int rp4 = red_rows_per_cta4(rows, c);
    k_bn_bwd_reduce_v4<<<cdiv(rows, rp4), kRedThreads, 0, s>>>(x, gy, rows, c, save_mean, save_invstd, gamma, beta, act, rp4, workspace);
    k_bn_bwd_apply_v4<<<grid_for(rows * (c / 4)), 256, 0, s>>>(x, gy, rows, c, save_mean, save_invstd, gamma, beta, act, workspace, gx,
                                                             ggamma, gbeta);
  } else {
    k_bn_bwd_reduce<<<cdiv(rows, rpc), kRedThreads, 0, s>>>(x, gy, rows, c, save_mean, save_invstd, gamma, beta, act, rpc, workspace);
    k_bn_bwd_apply<<<grid_for(rows * c), 256, 0, s>>>(x, gy, rows, c, save_mean, save_invstd, gamma, beta, act, workspace, gx, ggamma,
                                                      gbeta);
  }
  k_clear_f64<<<cdiv(2 * c, 256), 256, 0, s>>>(workspace, 2 * c);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_colsum(const float* g, long long rows, int c, float* out, double* workspace, spvd_stream_t stream) {
  SPVD_CHECK_ARG(g && out && workspace && rows >= 0 && c > 0, "spvd_colsum: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int rpc = red_rows_per_cta(rows);
  if (rows > 0) {
    const int rp4 = red_rows_per_cta4(rows, c);
    if (vec4_ok(c, g)) k_colsum_v4<<<cdiv(rows, rp4), kRedThreads, 0, s>>>(g, rows, c, rp4, workspace);
    else k_colsum<<<cdiv(rows, rpc), kRedThreads, 0, s>>>(g, rows, c, rpc, workspace);
  }
  k_f64_to_f32<<<cdiv(2 * c, 256), 256, 0, s>>>(workspace, c, 2 * c, out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_film_bwd(const float* x, const float* gy, const float* tt, int ldt, const int32_t* batch_offsets, int n_batch,
                             int max_rows_per_cloud, int c, float* gx, float* gtt, int ldg, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && gy && tt && batch_offsets && gx && gtt && n_batch > 0 && c > 0 && ldt >= 2 * c && ldg >= 2 * c,
                 "spvd_film_bwd: bad arguments");
  if (max_rows_per_cloud <= 0) return SPVD_OK;
  const int rows_per_cta = 64;
  dim3 grid(cdiv(max_rows_per_cloud, rows_per_cta), n_batch);
  k_film_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(x, gy, tt, ldt, batch_offsets, c, rows_per_cta, gx, gtt, ldg);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_layernorm_train_fwd(const float* x, const float* gamma, const float* beta, int rows, int c, float eps,
                                        float* y, float* mean, float* rstd, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && gamma && beta && y && rows >= 0 && c > 0, "spvd_layernorm_train_fwd: bad arguments");
  if (rows == 0) return SPVD_OK;
  k_layernorm_train<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, rows, c, eps, y, mean, rstd);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_layernorm_bwd(const float* x, const float* gy, const float* gamma, const float* mean, const float* rstd,
                                  int rows, int c, float* gx, float* ggamma, float* gbeta, double* workspace,
                                  spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && gy && gamma && mean && rstd && gx && workspace && rows >= 0 && c > 0, "spvd_layernorm_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (rows > 0) {
    k_layernorm_bwd_x<<<cdiv(rows, 8), 256, 0, s>>>(x, gy, gamma, mean, rstd, rows, c, gx);
    const int rpc = red_rows_per_cta(rows);
    if (ggamma || gbeta) k_layernorm_bwd_params<<<cdiv(rows, rpc), kRedThreads, 0, s>>>(x, gy, mean, rstd, rows, c, rpc, workspace);
  }
  if (ggamma) k_f64_to_f32<<<cdiv(c, 256), 256, 0, s>>>(workspace, c, c, ggamma);
  if (gbeta) k_f64_to_f32<<<cdiv(c, 256), 256, 0, s>>>(workspace + c, c, c, gbeta);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

#define SPVD_ATTN_DISPATCH(D_, CALL)                              \
  switch (D_) {                                                   \
    case 8: { constexpr int D = 8; CALL; } break;                 \
    case 16: { constexpr int D = 16; CALL; } break;               \
    case 24: { constexpr int D = 24; CALL; } break;               \
    case 32: { constexpr int D = 32; CALL; } break;               \
    case 48: { constexpr int D = 48; CALL; } break;               \
    case 64: { constexpr int D = 64; CALL; } break;               \
    default:                                                      \
      set_error("attention: head dim %d is not instantiated (8,16,24,32,48,64)", D_); \
      return SPVD_ERR_INVALID;                                    \
  }

extern "C" int spvd_attention_train_fwd(const float* qkv, const int32_t* batch_offsets, int n_batch, int max_len, int c, int heads,
                                        float* out, float* lse, spvd_stream_t stream) {
  SPVD_CHECK_ARG(qkv && batch_offsets && out && lse && n_batch > 0 && heads > 0 && c % heads == 0, "spvd_attention_train_fwd: bad arguments");
  if (max_len <= 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int d = c / heads;
  const float scale = 1.0f / sqrtf((float)d);
  dim3 grid(cdiv(max_len, 32), heads, n_batch);
  SPVD_ATTN_DISPATCH(d, (k_attention_train<D><<<grid, 128, 0, s>>>(qkv, batch_offsets, c, heads, scale, out, lse)));
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_attention_bwd(const float* qkv, const float* out, const float* gout, const float* lse,
                                  const int32_t* batch_offsets, int n_batch, int max_len, int c, int heads, float* gqkv,
                                  float* delta, spvd_stream_t stream) {
  SPVD_CHECK_ARG(qkv && out && gout && lse && batch_offsets && gqkv && delta && n_batch > 0 && heads > 0 && c % heads == 0,
                 "spvd_attention_bwd: bad arguments");
  if (max_len <= 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int d = c / heads;
  const float scale = 1.0f / sqrtf((float)d);
  dim3 grid(cdiv(max_len, 32), heads, n_batch);
  SPVD_ATTN_DISPATCH(d, (k_attention_bwd_q<D><<<grid, 128, 0, s>>>(qkv, out, gout, lse, batch_offsets, c, heads, scale, gqkv, delta)));
  SPVD_ATTN_DISPATCH(d, (k_attention_bwd_kv<D><<<grid, 128, 0, s>>>(qkv, gout, lse, delta, batch_offsets, c, heads, scale, gqkv)));
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_split_fwd(const float* g, long long rows, int ca, int cb, float* ga, float* gb, spvd_stream_t stream) {
  SPVD_CHECK_ARG(g && ga && gb && rows >= 0 && ca > 0 && cb > 0, "spvd_split_fwd: bad arguments");
  if (rows == 0) return SPVD_OK;
  k_split<<<grid_for(rows * (ca + cb)), 256, 0, (cudaStream_t)stream>>>(g, rows, ca, cb, ga, gb);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_cat_any_fwd(const float* a, const float* b, long long rows, int ca, int cb, float* y, spvd_stream_t stream) {
  SPVD_CHECK_ARG(a && b && y && rows >= 0 && ca > 0 && cb > 0, "spvd_cat_any_fwd: bad arguments");
  if (rows == 0) return SPVD_OK;
  k_cat<<<grid_for(rows * (ca + cb)), 256, 0, (cudaStream_t)stream>>>(a, b, rows, ca, cb, y);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_silu_fwd(const float* x, long long n, float* y, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && y && n >= 0, "spvd_silu_fwd: bad arguments");
  if (n == 0) return SPVD_OK;
  k_silu_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, n, y);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_silu_bwd(const float* x, const float* gy, long long n, float* gx, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && gy && gx && n >= 0, "spvd_silu_bwd: bad arguments");
  if (n == 0) return SPVD_OK;
  k_silu_bwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, gy, n, gx);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_add_fwd(const float* a, const float* b, long long n, float* y, spvd_stream_t stream) {
  SPVD_CHECK_ARG(a && b && y && n >= 0, "spvd_add_fwd: bad arguments");
  if (n == 0) return SPVD_OK;
  k_add<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, y);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_embedding_fwd(const float* table, const int64_t* idx, const float* base, int n_batch, int e, float* out,
                                  spvd_stream_t stream) {
  SPVD_CHECK_ARG(table && idx && out && n_batch > 0 && e > 0, "spvd_embedding_fwd: bad arguments");
  k_embedding_fwd<<<n_batch, 128, 0, (cudaStream_t)stream>>>(table, (const long long*)idx, base, e, out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_embedding_bwd(const float* g, const int64_t* idx, int n_batch, int e, float* gtable, spvd_stream_t stream) {
  SPVD_CHECK_ARG(g && idx && gtable && n_batch > 0 && e > 0, "spvd_embedding_bwd: bad arguments");
  k_embedding_bwd<<<n_batch, 128, 0, (cudaStream_t)stream>>>(g, (const long long*)idx, e, gtable);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_timestep_embedding(const int64_t* t, int n_batch, int d, float* out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(t && out && n_batch > 0 && d > 0 && d % 2 == 0, "spvd_timestep_embedding: bad arguments");
  k_timestep_embedding<<<n_batch, 64, 0, (cudaStream_t)stream>>>((const long long*)t, d, out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_mse_loss(const float* pred, const float* target, long long n, float grad_scale, float* loss, float* gpred,
                             double* workspace, spvd_stream_t stream) {
  SPVD_CHECK_ARG(pred && target && loss && workspace && n > 0, "spvd_mse_loss: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  k_mse<<<grid_for(n, 1024, 296), 256, 0, s>>>(pred, target, n, grad_scale, gpred, workspace);
  k_mse_finalize<<<1, 1, 0, s>>>(workspace, n, loss);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

// ------------------------------------------------------------------------------------------------
// Optimizer tail of a training step over the flat buffers of parallel.GradientAllReduce / optim.FlatAdam: one pass for the
// gradient norm, one pass for clip + (data-parallel average) + Adam.  torch.optim.Adam semantics (no amsgrad, L2 weight
// decay added to the gradient): m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sumsq(const float4* __restrict__ x, long long n4, double* __restrict__ acc) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;       // <= a few hundred terms per thread in fp32, then fp64
  }
  double d = (double)s;
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    atomicAdd(acc, t);
  }
}

__global__ void __launch_bounds__(256) k_adam(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                              float4* __restrict__ v, long long n4, float lr_c, float b1, float b2, float eps,
                                              float wd, float inv_bc2_sqrt, const double* __restrict__ sumsq, float max_norm,
                                              float pending, float* __restrict__ norm_out) {
  float scale = pending;
  if (sumsq) {
    const float norm = (float)sqrt(*sumsq) * pending;          // the norm of the averaged gradient
    if (max_norm > 0.f) scale = fminf(max_norm / (norm + 1e-6f), 1.f) * pending;     // torch.nn.utils.clip_grad_norm_
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  }
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 P = p[i], G = g[i], M = m[i], V = v[i];
#define SPVD_ADAM1(f)                                      \
  {                                                        \
    float gg = G.f * scale;                                \
    if (wd != 0.f) gg += wd * P.f;                         \
    M.f = b1 * M.f + c1 * gg;                              \
    V.f = b2 * V.f + c2 * gg * gg;                         \
    P.f -= lr_c * (M.f / (sqrtf(V.f) * inv_bc2_sqrt + eps)); \
  }
    SPVD_ADAM1(x) SPVD_ADAM1(y) SPVD_ADAM1(z) SPVD_ADAM1(w)
#undef SPVD_ADAM1
    p[i] = P;
    m[i] = M;
    v[i] = V;
  }
}

extern "C" int spvd_sumsq(const float* x, long long n, double* acc, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && acc && n > 0 && n % 4 == 0 && ((size_t)x & 15) == 0, "spvd_sumsq: bad arguments (n a multiple of 4, 16-byte aligned)");
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), s));
  k_sumsq<<<grid_for(n / 4, 256 * 8, 148 * 8), 256, 0, s>>>(reinterpret_cast<const float4*>(x), n / 4, acc);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                              float eps, float weight_decay, int step, const double* sumsq, float max_norm, float pending_scale,
                              float* norm_out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(p && g && m && v && n > 0 && n % 4 == 0 && step >= 1, "spvd_adam_step: bad arguments (n a multiple of 4, step >= 1)");
  SPVD_CHECK_ARG((((size_t)p | (size_t)g | (size_t)m | (size_t)v) & 15) == 0, "spvd_adam_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  k_adam<<<grid_for(n / 4, 256 * 4, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
      n / 4, (float)((double)lr / bc1), beta1, beta2, eps, weight_decay, (float)(1.0 / sqrt(bc2)), sumsq, max_norm, pending_scale,
      norm_out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

// y = x rounded to TF32 (cvt.rna) and, in the same pass, the column sums of x (bias gradient)
__global__ void __launch_bounds__(kRedThreads) k_round_colsum(const float* __restrict__ x, long long rows, int c, int rpc,
                                                              float* __restrict__ y, double* __restrict__ tot) {
  col_reduce2(rows, c, rpc, tot, [&](long long r, int ch, float* a, float* b) {
    const float v = x[r * c + ch];
    unsigned q;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(v));
    y[r * c + ch] = __uint_as_float(q);
    *a += v;
    (void)b;
  });
}

__global__ void __launch_bounds__(kRedThreads) k_round_colsum_v4(const float* __restrict__ x, long long rows, int c, int rpc,
                                                                 float* __restrict__ y, double* __restrict__ tot) {
  col_reduce4(rows, c, rpc, tot, [&](long long r, int ch, float4& a, float4& b) {
    const float4 v = *reinterpret_cast<const float4*>(x + r * c + ch);
    uint4 q;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q.x) : "f"(v.x));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q.y) : "f"(v.y));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q.z) : "f"(v.z));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q.w) : "f"(v.w));
    *reinterpret_cast<uint4*>(y + r * c + ch) = q;
    add4(a, v);
    (void)b;
  });
}

// tensor-core tile image (K-major SWIZZLE_128B, TF32-rounded: the layout of spvd_conv_pack_weights) of the TRANSPOSED
// weights W'[k'][co][ci] = W[k][ci][co], k = K-1-k' when flip -- the data-gradient weights, without materialising W'
__global__ void k_pack_weights_transposed(const float* __restrict__ w, int kvol, int cin, int cout, int flip,
                                          float* __restrict__ wp) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)kvol * cin * cout) return;
  const int n = (int)(t % cin);                 // output channel of the data gradient = input channel of the forward
  const long long q = t / cin;
  const int ci = (int)(q % cout);               // its input channel = forward output channel
  const int kp = (int)(q / cout);
  const int k = flip ? kvol - 1 - kp : kp;
  const int c = ci / 32, kk = ci % 32;
  const size_t dst = ((size_t)(kp * (cout / 32) + c) * cin + n) * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w[((size_t)k * cin + n) * cout + ci]));
  wp[dst] = __uint_as_float(r);
}

// Every per-step derivative of one weight tensor in ONE pass (training: the optimizer changed the weights): the forward
// tile image (spvd_conv_pack_weights layout), the transposed image for the data gradient (k_pack_weights_transposed
// layout, offsets flipped when `flip`) and, for an nn.Linear weight [cout, cin], its conv layout [1, cin, cout].
__global__ void k_pack_weights_train(const float* __restrict__ w, int kvol, int cin, int cout, int linear, int flip,
                                     float* __restrict__ w3, float* __restrict__ wp, float* __restrict__ wtp) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)kvol * cin * cout) return;
  const int co = (int)(t % cout);
  const long long q = t / cout;
  const int ci = (int)(q % cin);
  const int k = (int)(q / cin);
  const float v = linear ? w[(size_t)co * cin + ci] : w[t];
  if (w3) w3[t] = v;
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  const float vr = __uint_as_float(r);
  if (wp) {
    const int c = ci / 32, kk = ci % 32;
    wp[((size_t)(k * (cin / 32) + c) * cout + co) * 32 + ((((kk >> 2) ^ (co & 7)) << 2) | (kk & 3))] = vr;
  }
  if (wtp) {
    const int kp = flip ? kvol - 1 - k : k;
    const int c = co / 32, kk = co % 32;
    wtp[((size_t)(kp * (cout / 32) + c) * cin + ci) * 32 + ((((kk >> 2) ^ (ci & 7)) << 2) | (kk & 3))] = vr;
  }
}

extern "C" int spvd_conv_umma_supported(int, int, int);
extern "C" int spvd_conv_pack_weights_train(const float* w, int kvol, int cin, int cout, int w_is_linear, int flip_t, float* w3_out,
                                            float* wp_out, float* wtp_out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(w && kvol >= 1 && cin >= 1 && cout >= 1 && (w3_out || wp_out || wtp_out), "spvd_conv_pack_weights_train: bad arguments");
  SPVD_CHECK_ARG(!w_is_linear || kvol == 1, "spvd_conv_pack_weights_train: an nn.Linear weight has kernel size 1");
  SPVD_CHECK_ARG(!wp_out || spvd_conv_umma_supported(kvol, cin, cout), "spvd_conv_pack_weights_train: unsupported shape %dx%dx%d", kvol, cin, cout);
  SPVD_CHECK_ARG(!wtp_out || spvd_conv_umma_supported(kvol, cout, cin), "spvd_conv_pack_weights_train: unsupported transposed shape %dx%dx%d",
                 kvol, cout, cin);
  const long long total = (long long)kvol * cin * cout;
  k_pack_weights_train<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, kvol, cin, cout, w_is_linear, flip_t, w3_out, wp_out, wtp_out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

// ------------------------------------------------------------------------------------------------
// Whole backward of one sparse convolution / nn.Linear from ONE host call: TF32 rounding of the output gradient, weight
// transpose + tile image, data gradient (the forward kernels on the transposed weights and the gradient map), weight
// gradient (tcgen05 over the pair lists, fp32 SIMT fallback) and the bias gradient -- nine launches that the autograd
// functions used to issue one Python -> C call at a time (the training step was host-bound on exactly that).
// ------------------------------------------------------------------------------------------------
extern "C" {
int spvd_round_tf32(const float*, float*, long long, spvd_stream_t);
int spvd_conv_transpose_weights(const float*, int, int, int, int, float*, spvd_stream_t);
int spvd_conv_pack_weights(const float*, int, int, int, float*, spvd_stream_t);
int spvd_conv_umma_supported(int, int, int);
int spvd_conv_fwd_umma(const float*, const float*, const int32_t*, int, const uint32_t*, int, int, int, int, const float*,
                       const float*, float*, int, int, int, int, spvd_stream_t);
int spvd_conv_fwd_simt(const float*, const float*, const int32_t*, int, const uint32_t*, int, int, int, int, const float*,
                       const float*, float*, spvd_stream_t);
int spvd_conv_wgrad_umma(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int,
                         float*, spvd_stream_t);
int spvd_conv_wgrad(const float*, const float*, const int32_t*, int, int, int, int, int, float*, spvd_stream_t);
int spvd_conv_fwd_pairs(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int, int,
                        const float*, const float*, float*, int, spvd_stream_t);
}

static inline size_t al256(size_t b) { return (b + 255) / 256 * 256; }

extern "C" size_t spvd_conv_bwd_workspace_bytes(int n_in, int n_out, int kvol, int cin, int cout) {
  return al256((size_t)n_out * cout * 4) + al256((size_t)n_in * cin * 4) + 3 * al256((size_t)kvol * cin * cout * 4) + 1024;
}

extern "C" int spvd_conv_bwd(const spvd_conv_bwd_args_t* args, spvd_stream_t stream) {
  SPVD_CHECK_ARG(args, "spvd_conv_bwd: NULL arguments");
  const spvd_conv_bwd_args_t& a = *args;
  SPVD_CHECK_ARG(a.g && a.w && a.workspace && a.red_workspace && a.kvol >= 1 && a.cin >= 1 && a.cout >= 1 && a.n_in >= 0 && a.n_out >= 0,
                 "spvd_conv_bwd: bad arguments");
  SPVD_CHECK_ARG(a.workspace_bytes >= spvd_conv_bwd_workspace_bytes(a.n_in, a.n_out, a.kvol, a.cin, a.cout),
                 "spvd_conv_bwd: workspace too small");
  SPVD_CHECK_ARG(!a.gw || a.x, "spvd_conv_bwd: the weight gradient needs the forward input");
  char* ws = (char*)(((size_t)a.workspace + 255) / 256 * 256);
  float* gr = (float*)ws;                                   ws += al256((size_t)a.n_out * a.cout * 4);
  float* xr = (float*)ws;                                   ws += al256((size_t)a.n_in * a.cin * 4);
  float* wt = (float*)ws;                                   ws += al256((size_t)a.kvol * a.cin * a.cout * 4);
  float* wtp = (float*)ws;                                  ws += al256((size_t)a.kvol * a.cin * a.cout * 4);
  float* gw3 = (float*)ws;
  const bool tc = a.allow_tensor_cores != 0;
  const bool umma_d = tc && a.gx && spvd_conv_umma_supported(a.kvol, a.cout, a.cin);
  const bool umma_w = tc && a.gw && spvd_conv_umma_supported(a.kvol, a.cin, a.cout);
  int rc = SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const float* g_tc = a.g;
  bool bias_done = false;
  if ((umma_d || umma_w) && a.n_out > 0) {
    if (a.gbias) {       // rounding pass and bias gradient share one read of g
      const int rpc = red_rows_per_cta(a.n_out);
      const int rp4 = red_rows_per_cta4(a.n_out, a.cout);
      if (vec4_ok(a.cout, a.g, gr)) k_round_colsum_v4<<<cdiv(a.n_out, rp4), kRedThreads, 0, s>>>(a.g, a.n_out, a.cout, rp4, gr, a.red_workspace);
      else k_round_colsum<<<cdiv(a.n_out, rpc), kRedThreads, 0, s>>>(a.g, a.n_out, a.cout, rpc, gr, a.red_workspace);
      k_f64_to_f32<<<cdiv(2 * a.cout, 256), 256, 0, s>>>(a.red_workspace, a.cout, 2 * a.cout, a.gbias);
      SPVD_LAUNCH_CHECK();
      bias_done = true;
    } else if ((rc = spvd_round_tf32(a.g, gr, (long long)a.n_out * a.cout, stream)) != SPVD_OK) {
      return rc;
    }
    g_tc = gr;
  }
  if (a.gx) {
    // data gradient: out' = in, W'[k'] = W[k]^T (k' = K-1-k for submanifold maps), map = the gradient map
    if (umma_d) {
      // sparse levels (few active offsets per voxel): pair-major on the forward pair list with the roles swapped
      const bool pairs = a.pair_in && a.pair_out && a.pair_tile_k && a.n_pair_tiles > 0 && a.n_pairs > 0 &&
                         (long long)a.n_pairs < 10ll * a.n_in;
      const int need_flip = pairs ? 0 : (a.flip ? 1 : 0);
      if (a.w_packed_t && a.w_packed_t_flip == need_flip) {      // prepared with the forward image (spvd_conv_pack_weights_train)
        wtp = const_cast<float*>(a.w_packed_t);
      } else {
        const long long total = (long long)a.kvol * a.cin * a.cout;
        k_pack_weights_transposed<<<cdiv(total, 256), 256, 0, s>>>(a.w, a.kvol, a.cin, a.cout, need_flip, wtp);
        SPVD_LAUNCH_CHECK();
      }
      if (pairs)
        rc = spvd_conv_fwd_pairs(g_tc, wtp, a.pair_out, a.pair_in, a.pair_tile_k, a.n_pair_tiles, -1, a.n_in, a.kvol, a.cout, a.cin,
                                 nullptr, nullptr, a.gx, SPVD_A_TF32, stream);
      else
        rc = spvd_conv_fwd_umma(g_tc, wtp, a.gmap_t, a.gmap_ld, a.gmap_mask, a.n_in, a.kvol, a.cout, a.cin, nullptr, nullptr, a.gx, 0,
                                0, SPVD_A_TF32, 0, stream);
    } else {
      if ((rc = spvd_conv_transpose_weights(a.w, a.kvol, a.cin, a.cout, a.flip, wt, stream)) != SPVD_OK) return rc;
      rc = spvd_conv_fwd_simt(a.g, wt, a.gmap_t, a.gmap_ld, a.gmap_mask, a.n_in, a.kvol, a.cout, a.cin, nullptr, nullptr, a.gx, stream);
    }
    if (rc != SPVD_OK) return rc;
  }
  if (a.gw) {
    float* dst = a.gw_transposed ? gw3 : a.gw;
    if (umma_w) {
      const float* x_tc = a.x;
      if (!a.x_is_tf32 && a.n_in > 0) {
        if ((rc = spvd_round_tf32(a.x, xr, (long long)a.n_in * a.cin, stream)) != SPVD_OK) return rc;
        x_tc = xr;
      }
      rc = spvd_conv_wgrad_umma(x_tc, g_tc, a.pair_in, a.pair_out, a.pair_tile_k, a.n_pair_tiles, a.pair_in ? 0 : a.n_out, a.kvol,
                                a.cin, a.cout, dst, stream);
    } else {
      rc = spvd_conv_wgrad(a.x, a.g, a.map_t, a.map_ld, a.n_out, a.kvol, a.cin, a.cout, dst, stream);
    }
    if (rc != SPVD_OK) return rc;
    if (a.gw_transposed)     // nn.Linear keeps [out, in]: [1, in, out] -> [1, out, in]
      if ((rc = spvd_conv_transpose_weights(gw3, a.kvol, a.cin, a.cout, 0, a.gw, stream)) != SPVD_OK) return rc;
  }
  if (a.gbias && !bias_done) rc = spvd_colsum(a.g, a.n_out, a.cout, a.gbias, a.red_workspace, stream);
  return rc;
}
