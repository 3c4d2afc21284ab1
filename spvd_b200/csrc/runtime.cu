// Native inference executor: runs a whole SPVD denoiser forward (models/spvd.py:432-457 in eval
// mode) from ONE C call.  The host side (spvd_b200/runtime.py) compiles the module tree once into a
// flat program of spvd_op_t records over a single activation arena; per step the interpreter below
// resolves row counts from the planned geometry and enqueues ~180 kernels back to back, instead
// of ~480 Python-dispatched launches.  Also holds the kernels only the executor needs: LayerNorm,
// per-cloud attention (models/modules/transformer.py:40-67,118-125), FiLM fused with the following
// BatchNorm+SiLU (models/modules/graph.py:35-41 + models/ddpm_unet_attn.py:57), channel concat.
#include <math.h>
#include <string.h>

#include <cuda_fp16.h>
#include "common.cuh"

namespace spvd {

__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ float rna_tf32_f(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// y = act(((1 + tt[b,c]) * x + tt[b,C+c]) * scale[c] + shift[c]);  tt rows are ldt floats apart
__global__ void k_film_affine4(const float4* __restrict__ x, const float* __restrict__ tt, int ldt,
                               const int4* __restrict__ coords, const float4* __restrict__ scale,
                               const float4* __restrict__ shift, long long total4, int c4, int act,
                               float4* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total4) return;
  int row = (int)(t / c4), q = (int)(t % c4);
  int b = coords[row].x;
  const float4* tb = reinterpret_cast<const float4*>(tt + (size_t)b * ldt);
  float4 v = x[t], sc = tb[q], sh = tb[c4 + q];     // (the FiLM table is an activation: plain loads under PDL)
  float4 s = __ldg(scale + q), o = __ldg(shift + q);
  v.x = fmaf(fmaf(1.f + sc.x, v.x, sh.x), s.x, o.x);
  v.y = fmaf(fmaf(1.f + sc.y, v.y, sh.y), s.y, o.y);
  v.z = fmaf(fmaf(1.f + sc.z, v.z, sh.z), s.z, o.z);
  v.w = fmaf(fmaf(1.f + sc.w, v.w, sh.w), s.w, o.w);
  if (act & 1) {
    v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
  }
  if (act & 4) {
    store_half4(y, t, v);
    return;
  }
  if (act & 2) {
    v.x = rna_tf32_f(v.x); v.y = rna_tf32_f(v.y); v.z = rna_tf32_f(v.z); v.w = rna_tf32_f(v.w);
  }
  y[t] = v;
}

__global__ void k_cat4(const float4* __restrict__ a, const float4* __restrict__ b, long long rows, int ca4, int cb4,
                       float4* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c4 = ca4 + cb4;
  if (t >= rows * c4) return;
  long long row = t / c4;
  int q = (int)(t % c4);
  y[t] = q < ca4 ? a[row * ca4 + q] : b[row * cb4 + (q - ca4)];
}

// one warp per row
__global__ void k_layernorm(const float* __restrict__ x, const float* __restrict__ gamma,
                            const float* __restrict__ beta, int rows, int c, float eps, int act,
                            float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * c;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s += xr[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / c;
  float v = 0.f;
  for (int i = lane; i < c; i += 32) {
    float d = xr[i] - mean;
    v = fmaf(d, d, v);
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  float rstd = rsqrtf(v / c + eps);
  for (int i = lane; i < c; i += 32) {
    float r = (xr[i] - mean) * rstd * gamma[i] + beta[i];
    if (act & 4) reinterpret_cast<__half*>(y)[(size_t)row * c + i] = __float2half_rn(fminf(fmaxf(r, -65504.f), 65504.f));
    else y[(size_t)row * c + i] = (act & 2) ? rna_tf32_f(r) : r;
  }
}

// Per-cloud multi-head attention over the voxels of one level.  qkv [rows, 3C] with channel layout
// which*C + head*D + d (the reference's reshape(B,N,3,H,D)); rows of a cloud are contiguous
// (batch-major voxel order).  Four lanes per query, keys/values staged in shared memory 128 at a
// time, online softmax (the reference's exp(x)*mask/sum is the same function).
template <int D>
__global__ void __launch_bounds__(128) k_attention(const float* __restrict__ qkv, const int* __restrict__ offsets,
                                                   int C, float scale, int act, float* __restrict__ out) {
  // 32 queries per CTA, 4 lanes per query: lane s of a query visits the keys j = s (mod 4) and the four partial
  // (max, sum, weighted values) states are merged with two shuffles at the end (the clouds are short -- a few hundred
  // rows -- so one thread per query leaves most of the GPU idle)
  constexpr int KB = D <= 32 ? 128 : 64;   // static shared memory stays under 48 KB
  constexpr int QB = 32, S = 4;
  __shared__ float4 Ks[KB][D / 4];
  __shared__ float4 Vs[KB][D / 4];
  pdl_trigger();
  pdl_wait();
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
      float4 t = qp[i];
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
      int j = e / (D / 4), i = e % (D / 4);
      const float4* kp = reinterpret_cast<const float4*>(qkv + (size_t)(r0 + k0 + j) * 3 * C + C + h * D);
      Ks[j][i] = kp[i];
      Vs[j][i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(kp) + C)[i];
    }
    __syncthreads();
    for (int j0 = slice; j0 < nk; j0 += 8 * S) {
      float s[8], mc = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float d = -INFINITY;
        if (j0 + jj * S < nk) {
          d = 0.f;
#pragma unroll
          for (int i = 0; i < D / 4; ++i) {
            float4 k4 = Ks[j0 + jj * S][i];
            d = fmaf(q[4 * i], k4.x, d);
            d = fmaf(q[4 * i + 1], k4.y, d);
            d = fmaf(q[4 * i + 2], k4.z, d);
            d = fmaf(q[4 * i + 3], k4.w, d);
          }
        }
        s[jj] = d;
        mc = fmaxf(mc, d);
      }
      float mn = fmaxf(m, mc);
      float alpha = __expf(m - mn);
      m = mn;
      l *= alpha;
#pragma unroll
      for (int i = 0; i < D; ++i) acc[i] *= alpha;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (j0 + jj * S < nk) {
          float p = __expf(s[jj] - mn);
          l += p;
#pragma unroll
          for (int i = 0; i < D / 4; ++i) {
            float4 v4 = Vs[j0 + jj * S][i];
            acc[4 * i] = fmaf(p, v4.x, acc[4 * i]);
            acc[4 * i + 1] = fmaf(p, v4.y, acc[4 * i + 1]);
            acc[4 * i + 2] = fmaf(p, v4.z, acc[4 * i + 2]);
            acc[4 * i + 3] = fmaf(p, v4.w, acc[4 * i + 3]);
          }
        }
      }
    }
  }
  // merge the four key slices of the query (a slice that saw no key has m = -inf, l = 0: weight 0)
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
    float inv = 1.f / l;
    float4* op = reinterpret_cast<float4*>(out + (size_t)(r0 + qi) * C + h * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      if ((i & (S - 1)) != slice) continue;     // the four lanes hold the same result: each writes a quarter of it
      float4 o = make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv);
      if (act & 4) {
        store_half4(out, ((long long)(r0 + qi) * C + h * D) / 4 + i, o);
        continue;
      }
      if (act & 2) {
        o.x = rna_tf32_f(o.x); o.y = rna_tf32_f(o.y); o.z = rna_tf32_f(o.z); o.w = rna_tf32_f(o.w);
      }
      op[i] = o;
    }
  }
}

template <int D>
static void launch_attention(const float* qkv, const int* offsets, int C, int heads, int nb, int max_len, int act,
                             float* out, cudaStream_t s) {
  dim3 grid(cdiv(max_len, 32), heads, nb);
  launch_pdl(k_attention<D>, grid, dim3(128), 0, s, qkv, offsets, C, 1.0f / sqrtf((float)D), act, out);
}

}  // namespace spvd

using namespace spvd;

extern "C" {
int spvd_scatter_mean_fwd(const float*, const int32_t*, int, int, float*, float*, int, spvd_stream_t);
int spvd_devoxelize_fwd(const float*, const int32_t*, const float*, int, int, float*, spvd_stream_t);
int spvd_scatter_wmean_fwd(const float*, const int32_t*, const float*, int, int, float*, int, spvd_stream_t);
int spvd_segment_mean_fwd(const float*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, int, int, int, float*,
                          int, spvd_stream_t);
int spvd_conv_fwd_simt(const float*, const float*, const int32_t*, int, const uint32_t*, int, int, int, int,
                       const float*, const float*, float*, spvd_stream_t);
int spvd_conv_fwd_umma(const float*, const float*, const int32_t*, int, const uint32_t*, int, int, int, int,
                       const float*, const float*, float*, int, int, int, int, spvd_stream_t);
int spvd_bn_silu_fwd(const float*, const float*, const float*, long long, int, int, float*, spvd_stream_t);
int spvd_conv_fwd_pairs(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int,
                        int, const float*, const float*, float*, int, spvd_stream_t);
int spvd_conv_fwd_pairs_simt(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int,
                             int, int, const float*, float*, spvd_stream_t);
int spvd_conv_sp_f16(const spvd_conv_sp_args_t*, int, int, spvd_stream_t);
}

extern "C" int spvd_film_affine_fwd(const float* x, const float* tt, int ldt, const int32_t* coords,
                                    const float* scale, const float* shift, long long rows, int c, int act, float* y,
                                    spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && tt && coords && scale && shift && y && rows >= 0 && c > 0 && c % 4 == 0 && ldt % 4 == 0,
                 "spvd_film_affine_fwd: bad arguments (c, ldt multiples of 4)");
  SPVD_CHECK_ARG((((size_t)x | (size_t)y | (size_t)tt | (size_t)scale | (size_t)shift) & 15) == 0,
                 "spvd_film_affine_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return SPVD_OK;
  SPVD_CUDA(launch_pdl(k_film_affine4, dim3(cdiv(rows * (c / 4), 256)), dim3(256), 0, (cudaStream_t)stream,
                       (const float4*)x, tt, ldt, (const int4*)coords, (const float4*)scale, (const float4*)shift,
                       rows * (c / 4), c / 4, act, (float4*)y));
  return SPVD_OK;
}

// [a | b] -> (i) the concatenation as f16 (operand of the block's 1x1 shortcut convolution) and (ii)
// act(cat * scale + shift) as f16 (operand of its first 3^3 convolution): the up-path blocks need both and nothing else
__global__ void k_cat_affine4_f16(const float4* __restrict__ a, const float4* __restrict__ b, long long rows, int ca4, int cb4,
                                  const float4* __restrict__ scale, const float4* __restrict__ shift, int act,
                                  void* __restrict__ raw, void* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = ca4 + cb4;
  if (t >= rows * c4) return;
  const long long row = t / c4;
  const int q = (int)(t % c4);
  float4 v = q < ca4 ? a[row * ca4 + q] : b[row * cb4 + (q - ca4)];
  store_half4(raw, t, v);
  const float4 s = __ldg(scale + q), o = __ldg(shift + q);
  v.x = fmaf(v.x, s.x, o.x); v.y = fmaf(v.y, s.y, o.y); v.z = fmaf(v.z, s.z, o.z); v.w = fmaf(v.w, s.w, o.w);
  if (act & 1) {
    v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
  }
  store_half4(y, t, v);
}

extern "C" int spvd_cat_affine_f16_fwd(const float* a, const float* b, long long rows, int ca, int cb, const float* scale,
                                       const float* shift, int act, void* cat_f16, void* y_f16, spvd_stream_t stream) {
  SPVD_CHECK_ARG(a && b && scale && shift && cat_f16 && y_f16 && rows >= 0 && ca > 0 && cb > 0 && ca % 4 == 0 && cb % 4 == 0,
                 "spvd_cat_affine_f16_fwd: bad arguments (channel counts multiples of 4)");
  if (rows == 0) return SPVD_OK;
  SPVD_CUDA(launch_pdl(k_cat_affine4_f16, dim3(cdiv(rows * ((ca + cb) / 4), 256)), dim3(256), 0, (cudaStream_t)stream,
                       (const float4*)a, (const float4*)b, rows, ca / 4, cb / 4, (const float4*)scale, (const float4*)shift, act,
                       cat_f16, y_f16));
  return SPVD_OK;
}

extern "C" int spvd_cat_fwd(const float* a, const float* b, long long rows, int ca, int cb, float* y,
                            spvd_stream_t stream) {
  SPVD_CHECK_ARG(a && b && y && rows >= 0 && ca > 0 && cb > 0 && ca % 4 == 0 && cb % 4 == 0,
                 "spvd_cat_fwd: bad arguments (channel counts multiples of 4)");
  if (rows == 0) return SPVD_OK;
  SPVD_CUDA(launch_pdl(k_cat4, dim3(cdiv(rows * ((ca + cb) / 4), 256)), dim3(256), 0, (cudaStream_t)stream,
                       (const float4*)a, (const float4*)b, rows, ca / 4, cb / 4, (float4*)y));
  return SPVD_OK;
}

extern "C" int spvd_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int c, float eps,
                                  int act, float* y, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && gamma && beta && y && rows >= 0 && c > 0, "spvd_layernorm_fwd: bad arguments");
  if (rows == 0) return SPVD_OK;
  SPVD_CUDA(launch_pdl(k_layernorm, dim3(cdiv(rows, 8)), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, rows, c, eps,
                       act, y));
  return SPVD_OK;
}

extern "C" int spvd_attention_fwd(const float* qkv, const int32_t* batch_offsets, int n_batch, int max_len, int c,
                                  int heads, int act, float* out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(qkv && batch_offsets && out && n_batch > 0 && heads > 0 && c % heads == 0,
                 "spvd_attention_fwd: bad arguments");
  if (max_len <= 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int d = c / heads;
  switch (d) {
    case 8: launch_attention<8>(qkv, batch_offsets, c, heads, n_batch, max_len, act, out, s); break;
    case 16: launch_attention<16>(qkv, batch_offsets, c, heads, n_batch, max_len, act, out, s); break;
    case 24: launch_attention<24>(qkv, batch_offsets, c, heads, n_batch, max_len, act, out, s); break;
    case 32: launch_attention<32>(qkv, batch_offsets, c, heads, n_batch, max_len, act, out, s); break;
    case 48: launch_attention<48>(qkv, batch_offsets, c, heads, n_batch, max_len, act, out, s); break;
    case 64: launch_attention<64>(qkv, batch_offsets, c, heads, n_batch, max_len, act, out, s); break;
    default:
      set_error("spvd_attention_fwd: head dim %d is not instantiated (8,16,24,32,48,64)", d);
      return SPVD_ERR_INVALID;
  }
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

// ------------------------------------------------------------------------------------------------
// timestep embedding: sinusoid -> BatchNorm1d (eval) -> SiLU -> Linear -> SiLU -> Linear (+ class row) -> SiLU,
// one CTA per cloud (models/utils.py:15-19, models/spvd.py:380-392); params = freqs[half] | bn scale, shift [n_temb]
// | W1 [n_emb, n_temb] | b1 | W2 [n_emb, n_emb] | b2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_time_embed(const long long* __restrict__ t, const long long* __restrict__ cls,
                                                    const float* __restrict__ P, const float* __restrict__ cond, int n_temb,
                                                    int n_emb, int act, float* __restrict__ out) {
  extern __shared__ float sm_te[];
  float* x = sm_te;            // [n_temb]
  float* h = sm_te + n_temb;   // [n_emb]
  const int b = blockIdx.x, half = n_temb / 2;
  const float* freqs = P;
  const float* bn_s = freqs + half;
  const float* bn_b = bn_s + n_temb;
  const float* W1 = bn_b + n_temb;
  const float* b1 = W1 + (size_t)n_emb * n_temb;
  const float* W2 = b1 + n_emb;
  const float* b2 = W2 + (size_t)n_emb * n_emb;
  const float tv = (float)t[b];
  for (int i = threadIdx.x; i < 2 * half; i += blockDim.x) {
    const float e = tv * freqs[i < half ? i : i - half];
    const float v = i < half ? sinf(e) : cosf(e);
    x[i] = silu_f(v * bn_s[i] + bn_b[i]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int j = warp; j < n_emb; j += n_warps) {
    float acc = 0.f;
    for (int i = lane; i < n_temb; i += 32) acc = fmaf(W1[(size_t)j * n_temb + i], x[i], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) h[j] = silu_f(acc + b1[j]);
  }
  __syncthreads();
  const float* crow = (cond && cls) ? cond + (size_t)cls[b] * n_emb : nullptr;
  for (int j = warp; j < n_emb; j += n_warps) {
    float acc = 0.f;
    for (int i = lane; i < n_emb; i += 32) acc = fmaf(W2[(size_t)j * n_emb + i], h[i], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float v = acc + b2[j] + (crow ? crow[j] : 0.f);
      out[(size_t)b * n_emb + j] = (act & 1) ? silu_f(v) : v;
    }
  }
}

extern "C" int spvd_time_embed_fwd(const int64_t* t, const int64_t* cls, const float* params, const float* cond_table,
                                   int n_batch, int n_temb, int n_emb, int act, float* out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(t && params && out && n_batch > 0 && n_temb > 0 && n_temb % 2 == 0 && n_emb > 0 && (n_temb + n_emb) * 4 <= 48 * 1024,
                 "spvd_time_embed_fwd: bad arguments");
  k_time_embed<<<n_batch, 256, (size_t)(n_temb + n_emb) * 4, (cudaStream_t)stream>>>(
      (const long long*)t, (const long long*)cls, params, cond_table, n_temb, n_emb, act, out);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

// ------------------------------------------------------------------------------------------------
// program interpreter
// ------------------------------------------------------------------------------------------------
static inline long long rows_of(int kind, const spvd_level_geom_t* lv, int n_levels, int n_points, int n_batch) {
  if (kind == SPVD_ROWS_POINTS) return n_points;
  if (kind == SPVD_ROWS_BATCH) return n_batch;
  if (kind >= 0 && kind < n_levels) return lv[kind].n;
  return -1;
}

extern "C" int spvd_program_run(const spvd_op_t* ops, int n_ops, const spvd_level_geom_t* levels, int n_levels,
                                int n_points, int n_batch, void* arena, size_t arena_bytes, void* const* conv_events,
                                spvd_stream_t stream) {
  SPVD_CHECK_ARG(ops && levels && arena && n_ops >= 0 && n_levels > 0, "spvd_program_run: bad arguments");
  char* base = (char*)arena;
  int conv_index = 0;
  auto at = [&](long long off) -> float* { return off < 0 ? nullptr : (float*)(base + off); };
  for (int i = 0; i < n_ops; ++i) {
    const spvd_op_t& o = ops[i];
    const long long rin = rows_of(o.rows_in, levels, n_levels, n_points, n_batch);
    const long long rout = rows_of(o.rows_out, levels, n_levels, n_points, n_batch);
    if (rin < 0 || rout < 0) {
      set_error("spvd_program_run: op %d has a bad row space", i);
      return SPVD_ERR_INVALID;
    }
    const long long out_bytes = rout * (long long)((o.op == SPVD_OP_CAT || o.op == SPVD_OP_CAT_AFFINE) ? o.c0 + o.c1 : o.c1) * 4;  // c1 = output channels
    const long long need = (o.dst >= 0 ? o.dst : o.aux) + out_bytes;
    if ((o.dst < 0 && !(o.op == SPVD_OP_CONV_SP && o.aux >= 0)) || (size_t)need > arena_bytes) {
      set_error("spvd_program_run: op %d writes past the arena (%lld > %zu)", i, need, arena_bytes);
      return SPVD_ERR_WORKSPACE;
    }
    const spvd_level_geom_t* lv = (o.level >= 0 && o.level < n_levels) ? &levels[o.level] : nullptr;
    int rc = SPVD_OK;
    switch (o.op) {
      case SPVD_OP_SCATTER_MEAN:
        SPVD_CHECK_ARG(lv && lv->p2v, "op %d: level %d has no point->voxel index", i, o.level);
        if (lv->p2v_cnt && lv->p2v_start && lv->p2v_order && lv->p2v_units && o.c0 % 4 == 0)
          rc = spvd_segment_mean_fwd(at(o.src0), lv->p2v_start, lv->p2v_cnt, lv->p2v_order, lv->p2v_units, n_points, n_points,
                                     o.c0, at(o.dst), lv->n, stream);
        else if (lv->p2v_w && o.c0 % 4 == 0)
          rc = spvd_scatter_wmean_fwd(at(o.src0), lv->p2v, lv->p2v_w, n_points, o.c0, at(o.dst), lv->n, stream);
        else
          rc = spvd_scatter_mean_fwd(at(o.src0), lv->p2v, n_points, o.c0, at(o.dst), at(o.aux), lv->n, stream);
        break;
      case SPVD_OP_DEVOX:
        SPVD_CHECK_ARG(lv && lv->devox_idx && lv->devox_w, "op %d: level %d has no devoxelize query", i, o.level);
        rc = spvd_devoxelize_fwd(at(o.src0), lv->devox_idx, lv->devox_w, n_points, o.c0, at(o.dst), stream);
        break;
      case SPVD_OP_CONV: {
        const int32_t* map = nullptr;
        const uint32_t* mask = nullptr;
        int ld = 0;
        if (o.map_kind != SPVD_MAP_NONE) {
          SPVD_CHECK_ARG(lv, "op %d: conv without a level", i);
          if (o.map_kind == SPVD_MAP_K3) { map = lv->kmap3; mask = lv->mask3; ld = lv->ld3; }
          else if (o.map_kind == SPVD_MAP_DOWN) { map = lv->kmap_down; mask = lv->mask_down; ld = lv->ld_down; }
          else { map = lv->kmap_up; mask = lv->mask_up; ld = lv->ld_up; }
          SPVD_CHECK_ARG(map, "op %d: level %d lacks kernel map kind %d", i, o.level, o.map_kind);
        }
        if (conv_events) SPVD_CUDA(cudaEventRecord((cudaEvent_t)conv_events[2 * conv_index], (cudaStream_t)stream));
        // operand mode of the tcgen05 kernels: TF32 or f16 weights image, src0 pre-converted by its producer or not
        const int a_mode = (o.flags & SPVD_FLAG_F16) ? ((o.flags & SPVD_FLAG_A_TF32) ? SPVD_A_F16 : SPVD_A_F32_TO_F16)
                                                     : ((o.flags & SPVD_FLAG_A_TF32) ? SPVD_A_TF32 : SPVD_A_F32);
        if (o.w_packed && o.map_kind == SPVD_MAP_K3 && (o.flags & SPVD_FLAG_A_TF32) && !(o.flags & SPVD_FLAG_SIMT) &&
            lv->n_pair_tiles > 0 && lv->pair_in)
          rc = spvd_conv_fwd_pairs(at(o.src0), o.w_packed, lv->pair_in, lv->pair_out, lv->pair_tile_k, lv->n_pair_tiles,
                                   o.kvol / 2, (int)rout, o.kvol, o.c0, o.c1, o.bias, at(o.src1), at(o.dst), a_mode, stream);
        else if (o.w_packed && (o.map_kind == SPVD_MAP_DOWN || o.map_kind == SPVD_MAP_UP) && !(o.flags & SPVD_FLAG_SIMT) &&
                 lv->n_dpair_tiles > 0 && lv->dpair_in && !o.bias && o.src1 < 0)
          // stride-2 conv: (fine -> coarse) pairs; its transpose runs on the same lists with the roles swapped
          rc = spvd_conv_fwd_pairs(at(o.src0), o.w_packed, o.map_kind == SPVD_MAP_DOWN ? lv->dpair_in : lv->dpair_out,
                                   o.map_kind == SPVD_MAP_DOWN ? lv->dpair_out : lv->dpair_in, lv->dpair_tile_k,
                                   lv->n_dpair_tiles, -1, (int)rout, o.kvol, o.c0, o.c1, nullptr, nullptr, at(o.dst),
                                   a_mode, stream);
        else if (o.w_packed && !(o.flags & SPVD_FLAG_SIMT))
          rc = spvd_conv_fwd_umma(at(o.src0), o.w_packed, map, ld, mask, (int)rout, o.kvol, o.c0, o.c1, o.bias,
                                  at(o.src1), at(o.dst), 0, 0, a_mode, 0, stream);
        else if (o.map_kind == SPVD_MAP_K3 && o.c0 <= 4 && o.src1 < 0 && lv->n_pair_tiles > 0 && lv->pair_in)
          rc = spvd_conv_fwd_pairs_simt(at(o.src0), o.w, lv->pair_in, lv->pair_out, lv->pair_tile_k, lv->n_pair_tiles,
                                        o.kvol / 2, (int)rout, o.kvol, o.c0, o.c1, o.bias, at(o.dst), stream);
        else
          rc = spvd_conv_fwd_simt(at(o.src0), o.w, map, ld, mask, (int)rout, o.kvol, o.c0, o.c1, o.bias, at(o.src1),
                                  at(o.dst), stream);
        if (conv_events) SPVD_CUDA(cudaEventRecord((cudaEvent_t)conv_events[2 * conv_index + 1], (cudaStream_t)stream));
        ++conv_index;
        break;
      }
      case SPVD_OP_CONV_SP: {
        spvd_conv_sp_args_t a = {};
        a.in = at(o.src0);
        a.w = o.w_packed;
        if (o.map_kind == SPVD_MAP_K3) {
          SPVD_CHECK_ARG(lv && lv->kmap3, "op %d: level %d lacks the 3^3 kernel map", i, o.level);
          if (lv->kmap3_sorted && lv->mask3_sorted && lv->perm3) {
            a.map_t = lv->kmap3_sorted; a.tile_mask = lv->mask3_sorted; a.out_rows = lv->perm3;
          } else {
            a.map_t = lv->kmap3; a.tile_mask = lv->mask3;
          }
          a.ld = lv->ld3;
        } else {
          SPVD_CHECK_ARG(o.map_kind == SPVD_MAP_NONE, "op %d: the sorted-tile convolution takes 3^3 or identity maps", i);
        }
        a.in2 = at(o.src2);
        a.w2 = o.w2_packed;
        a.cin2 = o.c2;
        a.bias = o.bias;
        a.residual = at(o.src1);
        if (o.src3 >= 0) {
          SPVD_CHECK_ARG(lv && lv->coords, "op %d: FiLM epilogue without level coordinates", i);
          a.film = at(o.src3) + o.i1;
          a.film_ld = o.i0;
          a.coords = lv->coords;
        }
        a.aff_scale = o.p0;
        a.aff_shift = o.p1;
        a.act = (o.flags & SPVD_FLAG_SILU) ? 1 : 0;
        a.out32 = at(o.dst);
        a.out16 = o.aux >= 0 ? (void*)at(o.aux) : nullptr;
        a.n_out = (int)rout; a.n_in = (int)rin; a.kvol = o.kvol; a.cin = o.c0; a.cout = o.c1;
        if (conv_events) SPVD_CUDA(cudaEventRecord((cudaEvent_t)conv_events[2 * conv_index], (cudaStream_t)stream));
        rc = spvd_conv_sp_f16(&a, 0, 0, stream);
        if (conv_events) SPVD_CUDA(cudaEventRecord((cudaEvent_t)conv_events[2 * conv_index + 1], (cudaStream_t)stream));
        ++conv_index;
        break;
      }
      case SPVD_OP_AFFINE:
        rc = spvd_bn_silu_fwd(at(o.src0), o.p0, o.p1, rin, o.c0, (o.flags & 3) | ((o.flags & SPVD_FLAG_F16) ? 4 : 0), at(o.dst), stream);
        break;
      case SPVD_OP_FILM_AFFINE:
        SPVD_CHECK_ARG(lv, "op %d: FiLM without a level", i);
        rc = spvd_film_affine_fwd(at(o.src0), at(o.src1) + o.i1, o.i0, lv->coords, o.p0, o.p1, rin, o.c0, (o.flags & 3) | ((o.flags & SPVD_FLAG_F16) ? 4 : 0),
                                  at(o.dst), stream);
        break;
      case SPVD_OP_CAT:
        rc = spvd_cat_fwd(at(o.src0), at(o.src1), rin, o.c0, o.c1, at(o.dst), stream);
        break;
      case SPVD_OP_CAT_AFFINE:
        SPVD_CHECK_ARG(o.aux >= 0 && o.p0 && o.p1, "op %d: cat+affine needs the second output and the affine", i);
        rc = spvd_cat_affine_f16_fwd(at(o.src0), at(o.src1), rin, o.c0, o.c1, o.p0, o.p1, o.flags & 1, at(o.aux), at(o.dst), stream);
        break;
      case SPVD_OP_LAYERNORM: {
        float ln_eps;
        memcpy(&ln_eps, &o.i0, sizeof(float));       // the module's own epsilon (float bits in i0)
        rc = spvd_layernorm_fwd(at(o.src0), o.p0, o.p1, (int)rin, o.c0, ln_eps, (o.flags & 3) | ((o.flags & SPVD_FLAG_F16) ? 4 : 0), at(o.dst), stream);
        break;
      }
      case SPVD_OP_ATTENTION:
        SPVD_CHECK_ARG(lv && lv->batch_offsets, "op %d: level %d has no per-cloud offsets", i, o.level);
        rc = spvd_attention_fwd(at(o.src0), lv->batch_offsets, n_batch, lv->max_per_batch, o.c1, o.i0, (o.flags & 3) | ((o.flags & SPVD_FLAG_F16) ? 4 : 0),
                                at(o.dst), stream);
        break;
      case SPVD_OP_TIME_EMBED:
        rc = spvd_time_embed_fwd((const int64_t*)o.p1, (const int64_t*)o.bias, o.w, o.p0, n_batch, o.c0, o.c1, o.flags & 1,
                                 at(o.dst), stream);
        break;
      default:
        set_error("spvd_program_run: op %d has unknown opcode %d", i, o.op);
        return SPVD_ERR_INVALID;
    }
    if (rc != SPVD_OK) return rc;
  }
  return SPVD_OK;
}
