// Sparse convolution, fp32 SIMT path (C1-C4): out[o] = sum_k in[map_t[k][o]] @ W[k] (+ bias), and
// the weight gradient.  This path accepts any channel count (the network input has Cin = 3) and is
// the numerical cross-check of the tcgen05 path in conv_umma.cu; the data gradient is the same
// forward kernel run on the transposed weights and the transposed map (see spvd_b200.h).
#include "common.cuh"

namespace spvd {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) k_conv_fwd_simt(const float* __restrict__ in, const float* __restrict__ w,
                                                       const int* __restrict__ map_t, int ld,
                                                       const unsigned* __restrict__ tile_mask, int n_out, int kvol,
                                                       int cin, int cout, const float* __restrict__ bias,
                                                       const float* __restrict__ residual,
                                                       float* __restrict__ out) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN];
  __shared__ int rows[TM];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  unsigned mask = 0xffffffffu;
  if (tile_mask && map_t && kvol <= 32) mask = tile_mask[m0 >> 7];
  float acc[4][4] = {};
  for (int k = 0; k < kvol; ++k) {
    if (!((mask >> (k & 31)) & 1u)) continue;
    __syncthreads();
    int any = 0;
    if (tid < TM) {
      int o = m0 + tid;
      int r = o < n_out ? (map_t ? map_t[(size_t)k * ld + o] : o) : -1;
      rows[tid] = r;
      any = r >= 0;
    }
    if (!__syncthreads_or(any)) continue;
    const float* wk = w + (size_t)k * cin * cout;
    for (int c0 = 0; c0 < cin; c0 += TK) {
      {  // A tile: 64 rows x 16 channels, 4 consecutive channels per thread
        int r = tid >> 2, cc = (tid & 3) * 4;
        int row = rows[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + cc + j;
          As[cc + j][r] = (row >= 0 && c < cin) ? in[(size_t)row * cin + c] : 0.f;
        }
      }
      {  // B tile: 16 x 64
        int kk = tid >> 4, nn = (tid & 15) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + kk, n = n0 + nn + j;
          Bs[kk][nn + j] = (c < cin && n < cout) ? wk[(size_t)c * cout + n] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int o = m0 + ty * 4 + i;
    if (o >= n_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < cout)
        out[(size_t)o * cout + n] =
            acc[i][j] + (bias ? bias[n] : 0.f) + (residual ? residual[(size_t)o * cout + n] : 0.f);
    }
  }
}

// Tiny-Cin direct convolution (the network input has 3 channels: 3^3 x 3 -> 32 at stride 1): one thread per
// (output row, output channel), weights in shared memory, map and input reads are warp broadcasts.
__global__ void __launch_bounds__(256) k_conv_fwd_small_cin(const float* __restrict__ in, const float* __restrict__ w,
                                                            const int* __restrict__ map_t, int ld, int n_out, int kvol,
                                                            int cin, int cout, const float* __restrict__ bias,
                                                            const float* __restrict__ residual, float* __restrict__ out) {
  extern __shared__ float s_w[];   // [kvol][cin][cout]
  pdl_trigger();
  for (int i = threadIdx.x; i < kvol * cin * cout; i += blockDim.x) s_w[i] = w[i];
  pdl_wait();
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_out * cout) return;
  const int o = (int)(t / cout), c = (int)(t % cout);
  float acc = bias ? bias[c] : 0.f;
  if (residual) acc += residual[t];
  for (int k0 = 0; k0 < kvol; k0 += 9) {   // 9 independent map reads in flight
    int r[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) r[j] = (k0 + j < kvol) ? (map_t ? __ldg(map_t + (size_t)(k0 + j) * ld + o) : o) : -1;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      if (r[j] < 0) continue;
      const float* x = in + (size_t)r[j] * cin;
      const float* wk = s_w + ((k0 + j) * cin) * cout + c;
      for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], wk[ci * cout], acc);
    }
  }
  out[t] = acc;
}

// nn.Linear with a tiny input or output width (the 3 -> 32 point embedding, the 32 -> 3 output head; 65536 rows):
// pure streaming.  kLinIn: one thread per (row, 4 output channels), 128-bit residual loads and stores.
// kLinOut: one thread per row, the row read as float4, cout <= 4 dot products.
__global__ void __launch_bounds__(256) k_linear_small_cin4(const float* __restrict__ in, const float* __restrict__ w, int n_out,
                                                           int cin, int cout4, const float* __restrict__ bias,
                                                           const float* __restrict__ residual, float* __restrict__ out) {
  extern __shared__ float4 s_w4[];   // [cin][cout4]
  pdl_trigger();
  for (int i = threadIdx.x; i < cin * cout4; i += blockDim.x) s_w4[i] = reinterpret_cast<const float4*>(w)[i];
  pdl_wait();
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_out * cout4) return;
  const int o = (int)(t / cout4), q = (int)(t % cout4);
  float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (residual) {
    const float4 r = reinterpret_cast<const float4*>(residual)[t];
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  for (int ci = 0; ci < cin; ++ci) {
    const float x = in[(size_t)o * cin + ci];
    const float4 wv = s_w4[ci * cout4 + q];
    acc.x = fmaf(x, wv.x, acc.x); acc.y = fmaf(x, wv.y, acc.y); acc.z = fmaf(x, wv.z, acc.z); acc.w = fmaf(x, wv.w, acc.w);
  }
  reinterpret_cast<float4*>(out)[t] = acc;
}

__global__ void __launch_bounds__(256) k_linear_small_cout(const float* __restrict__ in, const float* __restrict__ w, int n_out,
                                                           int cin4, int cout, const float* __restrict__ bias,
                                                           const float* __restrict__ residual, float* __restrict__ out) {
  extern __shared__ float s_wo[];   // [cin][cout]
  pdl_trigger();
  for (int i = threadIdx.x; i < cin4 * 4 * cout; i += blockDim.x) s_wo[i] = w[i];
  pdl_wait();
  __syncthreads();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < cout; ++c) acc[c] = (bias ? bias[c] : 0.f) + (residual ? residual[(size_t)o * cout + c] : 0.f);
  const float4* xr = reinterpret_cast<const float4*>(in) + (size_t)o * cin4;
  for (int i = 0; i < cin4; ++i) {
    const float4 x = xr[i];
    const float* wr = s_wo + (size_t)i * 4 * cout;
    for (int c = 0; c < cout; ++c)
      acc[c] = fmaf(x.w, wr[3 * cout + c], fmaf(x.z, wr[2 * cout + c], fmaf(x.y, wr[cout + c], fmaf(x.x, wr[c], acc[c]))));
  }
  for (int c = 0; c < cout; ++c) out[(size_t)o * cout + c] = acc[c];
}

// The same on a pair list (spvd_kmap_pairs): one CTA per tile of 128 (input row, output row) pairs of ONE offset, one
// warp instruction per pair (lanes = output channels), vector-free RED into the zero-filled output; the centre-offset
// tiles, which cover every output row exactly once, add the bias.  At stride 1 a voxel has ~2 of 27 neighbours, so this
// touches 13x fewer map entries than the row-major kernel above.
__global__ void __launch_bounds__(128) k_conv_fwd_small_cin_pairs(const float* __restrict__ in, const float* __restrict__ w,
                                                                  const int* __restrict__ pair_in,
                                                                  const int* __restrict__ pair_out,
                                                                  const int* __restrict__ tile_k, int center_k, int cin,
                                                                  int cout, const float* __restrict__ bias,
                                                                  float* __restrict__ out) {
  extern __shared__ float s_wk[];   // [cin][cout] of this tile's offset
  pdl_trigger();
  const int k = tile_k[blockIdx.x];
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) s_wk[i] = w[(size_t)k * cin * cout + i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * 128 + warp * 32 + lane;
  const int my_in = pair_in[p], my_out = pair_out[p];
  pdl_wait();
  __syncthreads();
  const bool add_bias = bias != nullptr && k == center_k;
  // every lane fetches the input row of ITS pair (32 independent loads in flight), then the warp walks the 32 pairs
  float mx[4] = {0.f, 0.f, 0.f, 0.f};
  if (my_in >= 0)
    for (int ci = 0; ci < cin; ++ci) mx[ci] = in[(size_t)my_in * cin + ci];
  for (int j = 0; j < 32; ++j) {
    const int ro = __shfl_sync(0xffffffffu, my_in < 0 ? -1 : my_out, j);
    float x[4];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) x[ci] = __shfl_sync(0xffffffffu, mx[ci], j);
    if (ro < 0) continue;
    for (int c = lane; c < cout; c += 32) {
      float acc = add_bias ? bias[c] : 0.f;
      for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], s_wk[ci * cout + c], acc);
      atomicAdd(out + (size_t)ro * cout + c, acc);
    }
  }
}

// gw[k][ci][co] += sum over a chunk of output rows of in[map[k][o]][ci] * gout[o][co]
constexpr int WI = 32, WO = 64, WR = 16;

__global__ void __launch_bounds__(256) k_conv_wgrad_simt(const float* __restrict__ in, const float* __restrict__ gout,
                                                         const int* __restrict__ map_t, int ld, int n_out, int kvol,
                                                         int cin, int cout, int rows_per_cta, int n_chunks,
                                                         float* __restrict__ gw) {
  __shared__ float Ai[WR][WI];
  __shared__ float Go[WR][WO];
  __shared__ int rows[WR];
  const int tid = threadIdx.x;
  const int k = blockIdx.x / n_chunks, chunk = blockIdx.x % n_chunks;
  const int ci0 = blockIdx.y * WI, co0 = blockIdx.z * WO;
  const int ti = tid >> 4, tj = tid & 15;  // ci rows ti*2..+1, co cols tj*4..+3
  float acc[2][4] = {};
  const int r_begin = chunk * rows_per_cta;
  const int r_end = min(n_out, r_begin + rows_per_cta);
  for (int r0 = r_begin; r0 < r_end; r0 += WR) {
    __syncthreads();
    int any = 0;
    if (tid < WR) {
      int o = r0 + tid;
      int r = o < r_end ? (map_t ? map_t[(size_t)k * ld + o] : o) : -1;
      rows[tid] = r;
      any = r >= 0;
    }
    if (!__syncthreads_or(any)) continue;
    for (int e = tid; e < WR * WI; e += 256) {
      int rr = e / WI, c = ci0 + e % WI;
      int row = rows[rr];
      Ai[rr][e % WI] = (row >= 0 && c < cin) ? in[(size_t)row * cin + c] : 0.f;
    }
    for (int e = tid; e < WR * WO; e += 256) {
      int rr = e / WO, c = co0 + e % WO;
      Go[rr][e % WO] = (rows[rr] >= 0 && c < cout) ? gout[(size_t)(r0 + rr) * cout + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < WR; ++rr) {
      float a0 = Ai[rr][ti * 2], a1 = Ai[rr][ti * 2 + 1];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float g = Go[rr][tj * 4 + j];
        acc[0][j] = fmaf(a0, g, acc[0][j]);
        acc[1][j] = fmaf(a1, g, acc[1][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int ci = ci0 + ti * 2 + i;
    if (ci >= cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = co0 + tj * 4 + j;
      if (co < cout && acc[i][j] != 0.f) atomicAdd(gw + ((size_t)k * cin + ci) * cout + co, acc[i][j]);
    }
  }
}

__global__ void k_transpose_weights(const float* __restrict__ w, int kvol, int cin, int cout, int flip,
                                    float* __restrict__ wt) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long per = (long long)cin * cout;
  if (t >= per * kvol) return;
  int k = (int)(t / per);
  int rem = (int)(t % per);
  int co = rem / cin, ci = rem % cin;  // destination layout [k'][co][ci]
  int ks = flip ? kvol - 1 - k : k;
  wt[t] = w[((size_t)ks * cin + ci) * cout + co];
}

__global__ void k_round_tf32(const float* __restrict__ x, float* __restrict__ y, long long count) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x[t]));
  y[t] = __uint_as_float(r);
}

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_conv_fwd_simt(const float* in, const float* w, const int32_t* map_t, int ld,
                                  const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                                  const float* residual, float* out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(in && w && out && n_out >= 0 && kvol >= 1 && cin >= 1 && cout >= 1, "spvd_conv_fwd: bad arguments");
  SPVD_CHECK_ARG(map_t || kvol == 1, "spvd_conv_fwd: a NULL map needs kvol == 1");
  SPVD_CHECK_ARG(!map_t || ld >= n_out, "spvd_conv_fwd: ld %d < n_out %d", ld, n_out);
  if (n_out == 0) return SPVD_OK;
  const bool al16 = (((size_t)in | (size_t)w | (size_t)out | (size_t)bias | (size_t)residual) & 15) == 0;
  if (!map_t && kvol == 1 && cin <= 4 && cout % 4 == 0 && cin * cout * 4 <= 40 * 1024 && al16) {
    SPVD_CUDA(launch_pdl(k_linear_small_cin4, dim3(cdiv((long long)n_out * (cout / 4), 256)), dim3(256), (size_t)cin * cout * 4,
                         (cudaStream_t)stream, in, w, n_out, cin, cout / 4, bias, residual, out));
    return SPVD_OK;
  }
  if (!map_t && kvol == 1 && cout <= 4 && cin % 4 == 0 && cin * cout * 4 <= 40 * 1024 && al16) {
    SPVD_CUDA(launch_pdl(k_linear_small_cout, dim3(cdiv(n_out, 256)), dim3(256), (size_t)cin * cout * 4, (cudaStream_t)stream, in, w,
                         n_out, cin / 4, cout, bias, residual, out));
    return SPVD_OK;
  }
  if (cin <= 4 && (size_t)kvol * cin * cout * 4 <= 40 * 1024) {
    SPVD_CUDA(launch_pdl(k_conv_fwd_small_cin, dim3(cdiv((long long)n_out * cout, 256)), dim3(256),
                         (size_t)kvol * cin * cout * 4, (cudaStream_t)stream, in, w, map_t, ld, n_out, kvol, cin, cout, bias,
                         residual, out));
    return SPVD_OK;
  }
  dim3 grid(cdiv(n_out, TM), cdiv(cout, TN));
  SPVD_CUDA(launch_pdl(k_conv_fwd_simt, grid, dim3(256), 0, (cudaStream_t)stream, in, w, map_t, ld, tile_mask, n_out, kvol,
                       cin, cout, bias, residual, out));
  return SPVD_OK;
}

extern "C" int spvd_conv_fwd_pairs_simt(const float* in, const float* w, const int32_t* pair_in, const int32_t* pair_out,
                                        const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int kvol, int cin,
                                        int cout, const float* bias, float* out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(in && w && pair_in && pair_out && tile_k && out && n_out >= 0 && n_pair_tiles >= 0 && kvol >= 1,
                 "spvd_conv_fwd_pairs_simt: bad arguments");
  SPVD_CHECK_ARG(cin >= 1 && cin <= 4 && cout >= 1 && (size_t)cin * cout * 4 <= 40 * 1024,
                 "spvd_conv_fwd_pairs_simt: instantiated for cin <= 4 (the network input)");
  SPVD_CHECK_ARG((center_k >= 0 && center_k < kvol) || !bias, "spvd_conv_fwd_pairs_simt: a bias needs a centre offset");
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * cout * 4, s));
  if (n_out == 0 || n_pair_tiles == 0) return SPVD_OK;
  SPVD_CUDA(launch_pdl(k_conv_fwd_small_cin_pairs, dim3(n_pair_tiles), dim3(128), (size_t)cin * cout * 4, s, in, w, pair_in,
                       pair_out, tile_k, center_k, cin, cout, bias, out));
  return SPVD_OK;
}

extern "C" int spvd_conv_wgrad(const float* in, const float* gout, const int32_t* map_t, int ld, int n_out, int kvol,
                               int cin, int cout, float* gw, spvd_stream_t stream) {
  SPVD_CHECK_ARG(in && gout && gw && n_out >= 0 && kvol >= 1 && cin >= 1 && cout >= 1, "spvd_conv_wgrad: bad arguments");
  SPVD_CHECK_ARG(map_t || kvol == 1, "spvd_conv_wgrad: a NULL map needs kvol == 1");
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(gw, 0, (size_t)kvol * cin * cout * 4, s));
  if (n_out == 0) return SPVD_OK;
  int rows_per_cta = 1024;
  int n_chunks = cdiv(n_out, rows_per_cta);
  dim3 grid(kvol * n_chunks, cdiv(cin, WI), cdiv(cout, WO));
  k_conv_wgrad_simt<<<grid, 256, 0, s>>>(in, gout, map_t, ld, n_out, kvol, cin, cout, rows_per_cta, n_chunks, gw);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_conv_transpose_weights(const float* w, int kvol, int cin, int cout, int flip, float* w_t,
                                           spvd_stream_t stream) {
  SPVD_CHECK_ARG(w && w_t && kvol >= 1 && cin >= 1 && cout >= 1, "spvd_conv_transpose_weights: bad arguments");
  long long total = (long long)kvol * cin * cout;
  k_transpose_weights<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, kvol, cin, cout, flip, w_t);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_round_tf32(const float* x, float* y, long long count, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && y && count >= 0, "spvd_round_tf32: bad arguments");
  if (count == 0) return SPVD_OK;
  k_round_tf32<<<cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(x, y, count);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
