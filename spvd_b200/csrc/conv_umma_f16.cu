// f16-operand entry points of the tcgen05 sparse convolution (kernel and launchers: conv_umma_impl.cuh).
//
// Same tiles, same shared-memory image and descriptors as the TF32 path, but A and B are IEEE half (11-bit significand,
// exactly TF32's precision; fp32 accumulation in TMEM): tcgen05.mma.kind::f16 issues K = 16 per instruction at twice the
// TF32 rate, and a 128-byte row segment holds 64 input channels instead of 32, so a layer needs half the pipeline
// iterations and half the operand bytes.  Inputs are either f16 tensors written by the producing kernel (cp.async
// gather) or fp32 tensors converted on the fly (cvt.rn.f16x2, clamped to the finite range).
#include <cuda_fp16.h>

#include "conv_umma_impl.cuh"

namespace spvd {

// weights [kvol][cin][cout] fp32 -> per (k, 64-channel slab) image of the K-major SWIZZLE_128B B tile in f16:
// element (n, kk) at half offset ((k*nchunks + c)*cout + n)*64 + (((kk>>3) ^ (n&7))<<3) + (kk&7); channels past cin are 0
static __global__ void k_pack_weights_f16(const float* __restrict__ w, int kvol, int cin, int cout, int nchunks,
                                          __half* __restrict__ wp) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)kvol * nchunks * 64 * cout;
  if (t >= total) return;
  int n = (int)(t % cout);
  long long q = t / cout;
  int ci = (int)(q % (nchunks * 64));
  int k = (int)(q / (nchunks * 64));
  int c = ci / 64, kk = ci % 64;
  float v = ci < cin ? w[((size_t)k * cin + ci) * cout + n] : 0.f;
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  size_t dst = ((size_t)(k * nchunks + c) * cout + n) * 64 + ((((kk >> 3) ^ (n & 7)) << 3) | (kk & 7));
  wp[dst] = __float2half_rn(v);
}

template <int BN>
static int launch_f16_bn(bool a_half, const float* in, const float* wp, const int32_t* map_t, int ld, const uint32_t* tile_mask,
                         int n_out, int kvol, int cin, int cout, const float* bias, const float* residual, float* out,
                         int nsplit, cudaStream_t s) {
  if (a_half)
    return launch_umma<BN, 1, true, true, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
  return launch_umma<BN, 1, false, true, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
}

template <int BN>
static int launch_f16_pairs_bn(bool a_half, const float* in, const float* wp, const int32_t* pair_in, const int32_t* pair_out,
                               const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int kvol, int cin, int cout,
                               const float* bias, const float* residual, float* out, cudaStream_t s) {
  if (a_half)
    return launch_umma_pairs_a<BN, true, true>(in, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin,
                                               cout, bias, residual, out, s);
  return launch_umma_pairs_a<BN, false, true>(in, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin,
                                              cout, bias, residual, out, s);
}

int conv_fwd_umma_f16(const void* in, bool a_half, const void* w_packed, const int32_t* map_t, int ld,
                      const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                      const float* residual, float* out, int bn, int nsplit, cudaStream_t s) {
  const float* inf = reinterpret_cast<const float*>(in);
  const float* wp = reinterpret_cast<const float*>(w_packed);
  if (bn <= 0) bn = pick_bn(cout, cdiv(n_out, BM), kvol);
  SPVD_CHECK_ARG(cout % bn == 0, "spvd_conv_fwd_umma: slice width %d does not divide cout %d", bn, cout);
  const int max_iters = kvol * ((cin + 63) / 64);
  if (nsplit <= 0) nsplit = pick_split(cdiv(n_out, BM) * (cout / bn), max_iters);
  if (nsplit > max_iters) nsplit = max_iters;
  switch (bn) {
    case 32: return launch_f16_bn<32>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
    case 64: return launch_f16_bn<64>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
    case 96: return launch_f16_bn<96>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
    case 128: return launch_f16_bn<128>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
    case 192: return launch_f16_bn<192>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
    case 256: return launch_f16_bn<256>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s);
  }
  set_error("spvd_conv_fwd_umma: slice width %d is not instantiated", bn);
  return SPVD_ERR_INVALID;
}

int conv_fwd_pairs_f16(const void* in, bool a_half, const void* w_packed, const int32_t* pair_in, const int32_t* pair_out,
                       const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int kvol, int cin, int cout,
                       const float* bias, const float* residual, float* out, cudaStream_t s) {
  const float* inf = reinterpret_cast<const float*>(in);
  const float* wp = reinterpret_cast<const float*>(w_packed);
  switch (pick_bn(cout)) {
    case 32: return launch_f16_pairs_bn<32>(a_half, inf, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 64: return launch_f16_pairs_bn<64>(a_half, inf, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 96: return launch_f16_pairs_bn<96>(a_half, inf, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 128: return launch_f16_pairs_bn<128>(a_half, inf, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 192: return launch_f16_pairs_bn<192>(a_half, inf, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 256: return launch_f16_pairs_bn<256>(a_half, inf, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
  }
  return SPVD_ERR_INVALID;
}

}  // namespace spvd

using namespace spvd;

extern "C" size_t spvd_conv_packed_weight_count_f16(int kvol, int cin, int cout) {
  return (size_t)kvol * ((cin + 63) / 64) * 64 * cout;
}

extern "C" int spvd_conv_pack_weights_f16(const float* w, int kvol, int cin, int cout, void* w_packed, spvd_stream_t stream) {
  SPVD_CHECK_ARG(w && w_packed, "spvd_conv_pack_weights_f16: bad arguments");
  SPVD_CHECK_ARG(spvd_conv_umma_supported(kvol, cin, cout), "spvd_conv_pack_weights_f16: unsupported shape %dx%dx%d", kvol, cin, cout);
  const int nchunks = (cin + 63) / 64;
  long long total = (long long)kvol * nchunks * 64 * cout;
  k_pack_weights_f16<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, kvol, cin, cout, nchunks, (__half*)w_packed);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}
