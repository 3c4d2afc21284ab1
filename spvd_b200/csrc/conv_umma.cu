// TF32-operand entry points of the tcgen05 sparse convolution (kernel and launchers: conv_umma_impl.cuh).
#include "conv_umma_impl.cuh"


namespace spvd {
// conv_umma_f16.cu
int conv_fwd_umma_f16(const void* in, bool a_half, const void* w_packed, const int32_t* map_t, int ld,
                      const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                      const float* residual, float* out, int bn, int nsplit, int mt, cudaStream_t s,
                      const int32_t* out_rows = nullptr);
int conv_fwd_pairs_f16(const void* in, bool a_half, const void* w_packed, const int32_t* pair_in, const int32_t* pair_out,
                       const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int kvol, int cin, int cout,
                       const float* bias, const float* residual, float* out, cudaStream_t s);
}  // namespace spvd

using namespace spvd;

extern "C" size_t spvd_conv_packed_weight_count(int kvol, int cin, int cout) { return (size_t)kvol * cin * cout; }

extern "C" int spvd_conv_umma_supported(int kvol, int cin, int cout) {
  return kvol >= 1 && kvol <= 32 && cin % 32 == 0 && cout % 32 == 0 && cin > 0 && cout > 0;
}

extern "C" int spvd_conv_pack_weights(const float* w, int kvol, int cin, int cout, float* w_packed,
                                      spvd_stream_t stream) {
  SPVD_CHECK_ARG(w && w_packed, "spvd_conv_pack_weights: bad arguments");
  SPVD_CHECK_ARG(spvd_conv_umma_supported(kvol, cin, cout), "spvd_conv_pack_weights: unsupported shape %dx%dx%d", kvol, cin, cout);
  long long total = (long long)kvol * cin * cout;
  k_pack_weights<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, kvol, cin, cout, w_packed);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_conv_fwd_umma(const float* in, const float* w_packed, const int32_t* map_t, int ld,
                                  const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout,
                                  const float* bias, const float* residual, float* out, int bn, int nsplit,
                                  int a_is_tf32, int mt, spvd_stream_t stream) {
  return spvd_conv_fwd_umma_rows(in, w_packed, map_t, ld, tile_mask, nullptr, n_out, kvol, cin, cout, bias, residual, out, bn,
                                 nsplit, a_is_tf32, mt, stream);
}

extern "C" int spvd_conv_fwd_umma_rows(const float* in, const float* w_packed, const int32_t* map_t, int ld,
                                       const uint32_t* tile_mask, const int32_t* out_rows, int n_out, int kvol, int cin,
                                       int cout, const float* bias, const float* residual, float* out, int bn, int nsplit,
                                       int a_is_tf32, int mt, spvd_stream_t stream) {
  SPVD_CHECK_ARG(in && w_packed && out && n_out >= 0, "spvd_conv_fwd_umma: bad arguments");
  SPVD_CHECK_ARG(residual != out || residual == nullptr, "spvd_conv_fwd_umma: residual must not alias out");
  SPVD_CHECK_ARG(spvd_conv_umma_supported(kvol, cin, cout), "spvd_conv_fwd_umma: unsupported shape %dx%dx%d", kvol, cin, cout);
  SPVD_CHECK_ARG(map_t || kvol == 1, "spvd_conv_fwd_umma: a NULL map needs kvol == 1");
  SPVD_CHECK_ARG(!map_t || (ld % 128 == 0 && ld >= n_out), "spvd_conv_fwd_umma: ld %d must be a multiple of 128 >= n_out", ld);
  SPVD_CHECK_ARG(((size_t)in & 15) == 0 && ((size_t)out & 15) == 0 && ((size_t)w_packed & 15) == 0 &&
                     (!bias || ((size_t)bias & 15) == 0),
                 "spvd_conv_fwd_umma: pointers must be 16-byte aligned");
  if (n_out == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (a_is_tf32 >= SPVD_A_F32_TO_F16)   // f16 operands: w_packed is the spvd_conv_pack_weights_f16 image
    return conv_fwd_umma_f16(in, a_is_tf32 == SPVD_A_F16, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual,
                             out, bn, nsplit, mt, s, out_rows);
  if (bn <= 0) bn = pick_bn(cout, cdiv(n_out, BM), kvol);
  SPVD_CHECK_ARG(cout % bn == 0, "spvd_conv_fwd_umma: slice width %d does not divide cout %d", bn, cout);
  const int max_iters = kvol * (cin / 32);
  const bool aa = a_is_tf32 != 0;
  // default: one tile per CTA with a two-stage ring (<= 111 KB of shared memory), so two CTAs share an SM and one's
  // prologue / epilogue overlaps the other's main loop -- measured 10-25% faster than two tiles per CTA sharing the
  // weight tile (tools/conv_bench.py)
  if (mt <= 0) mt = 3;
  if (!aa && (mt == 2 || mt >= 4)) mt = mt == 2 ? 1 : 3;
#ifndef SPVD_EXPERIMENTAL
  if (mt == 2 || mt >= 4) mt = 3;
#endif
  if (nsplit <= 0) nsplit = pick_split(cdiv(n_out, BM * (mt == 2 ? 2 : 1)) * (cout / bn), max_iters);
  if (nsplit > max_iters) nsplit = max_iters;
  switch (bn) {
    case 32: return launch_umma_bn<32>(aa, mt, in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
    case 64: return launch_umma_bn<64>(aa, mt, in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
    case 96: return launch_umma_bn<96>(aa, mt, in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
    case 128: return launch_umma_bn<128>(aa, mt, in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
    case 192: return launch_umma_bn<192>(aa, mt, in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
    case 256: return launch_umma_bn<256>(aa, mt, in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  }
  set_error("spvd_conv_fwd_umma: slice width %d is not instantiated", bn);
  return SPVD_ERR_INVALID;
}

extern "C" int spvd_conv_fwd_pairs(const float* in, const float* w_packed, const int32_t* pair_in,
                                   const int32_t* pair_out, const int32_t* tile_k, int n_pair_tiles, int center_k,
                                   int n_out, int kvol, int cin, int cout, const float* bias, const float* residual,
                                   float* out, int a_is_tf32, spvd_stream_t stream) {
  SPVD_CHECK_ARG(in && w_packed && pair_in && pair_out && tile_k && out && n_out >= 0 && n_pair_tiles >= 0,
                 "spvd_conv_fwd_pairs: bad arguments");
  SPVD_CHECK_ARG(spvd_conv_umma_supported(kvol, cin, cout), "spvd_conv_fwd_pairs: unsupported shape %dx%dx%d", kvol, cin, cout);
  SPVD_CHECK_ARG((center_k >= 0 && center_k < kvol) || (!bias && !residual),
                 "spvd_conv_fwd_pairs: bias / residual need a centre offset (submanifold map)");
  const bool aa = a_is_tf32 != 0;
  SPVD_CHECK_ARG(residual != out || residual == nullptr, "spvd_conv_fwd_pairs: residual must not alias out");
  if (n_out == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (a_is_tf32 >= SPVD_A_F32_TO_F16) {
    return conv_fwd_pairs_f16(in, a_is_tf32 == SPVD_A_F16, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol,
                              cin, cout, bias, residual, out, s);
  }
  const int bn = pick_bn(cout);
  switch (bn) {
    case 32: return launch_umma_pairs<32>(aa, in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 64: return launch_umma_pairs<64>(aa, in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 96: return launch_umma_pairs<96>(aa, in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 128: return launch_umma_pairs<128>(aa, in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 192: return launch_umma_pairs<192>(aa, in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
    case 256: return launch_umma_pairs<256>(aa, in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol, cin, cout, bias, residual, out, s);
  }
  return SPVD_ERR_INVALID;
}

extern "C" int spvd_conv_fwd_simt(const float* in, const float* w, const int32_t* map_t, int ld,
                                  const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                                  const float* residual, float* out, spvd_stream_t stream);

extern "C" int spvd_conv_fwd(const float* in, const float* w, const float* w_packed, const int32_t* map_t, int ld,
                             const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                             float* out, int impl, spvd_stream_t stream) {
  bool umma_ok = w_packed && spvd_conv_umma_supported(kvol, cin, cout) && (!map_t || ld % 128 == 0);
  if (impl == SPVD_CONV_UMMA || (impl == SPVD_CONV_AUTO && umma_ok)) {
    SPVD_CHECK_ARG(umma_ok, "spvd_conv_fwd: the tcgen05 path needs packed weights, cin %% 32 == 0 and cout %% 32 == 0");
    return spvd_conv_fwd_umma(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, nullptr, out, 0, 0, 0, 0, stream);
  }
  SPVD_CHECK_ARG(w, "spvd_conv_fwd: the SIMT path needs the unpacked weights");
  return spvd_conv_fwd_simt(in, w, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, nullptr, out, stream);
}
