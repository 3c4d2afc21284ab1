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


// ------------------------------------------------------------------------------------------------
// Persistent pair-major convolution (f16 operands).  A pair tile has only cin/64 pipeline iterations, so a CTA per tile
// spends most of its life in the prologue (barrier init, TMEM allocation), the first load latency and the RED epilogue.
// Here a CTA walks many pair tiles: 4 producer warps run ahead through the (tile, slab) sequence on one stage ring, the
// MMA warp alternates between two TMEM accumulators, and 4 epilogue warps scatter tile i (tcgen05.ld + RED) while tile
// i + 1 is loaded and multiplied.  Two CTAs per SM (<= 128 accumulator columns x 2 each).
// ------------------------------------------------------------------------------------------------
constexpr int kPpThreads = 32 * 9;   // warps 0-3 producers, 4 MMA, 5-8 epilogue

template <int BN>
struct PairsCfg {
  static constexpr int kBTileBytes = BN * 128;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (4 * kStageBytes <= 100 * 1024) ? 4 : (3 * kStageBytes <= 100 * 1024) ? 3 : 2;
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(BN <= 128, "two accumulators of two co-resident CTAs must fit the 512 TMEM columns");
};

template <int BN, bool kHalfIn>
__global__ void __launch_bounds__(kPpThreads, 2) k_conv_pairs_persistent(
    const void* __restrict__ in_v, const void* __restrict__ w_packed, const int* __restrict__ pair_in,
    const int* __restrict__ pair_out, const int* __restrict__ tile_k, int n_tiles, int center_k, int n_out, int cin, int cout,
    const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ out) {
  using Cfg = PairsCfg<BN>;
  constexpr int S = Cfg::kStages;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_full = base + S * Cfg::kStageBytes;       // full[S]
  const uint32_t bar_empty = bar_full + 8 * S;                 // empty[S]
  const uint32_t bar_accfull = bar_empty + 8 * S;              // accumulator ready [2]
  const uint32_t bar_accempty = bar_accfull + 16;              // accumulator drained [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * Cfg::kStageBytes + 8 * (2 * S + 4));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;
  const int nchunks = (cin + 63) / 64;
  const uint8_t* w_b = reinterpret_cast<const uint8_t*>(w_packed);

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, kProducerThreads + 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accfull + 8 * b, 1);
      mbar_init(bar_accempty + 8 * b, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_wait();   // geometry (pair lists) and weights were ready before; activations and the zeroed output follow

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const int chunk16 = tid & 7, r0 = tid >> 3;
    int li = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int k = __ldg(tile_k + tile);
      int rows[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rows[i] = __ldg(pair_in + (size_t)tile * BM + r0 + 16 * i);
      for (int c = 0; c < nchunks; ++c, ++li) {
        const int stage = li % S;
        const uint32_t parity = (uint32_t)(li / S) & 1u;
        const uint32_t a_s = base + stage * Cfg::kStageBytes;
        if constexpr (kHalfIn) {
          const uint8_t* in_b = reinterpret_cast<const uint8_t*>(in_v);
          const uint32_t row_bytes = (uint32_t)cin * 2u;
          mbar_wait(bar_empty + 8 * stage, parity ^ 1u);
          if (tid == 0) {
            mbar_arrive_expect_tx(bar_full + 8 * stage, Cfg::kBTileBytes);
            bulk_g2s(a_s + kATileBytes, w_b + ((size_t)(k * nchunks + c) * cout + n0) * 128, Cfg::kBTileBytes, bar_full + 8 * stage);
          }
          const uint32_t col_b = (uint32_t)c * 128u + (uint32_t)chunk16 * 16u;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r0 + 16 * i;
            const uint32_t dst = a_s + r * 128 + ((chunk16 ^ (r & 7)) << 4);
            const bool ok = rows[i] >= 0 && col_b < row_bytes;
            cp_async16(dst, ok ? in_b + (size_t)rows[i] * row_bytes + col_b : in_b, ok ? 16u : 0u);
          }
          cp_async_mbar_arrive_noinc(bar_full + 8 * stage);
        } else {
          // fp32 rows converted on the fly, four rows at a time (this kernel keeps two CTAs per SM: <= 112 registers)
          const float* in = reinterpret_cast<const float*>(in_v);
          const int ch = c * 64 + chunk16 * 8;
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[2 * i] = v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rows[i] >= 0 && ch < cin) {
              const float4* p = reinterpret_cast<const float4*>(in + (size_t)rows[i] * cin + ch);
              v[2 * i] = *p;
              v[2 * i + 1] = *(p + 1);
            }
          }
          mbar_wait(bar_empty + 8 * stage, parity ^ 1u);
          if (tid == 0) {
            mbar_arrive_expect_tx(bar_full + 8 * stage, Cfg::kBTileBytes);
            bulk_g2s(a_s + kATileBytes, w_b + ((size_t)(k * nchunks + c) * cout + n0) * 128, Cfg::kBTileBytes, bar_full + 8 * stage);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                v[2 * i] = v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rows[4 + i] >= 0 && ch < cin) {
                  const float4* p = reinterpret_cast<const float4*>(in + (size_t)rows[4 + i] * cin + ch);
                  v[2 * i] = *p;
                  v[2 * i + 1] = *(p + 1);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r0 + 16 * (4 * h + i);
              const uint32_t dst = a_s + r * 128 + ((chunk16 ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_half2_sat(v[2 * i].x, v[2 * i].y)),
                           "r"(pack_half2_sat(v[2 * i].z, v[2 * i].w)), "r"(pack_half2_sat(v[2 * i + 1].x, v[2 * i + 1].y)),
                           "r"(pack_half2_sat(v[2 * i + 1].z, v[2 * i + 1].w))
                           : "memory");
            }
          }
          fence_proxy_async();
          mbar_arrive(bar_full + 8 * stage);
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    int li = 0, t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      mbar_wait(bar_accempty + 8 * buf, (((uint32_t)t >> 1) & 1u) ^ 1u);   // the epilogue drained this accumulator
      tc_fence_after();
      for (int c = 0; c < nchunks; ++c, ++li) {
        const int stage = li % S;
        mbar_wait(bar_full + 8 * stage, (uint32_t)(li / S) & 1u);
        if constexpr (kHalfIn) fence_proxy_async();
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_s = base + stage * Cfg::kStageBytes;
          const uint64_t adesc = make_desc(a_s), bdesc = make_desc(a_s + kATileBytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_d + (uint32_t)(buf * BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), Cfg::kIdescF16,
                     (c > 0 || kk > 0) ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);
          if (c == nchunks - 1) umma_commit(bar_accfull + 8 * buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      int row = __ldg(pair_out + (size_t)tile * BM + q * 32 + lane);
      const bool first = __ldg(tile_k + tile) == center_k;
      mbar_wait(bar_accfull + 8 * buf, ((uint32_t)t >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
              "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
              "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row >= 0 && row < n_out) {
          float* dstf = out + (size_t)row * cout + n0 + cb * 32;
          const float4* bp = (bias && first) ? reinterpret_cast<const float4*>(bias + n0 + cb * 32) : nullptr;
          const float4* rp = (residual && first) ? reinterpret_cast<const float4*>(residual + (size_t)row * cout + n0 + cb * 32) : nullptr;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                   __uint_as_float(r[4 * j + 3]));
            if (bp) {
              const float4 b = __ldg(bp + j);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (rp) {
              const float4 b = rp[j];
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            red_add_v4(dstf + 4 * j, o);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_accempty + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

template <int BN, bool kHalfIn>
static int launch_pairs_persistent(const void* in, const void* wp, const int32_t* pair_in, const int32_t* pair_out,
                                   const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int cin, int cout,
                                   const float* bias, const float* residual, float* out, cudaStream_t s) {
  using Cfg = PairsCfg<BN>;
  auto kern = k_conv_pairs_persistent<BN, kHalfIn>;
  static bool configured_dev[64] = {};   // the attribute is per device: one flag per device, not per process
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * cout * 4, s));
  if (n_pair_tiles == 0) return SPVD_OK;
  const int slices = cout / BN;
  int ctas = (2 * 148) / slices;            // two CTAs per SM, all slices of a tile range resident together
  if (ctas < 1) ctas = 1;
  if (ctas > n_pair_tiles) ctas = n_pair_tiles;
  SPVD_CUDA(launch_pdl(kern, dim3(ctas, slices), dim3(kPpThreads), (size_t)Cfg::kSmemBytes, s, in, wp, pair_in, pair_out, tile_k,
                       n_pair_tiles, center_k, n_out, cin, cout, bias, residual, out));
  return SPVD_OK;
}

template <int BN>
static int launch_f16_bn(bool a_half, const float* in, const float* wp, const int32_t* map_t, int ld, const uint32_t* tile_mask,
                         int n_out, int kvol, int cin, int cout, const float* bias, const float* residual, float* out,
                         int nsplit, cudaStream_t s, int mt, const int32_t* out_rows) {
#ifdef SPVD_EXPERIMENTAL   // measured no faster on any SPVD layer (tools/_sorted_sweep.py): not part of the product library
  if (a_half && mt == 4)   // weight tiles multicast across a cluster of 2 CTAs
    return launch_umma_cluster<BN, 2, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  if (a_half && mt == 2)   // two tiles per CTA share every weight tile
    return launch_umma<BN, 2, true, false, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
#else
  (void)mt;
#endif
  if (a_half)
    return launch_umma<BN, 1, true, true, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  return launch_umma<BN, 1, false, true, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
}

int conv_fwd_umma_f16(const void* in, bool a_half, const void* w_packed, const int32_t* map_t, int ld,
                      const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                      const float* residual, float* out, int bn, int nsplit, int mt, cudaStream_t s, const int32_t* out_rows) {
  const float* inf = reinterpret_cast<const float*>(in);
  const float* wp = reinterpret_cast<const float*>(w_packed);
  if (bn <= 0) bn = pick_bn(cout, cdiv(n_out, BM), kvol);
  SPVD_CHECK_ARG(cout % bn == 0, "spvd_conv_fwd_umma: slice width %d does not divide cout %d", bn, cout);
  const int max_iters = kvol * ((cin + 63) / 64);
  if (nsplit <= 0) nsplit = pick_split(cdiv(n_out, BM) * (cout / bn), max_iters);
  if (nsplit > max_iters) nsplit = max_iters;
  switch (bn) {
    case 32: return launch_f16_bn<32>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, mt, out_rows);
    case 64: return launch_f16_bn<64>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, mt, out_rows);
    case 96: return launch_f16_bn<96>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, mt, out_rows);
    case 128: return launch_f16_bn<128>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, mt, out_rows);
    case 192: return launch_f16_bn<192>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, mt, out_rows);
    case 256: return launch_f16_bn<256>(a_half, inf, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, mt, out_rows);
  }
  set_error("spvd_conv_fwd_umma: slice width %d is not instantiated", bn);
  return SPVD_ERR_INVALID;
}

int conv_fwd_pairs_f16(const void* in, bool a_half, const void* w_packed, const int32_t* pair_in, const int32_t* pair_out,
                       const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int kvol, int cin, int cout,
                       const float* bias, const float* residual, float* out, cudaStream_t s) {
  (void)kvol;
  // widest slice <= 128 that divides cout (two accumulators per CTA, two CTAs per SM)
  const int widths[] = {128, 96, 64, 32};
  int bn = 32;
  for (int w : widths)
    if (cout % w == 0) { bn = w; break; }
#define SPVD_PP(BNV)                                                                                                        \
  case BNV:                                                                                                                 \
    return a_half ? launch_pairs_persistent<BNV, true>(in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, \
                                                       cin, cout, bias, residual, out, s)                                   \
                  : launch_pairs_persistent<BNV, false>(in, w_packed, pair_in, pair_out, tile_k, n_pair_tiles, center_k,    \
                                                        n_out, cin, cout, bias, residual, out, s);
  switch (bn) {
    SPVD_PP(32)
    SPVD_PP(64)
    SPVD_PP(96)
    SPVD_PP(128)
  }
#undef SPVD_PP
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
