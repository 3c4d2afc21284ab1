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
              v[2 * i] = __ldg(p);
              v[2 * i + 1] = __ldg(p + 1);
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
                  v[2 * i] = __ldg(p);
                  v[2 * i + 1] = __ldg(p + 1);
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
  static bool configured = false;
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

// ------------------------------------------------------------------------------------------------
// Persistent output-stationary convolution (f16 rows, f16 weights): "stream-K" over the flattened
// (tile, active offset, slab) iteration space.  The one-tile-per-CTA kernel is bound by the tensor work ISSUED per SM
// (measured: with every load removed it keeps 70-90 % of its time), and with 150-450 tiles on 148 SMs some SMs carry
// twice the work of others.  Here every CTA (one per SM) takes an equal, contiguous share of all iterations; tiles cut
// by a share boundary are finished with RED into the zero-filled output, whole tiles are stored.  4 producer warps run
// ahead on a deep stage ring (the whole shared memory of the SM), the MMA warp alternates between two TMEM accumulators,
// 4 epilogue warps drain one while the other is being filled.
// ------------------------------------------------------------------------------------------------
constexpr int kSkMaxTiles = 1024;
constexpr int kSkProducers = 128;                 // 4 producer warps
constexpr int kSkThreads = kSkProducers + 32 + 128; // + MMA warp + 4 epilogue warps
constexpr int kSkRows = BM * 8 / kSkProducers;     // rows gathered per producer thread and stage

template <int BN>
struct StreamKCfg {
  static constexpr int kBTileBytes = BN * 128;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  // two CTAs per SM (measured: ONE CTA's tcgen05 stream reaches ~55 % of the f16 rate, two co-resident CTAs ~90 %)
  static constexpr int kStagesFit = (104 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesFit > 4 ? 4 : kStagesFit;
  static constexpr int kBufs = 4 * BN <= 512 ? 2 : 1;      // accumulators per CTA: 2 x 2 x BN columns must fit TMEM
  static constexpr int kCols = kBufs * BN;
  static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + (kSkMaxTiles + 1) * 4;
  static constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(kStages >= 2 && kCols <= 256, "two CTAs per SM");
};

template <int BN>
__global__ void __launch_bounds__(kSkThreads, 2) k_conv_streamk_f16(
    const void* __restrict__ in_v, const void* __restrict__ w_packed, const int* __restrict__ map_t, int ld,
    const unsigned* __restrict__ tile_mask, int n_out, int kvol, int cin, int cout, const float* __restrict__ bias,
    const float* __restrict__ residual, float* __restrict__ out) {
  using Cfg = StreamKCfg<BN>;
  constexpr int S = Cfg::kStages;
  constexpr int NB = Cfg::kBufs;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_full = base + S * Cfg::kStageBytes;
  const uint32_t bar_empty = bar_full + 8 * S;
  const uint32_t bar_accfull = bar_empty + 8 * S;
  const uint32_t bar_accempty = bar_accfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * Cfg::kStageBytes + 8 * (2 * S + 4));
  int* s_pref = reinterpret_cast<int*>(smem + S * Cfg::kStageBytes + 256);   // [n_tiles + 1] iterations before each tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = (cin + 63) / 64;
  const int n_tiles = (n_out + BM - 1) / BM;
  const unsigned kmask = kvol >= 32 ? 0xffffffffu : ((1u << kvol) - 1u);
  const uint8_t* w_b = reinterpret_cast<const uint8_t*>(w_packed);
  const uint8_t* in_b = reinterpret_cast<const uint8_t*>(in_v);
  const uint32_t row_bytes = (uint32_t)cin * 2u;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, kSkProducers + 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accfull + 8 * b, 1);
      mbar_init(bar_accempty + 8 * b, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kSkProducers / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // iterations of every tile (geometry only: before the dependency wait), then an exclusive scan by warp 0
  for (int t = tid; t < n_tiles; t += kSkThreads)
    s_pref[t + 1] = __popc((tile_mask ? tile_mask[t] : kmask) & kmask) * nchunks;
  if (tid == 0) s_pref[0] = 0;
  __syncthreads();
  if (warp == 0) {
    const int per = (n_tiles + 31) / 32;
    const int b0 = 1 + lane * per, b1 = min(n_tiles + 1, b0 + per);
    int sum = 0;
    for (int i = b0; i < b1; ++i) sum += s_pref[i];
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - sum;
    for (int i = b0; i < b1; ++i) {
      run += s_pref[i];
      s_pref[i] = run;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const int total = s_pref[n_tiles];
  const int share = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int g0 = min(total, (int)blockIdx.x * share), g1 = min(total, g0 + share);
  int first_tile = 0;
  {
    int lo = 0, hi = n_tiles;            // last tile whose start is <= g0
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_pref[mid] <= g0) lo = mid; else hi = mid;
    }
    first_tile = lo;
  }
  pdl_wait();

  if (warp < kSkProducers / 32) {
    // ------------------------------------------------------------------ producers
    const int chunk16 = tid & 7, r0 = tid >> 3;   // rows r0 + (kSkProducers / 8) * i
    int li = 0;
    for (int g = g0, tile = first_tile; g < g1; ++tile) {
      const int tb = s_pref[tile], te = s_pref[tile + 1];
      if (te <= g) continue;
      const int it0 = g - tb, it1 = min(te, g1) - tb;
      const unsigned mask = (tile_mask ? tile_mask[tile] : kmask) & kmask;
      int rows[kSkRows], rows_next[kSkRows];
      int ka_loaded = -1;
      for (int it = it0; it < it1; ++it, ++li) {
        const int ka = it / nchunks, c = it - ka * nchunks;
        const int k = __fns(mask, 0, ka + 1);
        if (ka != ka_loaded) {
          if (ka_loaded >= 0 && ka == ka_loaded + 1) {
#pragma unroll
            for (int i = 0; i < kSkRows; ++i) rows[i] = rows_next[i];
          } else {
#pragma unroll
            for (int i = 0; i < kSkRows; ++i) {
              const int o = tile * BM + r0 + (kSkProducers / 8) * i;
              rows[i] = map_t ? __ldg(map_t + (size_t)k * ld + o) : (o < n_out ? o : -1);
            }
          }
          ka_loaded = ka;
          // the next offset's row indices travel under this offset's gathers
          const unsigned kn = __fns(mask, 0, ka + 2);   // 0xffffffff when this was the last active offset
          if (kn < 32u && map_t) {
#pragma unroll
            for (int i = 0; i < kSkRows; ++i) rows_next[i] = __ldg(map_t + (size_t)kn * ld + tile * BM + r0 + (kSkProducers / 8) * i);
          } else {
#pragma unroll
            for (int i = 0; i < kSkRows; ++i) rows_next[i] = -1;
          }
        }
        const int stage = li % S;
        const uint32_t parity = (uint32_t)(li / S) & 1u;
        const uint32_t a_s = base + stage * Cfg::kStageBytes;
        mbar_wait(bar_empty + 8 * stage, parity ^ 1u);
        if (tid == 0) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, Cfg::kBTileBytes);
          bulk_g2s(a_s + kATileBytes, w_b + ((size_t)(k * nchunks + c) * cout) * 128, Cfg::kBTileBytes, bar_full + 8 * stage);
        }
        const uint32_t col_b = (uint32_t)c * 128u + (uint32_t)chunk16 * 16u;
#pragma unroll
        for (int i = 0; i < kSkRows; ++i) {
          const int r = r0 + (kSkProducers / 8) * i;
          const uint32_t dst = a_s + r * 128 + ((chunk16 ^ (r & 7)) << 4);
          const bool ok = rows[i] >= 0 && col_b < row_bytes;
          cp_async16(dst, ok ? in_b + (size_t)rows[i] * row_bytes + col_b : in_b, ok ? 16u : 0u);
        }
        cp_async_mbar_arrive_noinc(bar_full + 8 * stage);
      }
      g = tb + it1;
    }
  } else if (warp == kSkProducers / 32) {
    // ------------------------------------------------------------------ MMA issuer
    int li = 0, t = 0;
    for (int g = g0, tile = first_tile; g < g1; ++tile) {
      const int tb = s_pref[tile], te = s_pref[tile + 1];
      if (te <= g) continue;
      const int it0 = g - tb, it1 = min(te, g1) - tb;
      const int buf = t % NB;
      mbar_wait(bar_accempty + 8 * buf, (((uint32_t)(t / NB)) & 1u) ^ 1u);
      tc_fence_after();
      // MMAs of kGroup consecutive stages are issued before their commits: a tcgen05.commit right behind every stage's
      // four MMAs leaves the tensor pipe idle while it drains (one CTA per SM: nobody else fills the gap)
      constexpr int kGroup = 2;
      for (int it = it0; it < it1; it += kGroup) {
        const int n_in = min(kGroup, it1 - it);
        for (int u = 0; u < n_in; ++u) {
          const int stage = (li + u) % S;
          mbar_wait(bar_full + 8 * stage, (uint32_t)((li + u) / S) & 1u);
          fence_proxy_async();
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_s = base + stage * Cfg::kStageBytes;
            const uint64_t adesc = make_desc(a_s), bdesc = make_desc(a_s + kATileBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_d + (uint32_t)(buf * BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), Cfg::kIdescF16,
                       (it + u > it0 || kk > 0) ? 1u : 0u);
          }
          __syncwarp();
        }
        if (lane == 0) {
          for (int u = 0; u < n_in; ++u) umma_commit(bar_empty + 8 * ((li + u) % S));
          if (it + n_in >= it1) umma_commit(bar_accfull + 8 * buf);
        }
        __syncwarp();
        li += n_in;
      }
      g = tb + it1;
      ++t;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    const bool wide = ((reinterpret_cast<size_t>(out) | reinterpret_cast<size_t>(residual)) & 31) == 0;
    int t = 0;
    for (int g = g0, tile = first_tile; g < g1; ++tile) {
      const int tb = s_pref[tile], te = s_pref[tile + 1];
      if (te <= g) continue;
      const int it0 = g - tb, it1 = min(te, g1) - tb;
      const int buf = t % NB;
      const bool first = it0 == 0;                       // this share starts the tile: it adds bias and residual
      const bool whole = first && it1 == te - tb;        // nobody else touches the tile: plain stores
      const int row = tile * BM + q * 32 + lane;
      mbar_wait(bar_accfull + 8 * buf, ((uint32_t)(t / NB)) & 1u);
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
        if (row < n_out) {
          float* dstf = out + (size_t)row * cout + cb * 32;
          const float4* bp = (bias && first) ? reinterpret_cast<const float4*>(bias + cb * 32) : nullptr;
          const float4* rp = (residual && first) ? reinterpret_cast<const float4*>(residual + (size_t)row * cout + cb * 32) : nullptr;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r[8 * j + e]);
            if (bp) {
              const float4 b0 = __ldg(bp + 2 * j), b1 = __ldg(bp + 2 * j + 1);
              o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
            }
            if (rp) {
              const float4 b0 = rp[2 * j], b1 = rp[2 * j + 1];
              o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
            }
            if (whole && wide) {
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dstf + 8 * j), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                           "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                           : "memory");
            } else if (whole) {
              reinterpret_cast<float4*>(dstf)[2 * j] = make_float4(o[0], o[1], o[2], o[3]);
              reinterpret_cast<float4*>(dstf)[2 * j + 1] = make_float4(o[4], o[5], o[6], o[7]);
            } else {
              red_add_v4(dstf + 8 * j, make_float4(o[0], o[1], o[2], o[3]));
              red_add_v4(dstf + 8 * j + 4, make_float4(o[4], o[5], o[6], o[7]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_accempty + 8 * buf);
      g = tb + it1;
      ++t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSkProducers / 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

template <int BN>
static int launch_streamk(const void* in, const void* wp, const int32_t* map_t, int ld, const uint32_t* tile_mask, int n_out,
                          int kvol, int cin, int cout, const float* bias, const float* residual, float* out, cudaStream_t s) {
  using Cfg = StreamKCfg<BN>;
  auto kern = k_conv_streamk_f16<BN>;
  static bool configured = false;
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  // tiles cut by a share boundary are accumulated with RED: the output starts from zero
  SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * cout * 4, s));
  const int n_tiles = cdiv(n_out, BM);
  const int max_iters = n_tiles * (kvol > 32 ? 32 : kvol) * ((cin + 63) / 64);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    SPVD_CUDA(cudaGetDevice(&dev));
    SPVD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  int ctas = 2 * n_sm;                      // two persistent CTAs per SM
  if (ctas > max_iters) ctas = max_iters;
  SPVD_CUDA(launch_pdl(kern, dim3(ctas), dim3(kSkThreads), (size_t)Cfg::kSmemBytes, s, in, wp, map_t, ld, tile_mask, n_out, kvol,
                       cin, cout, bias, residual, out));
  return SPVD_OK;
}

// ------------------------------------------------------------------------------------------------
// CTA-pair convolution (tcgen05 cta_group::2, f16 rows): two CTAs on the two SMs of a TPC own two consecutive 128-row
// tiles and walk the same (offset, slab) sequence (union of their tile masks); each gathers its own A tile but loads only
// HALF of the weight tile (its BN/2 output channels), and the leader issues M = 256 MMAs that read A and the B halves from
// both CTAs' shared memory and accumulate into both CTAs' TMEM.  The weight tile is 60 % of a stage, so a stage shrinks from
// 16 + BN/8 KB to 16 + BN/16 KB and the ring gets deeper in the same shared memory (BN = 192: 3 stages instead of 2).
//   full[s]  (leader): 128 cp.async arrivals + own weight bytes + one arrive relayed by the peer's idle MMA warp once the
//                      peer's own full[s] (rows + its weight half) completed;
//   empty[s], accum  : tcgen05.commit.cta_group::2 multicast to both CTAs.
// ------------------------------------------------------------------------------------------------
template <int BN>
struct TwoSmCfg {
  static constexpr int kBHalfBytes = BN * 64;                       // BN/2 rows x 128 B
  static constexpr int kStageBytes = kATileBytes + kBHalfBytes;
  static constexpr int kRowsBytes = 27 * BM * 4;
  static constexpr int kStagesFit = (112 * 1024 - kRowsBytes - 1280) / kStageBytes;
  static constexpr int kStages = kStagesFit > 4 ? 4 : kStagesFit;
  static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kRowsBytes;
  // M = 256 over the pair, N = BN, A = B = f16, D = f32
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  static_assert(BN % 32 == 0 && BN <= 256 && kStages >= 2, "CTA-pair tiles");
};

__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives on `bar` of both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(local_bar),
      "r"(cta_rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kUmmaThreads) k_conv_fwd_2sm_f16(const void* __restrict__ in_v, const void* __restrict__ w_packed,
                                                                  const int* __restrict__ map_t, int ld,
                                                                  const unsigned* __restrict__ tile_mask, int n_out, int kvol,
                                                                  int cin, int cout, const float* __restrict__ bias,
                                                                  const float* __restrict__ residual, float* __restrict__ out) {
  using Cfg = TwoSmCfg<BN>;
  constexpr int S = Cfg::kStages;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_full = base + S * Cfg::kStageBytes;     // [S]
  const uint32_t bar_empty = bar_full + 8 * S;               // [S]
  const uint32_t bar_accum = bar_empty + 8 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * Cfg::kStageBytes + 8 * (2 * S + 1));
  int* active_k = reinterpret_cast<int*>(smem + S * Cfg::kStageBytes + 8 * (2 * S + 1) + 16);
  int* s_rows = reinterpret_cast<int*>(smem + S * Cfg::kStageBytes + 256);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nchunks = (cin + 63) / 64;
  const int n_tiles = (n_out + BM - 1) / BM;
  const uint8_t* in_b = reinterpret_cast<const uint8_t*>(in_v);
  const uint8_t* w_b = reinterpret_cast<const uint8_t*>(w_packed);
  const uint32_t row_bytes = (uint32_t)cin * 2u;

  unsigned mask = kvol >= 32 ? 0xffffffffu : ((1u << kvol) - 1u);
  if (tile_mask) {
    const int t0 = ((int)blockIdx.x >> 1) << 1;
    unsigned u = 0;
    for (int t = t0; t < t0 + 2; ++t)
      if (t < n_tiles) u |= tile_mask[t];
    mask &= u;
  }
  const int total_iters = __popc(mask) * nchunks;
  const int per_split = (total_iters + (int)gridDim.z - 1) / (int)gridDim.z;
  const int it_begin = min(total_iters, (int)blockIdx.z * per_split);
  const int it_end = min(total_iters, it_begin + per_split);
  const int num_iters = it_end - it_begin;
  const bool split = gridDim.z > 1;
  if (split && num_iters == 0) return;   // uniform across the pair: before any barrier / allocation

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, kProducerThreads + 1 + (leader ? 1 : 0));   // + the peer's relay on the leader
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int j = 0;
    for (int k = 0; k < kvol && k < 32; ++k)
      if ((mask >> k) & 1u) active_k[j++] = k;
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before anything remote arrives
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  const int ka_begin = num_iters > 0 ? it_begin / nchunks : 0;
  const int ka_end = num_iters > 0 ? (it_end - 1) / nchunks + 1 : 0;
  for (int e = tid; e < (ka_end - ka_begin) * BM; e += kUmmaThreads) {
    const int slot = e / BM, r = e - slot * BM;
    const int o = m0 + r;
    int v = -1;
    if (map_t) {
      if (o < ld) v = __ldg(map_t + (size_t)active_k[ka_begin + slot] * ld + o);
    } else if (o < n_out) {
      v = o;
    }
    s_rows[e] = v;
  }
  pdl_wait();
  __syncthreads();

  if (warp < 4) {
    // ------------------------------------------------------------------ producers (both CTAs)
    const int chunk16 = tid & 7, r0 = tid >> 3;
    int rows[8];
    for (int it = it_begin; it < it_end; ++it) {
      const int li = it - it_begin;
      const int stage = li % S;
      const uint32_t parity = (uint32_t)(li / S) & 1u;
      const int ka = it / nchunks, c = it - ka * nchunks;
      const int k = active_k[ka];
      if (c == 0 || li == 0) {
        const int* sr = s_rows + (ka - ka_begin) * BM;
#pragma unroll
        for (int i = 0; i < 8; ++i) rows[i] = sr[r0 + 16 * i];
      }
      mbar_wait(bar_empty + 8 * stage, parity ^ 1u);
      const uint32_t a_s = base + stage * Cfg::kStageBytes;
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_full + 8 * stage, Cfg::kBHalfBytes);
        // this CTA's half of the weight tile: output channels n0 + rank * BN/2 ...
        bulk_g2s(a_s + kATileBytes, w_b + ((size_t)(k * nchunks + c) * cout + n0 + rank * (BN / 2)) * 128, Cfg::kBHalfBytes,
                 bar_full + 8 * stage);
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
    }
    // ------------------------------------------------------------------ epilogue (each CTA: its own 128 rows)
    if (num_iters > 0) {
      mbar_wait(bar_accum, 0);
      tc_fence_after();
    }
    const bool first = blockIdx.z == 0;
    const bool add_bias = bias != nullptr && first;
    const bool add_res = residual != nullptr && first;
    const int row = m0 + warp * 32 + lane;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      uint32_t r[32];
      if (num_iters > 0) {
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32);
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
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (row < n_out) {
        float* dstf = out + (size_t)row * cout + n0 + cb * 32;
        const float4* bp = add_bias ? reinterpret_cast<const float4*>(bias + n0 + cb * 32) : nullptr;
        const float4* rp = add_res ? reinterpret_cast<const float4*>(residual + (size_t)row * cout + n0 + cb * 32) : nullptr;
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
          if (split) red_add_v4(dstf + 4 * j, o);
          else reinterpret_cast<float4*>(dstf)[j] = o;
        }
      }
    }
  } else if (leader) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    for (int li = 0; li < num_iters; ++li) {
      const int stage = li % S;
      mbar_wait_cluster(bar_full + 8 * stage, (uint32_t)(li / S) & 1u);
      fence_proxy_async();
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_s = base + stage * Cfg::kStageBytes;
        const uint64_t adesc = make_desc(a_s), bdesc = make_desc(a_s + kATileBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_2sm(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), Cfg::kIdesc, (li > 0 || kk > 0) ? 1u : 0u);
        umma_commit_2sm(bar_empty + 8 * stage);
        if (li == num_iters - 1) umma_commit_2sm(bar_accum);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ relay (peer CTA): my stage is complete -> tell the leader
    for (int li = 0; li < num_iters; ++li) {
      const int stage = li % S;
      mbar_wait(bar_full + 8 * stage, (uint32_t)(li / S) & 1u);
      fence_proxy_async();      // my cp.async rows become visible to the tensor core that the leader drives
      if (lane == 0) mbar_arrive_remote(bar_full + 8 * stage, 0);
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while the pair may still read its shared memory / signal its barriers
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

template <int BN>
static int launch_2sm(const void* in, const void* wp, const int32_t* map_t, int ld, const uint32_t* tile_mask, int n_out, int kvol,
                      int cin, int cout, const float* bias, const float* residual, float* out, int nsplit, cudaStream_t s) {
  using Cfg = TwoSmCfg<BN>;
  auto kern = k_conv_fwd_2sm_f16<BN>;
  static bool configured = false;
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  if (nsplit > 1) SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * cout * 4, s));
  const int tiles = cdiv(cdiv(n_out, BM), 2) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, cout / BN, nsplit);
  cfg.blockDim = dim3(kUmmaThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  SPVD_CUDA(cudaLaunchKernelEx(&cfg, kern, in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out));
  return SPVD_OK;
}

template <int BN>
static int launch_f16_bn(bool a_half, const float* in, const float* wp, const int32_t* map_t, int ld, const uint32_t* tile_mask,
                         int n_out, int kvol, int cin, int cout, const float* bias, const float* residual, float* out,
                         int nsplit, cudaStream_t s, int mt, const int32_t* out_rows) {
  if (a_half && mt == 4)   // weight tiles multicast across a cluster of 2 CTAs
    return launch_umma_cluster<BN, 2, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  if constexpr (BN <= 256) {
    if (a_half && mt == 2)   // two tiles per CTA share every weight tile
      return launch_umma<BN, 2, true, false, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  }
  if (a_half)
    return launch_umma<BN, 1, true, true, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  return launch_umma<BN, 1, false, true, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
}

int conv_fwd_umma_f16(const void* in, bool a_half, const void* w_packed, const int32_t* map_t, int ld,
                      const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                      const float* residual, float* out, int bn, int nsplit, int mt, cudaStream_t s, const int32_t* out_rows) {
  const float* inf = reinterpret_cast<const float*>(in);
  const float* wp = reinterpret_cast<const float*>(w_packed);
  // f16 rows, one slice over all output channels, automatic configuration: the persistent stream-K kernel
  // mt == 7: CTA pairs (cta_group::2), each CTA loads half of every weight tile
  if (mt == 7 && a_half && !out_rows && cdiv(n_out, BM) >= 2) {
    int b = bn > 0 ? bn : pick_bn(cout, cdiv(n_out, BM), kvol);
    const int max_iters = kvol * ((cin + 63) / 64);
    int ns = nsplit > 0 ? nsplit : pick_split(cdiv(cdiv(n_out, BM), 2) * 2 * (cout / b), max_iters);
    if (ns > max_iters) ns = max_iters;
    switch (b) {
      case 64: return launch_2sm<64>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, ns, s);
      case 96: return launch_2sm<96>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, ns, s);
      case 128: return launch_2sm<128>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, ns, s);
      case 192: return launch_2sm<192>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, ns, s);
      case 256: return launch_2sm<256>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, ns, s);
      default: break;
    }
  }
  // mt == 6 (experimental, slower than the one-tile kernel on every SPVD layer -- see DESIGN.md 4.1): persistent stream-K
  if (mt == 6 && a_half && !out_rows && map_t && kvol > 1 && cdiv(n_out, BM) <= kSkMaxTiles) {
    switch (cout) {
      case 64: return launch_streamk<64>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, s);
      case 96: return launch_streamk<96>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, s);
      case 128: return launch_streamk<128>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, s);
      case 192: return launch_streamk<192>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, s);
      case 256: return launch_streamk<256>(in, w_packed, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, s);
      default: break;
    }
  }
  if (bn <= 0) bn = pick_bn(cout, cdiv(n_out, BM), kvol);
  SPVD_CHECK_ARG(cout % bn == 0, "spvd_conv_fwd_umma: slice width %d does not divide cout %d", bn, cout);
  const int max_iters = kvol * ((cin + 63) / 64);
  if (nsplit <= 0) nsplit = pick_split(cdiv(n_out, BM * (mt == 2 ? 2 : 1)) * (cout / bn), max_iters);
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
