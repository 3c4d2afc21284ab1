// Persistent "sorted implicit GEMM" sparse convolution on tcgen05 with a fused epilogue (f16 operands, fp32 accumulate
// in TMEM) -- the convolution kernel of the inference executor (C1-C4 of SURVEY.md section 8; replaces the conv launch
// inside spnn.Conv3d, models/spvd.py:35,38,113,229, plus the glue around it, models/ddpm_unet_attn.py:48-65).
//
// What it changes with respect to the one-tile-per-CTA kernel (conv_umma_impl.cuh), and why (tools/_conv_trace.py):
//   * a CTA of the old kernel spent ~10 us staging map rows and 4-5 us in a serial epilogue for tiles whose main loop is a
//     few microseconds; here CTAs are persistent (two per SM), 4 producer warps run ahead through the (tile, offset,
//     slab) sequence on one stage ring that never drains between tiles, the MMA warp alternates between two TMEM
//     accumulators and 4 epilogue warps drain tile i while tile i+1 is being multiplied;
//   * tiles are MASK-SORTED (spvd_kmap_sort): rows that share their set of active kernel offsets are grouped, so an
//     output-stationary tile issues few zero rows even on sparse levels (stride 1: 0.08 -> 0.36 of the issued rows are
//     real pairs, stride 2: 0.28 -> 0.43) -- no pair-major RED scatter, no zero fill of the output, plain stores;
//   * the epilogue applies what followed the convolution as separate kernels: bias, residual, FiLM
//     ((1 + scale[b]) * x + shift[b], models/modules/graph.py:35-41), the next BatchNorm (eval: affine) + SiLU, and writes
//     the fp32 result and / or the f16 operand of the next convolution;
//   * a second operand source extends the K loop: the 1x1 shortcut (idconv, models/ddpm_unet_attn.py:61) of a residual
//     block runs as extra slabs of its second 3^3 convolution instead of as a launch of its own.
//   * map rows reach the producers through a ring of 512-byte segments that thread 0 bulk-copies (cp.async.bulk) eight
//     (tile, offset) steps ahead, and the offset masks of a CTA's tiles are staged once: no role has a dependent global
//     load on its critical path.
// Measured and rejected: gathering the rows with the TMA (cp.async.bulk.tensor ... tile::gather4, 512 B per instruction)
// is correct but about twice as slow as the per-thread cp.async gather on every SPVD layer (profiles/README.md).
#include <cuda_fp16.h>

#include "conv_umma_impl.cuh"

namespace spvd {

constexpr int kSpThreads = 32 * 9;   // warps 0-3 producers, 4 MMA issuer, 5-8 epilogue
constexpr int kSpMaxStages = 8;
constexpr int kSpIdxSlots = 8;       // ring of map-row segments (512 B each) prefetched ahead of the gathers
constexpr int kSpMaskCache = 256;    // offset masks of this CTA's tiles, staged once
constexpr int kSpTailBytes = 256 /*barriers*/ + kSpIdxSlots * 512 + kSpMaskCache * 4;

template <int BN>
struct SpCfg {
  static constexpr int kBTileBytes = BN * 128;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kBufs = BN <= 128 ? 2 : 1;   // two CTAs per SM share the 512 TMEM columns
  static constexpr int kCols = kBufs * BN;
  static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : 256;
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(kCols <= 256, "two CTAs per SM");
};

__device__ __forceinline__ float sp_silu(float v) { return v / (1.f + __expf(-v)); }

template <int BN>
__global__ void __launch_bounds__(kSpThreads, 2) k_conv_sp_f16(const __grid_constant__ spvd_conv_sp_args_t a, int S) {
  using Cfg = SpCfg<BN>;
  constexpr int NB = Cfg::kBufs;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_full = base + S * Cfg::kStageBytes;          // full[8]
  const uint32_t bar_empty = bar_full + 8 * kSpMaxStages;         // empty[8]
  const uint32_t bar_accfull = bar_empty + 8 * kSpMaxStages;      // [2]
  const uint32_t bar_accempty = bar_accfull + 16;                 // [2]
  const uint32_t bar_idx = bar_accempty + 16;                     // [kSpIdxSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * Cfg::kStageBytes + 16 * kSpMaxStages + 32 + 8 * kSpIdxSlots);
  int* s_idx = reinterpret_cast<int*>(smem + S * Cfg::kStageBytes + 256);                       // [kSpIdxSlots][128]
  unsigned* s_mask = reinterpret_cast<unsigned*>(smem + S * Cfg::kStageBytes + 256 + kSpIdxSlots * 512);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;
  const int nchunks = (a.cin + 63) / 64;
  const int nchunks2 = a.in2 ? (a.cin2 + 63) / 64 : 0;
  const int n_tiles = (a.n_out + BM - 1) / BM;
  const unsigned kmask = a.kvol >= 32 ? 0xffffffffu : ((1u << a.kvol) - 1u);

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, kProducerThreads + 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accfull + 8 * b, 1);
      mbar_init(bar_accempty + 8 * b, 128);
    }
    for (int i = 0; i < kSpIdxSlots; ++i) mbar_init(bar_idx + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // offset masks of this CTA's tiles (every role needs them per tile: no dependent global load on any critical path)
  if (a.map_t && a.tile_mask)
    for (int i = tid; i < kSpMaskCache; i += kSpThreads) {
      const int tile = blockIdx.x + i * gridDim.x;
      s_mask[i] = tile < n_tiles ? (a.tile_mask[tile] & kmask) : 0u;
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
  pdl_wait();   // geometry and weights were complete long before; activations follow

  auto tile_mask_of = [&](int tile, int i) -> unsigned {           // tile = blockIdx.x + i * gridDim.x
    if (!a.map_t) return 1u;                                       // identity map: a single "offset"
    if (!a.tile_mask) return kmask;
    return i < kSpMaskCache ? s_mask[i] : (a.tile_mask[tile] & kmask);
  };

  if (warp < 4) {
    // ------------------------------------------------------------------ producers (thread t: 16-byte piece t & 7 of rows (t >> 3) + 16 i)
    const int chunk16 = tid & 7, r0 = tid >> 3;
    const uint8_t* w_b = reinterpret_cast<const uint8_t*>(a.w);
    int li = 0;
    auto fill = [&](const uint8_t* src, uint32_t row_bytes, const int* rows, const uint8_t* wsrc, int c) {
      const int stage = li % S;
      const uint32_t parity = (uint32_t)(li / S) & 1u;
      const uint32_t a_s = base + stage * Cfg::kStageBytes;
      mbar_wait(bar_empty + 8 * stage, parity ^ 1u);
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_full + 8 * stage, Cfg::kBTileBytes);
        bulk_g2s(a_s + kATileBytes, wsrc, Cfg::kBTileBytes, bar_full + 8 * stage);
      }
      __syncwarp();
      const uint32_t col_b = (uint32_t)c * 128u + (uint32_t)chunk16 * 16u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + 16 * i;
        const uint32_t dst = a_s + r * 128 + ((chunk16 ^ (r & 7)) << 4);
        const bool ok = rows[i] >= 0 && col_b < row_bytes;
        cp_async16(dst, ok ? src + (size_t)rows[i] * row_bytes + col_b : src, ok ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(bar_full + 8 * stage);
      ++li;
    };
    // map rows travel through a ring of 512-byte segments, bulk-copied kSpIdxSlots (tile, offset) steps ahead by thread 0
    int pf_i = 0, pf_tile = blockIdx.x, pf_q = 0;      // prefetch cursor (thread 0): i-th tile of this CTA, remaining offsets
    unsigned pf_rem = (a.map_t && pf_tile < n_tiles) ? tile_mask_of(pf_tile, 0) : 0u;
    auto prefetch_one = [&]() {                         // thread 0 only
      while (pf_tile < n_tiles && pf_rem == 0u) {
        pf_tile += gridDim.x;
        ++pf_i;
        pf_rem = pf_tile < n_tiles ? tile_mask_of(pf_tile, pf_i) : 0u;
      }
      if (pf_tile >= n_tiles) return;
      const int k = __ffs(pf_rem) - 1;
      pf_rem &= pf_rem - 1;
      const int slot = pf_q % kSpIdxSlots;
      fence_proxy_async();   // the generic-proxy reads of this segment (ordered by the barrier) precede the async-proxy refill
      mbar_arrive_expect_tx(bar_idx + 8 * slot, 512u);
      bulk_g2s(smem_u32(s_idx + slot * BM), a.map_t + (size_t)k * a.ld + (size_t)pf_tile * BM, 512u, bar_idx + 8 * slot);
      ++pf_q;
    };
    if (a.map_t && tid == 0)
      for (int i = 0; i < kSpIdxSlots; ++i) prefetch_one();
    __syncwarp();
    int q = 0, ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      int rows[8];
      if (a.map_t) {
        unsigned rem = tile_mask_of(tile, ti);
        while (rem) {
          const int k = __ffs(rem) - 1;
          rem &= rem - 1;
          const int slot = q % kSpIdxSlots;
          mbar_wait(bar_idx + 8 * slot, (uint32_t)(q / kSpIdxSlots) & 1u);
          // (volatile asm: these loads must stay in front of the barrier below, which releases the segment for refill --
          //  as plain loads they were sunk behind it and a handful of rows per launch gathered a refilled segment)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rows[i]) : "r"(smem_u32(s_idx + slot * BM + r0 + 16 * i)) : "memory");
          ++q;
          // warp 0 diverges around thread 0's bulk copies: reconverge first (an aligned barrier reached by a split warp is
          // undefined) and use the unaligned form
          __syncwarp();
          asm volatile("barrier.sync 1, 128;" ::: "memory");
          if (tid == 0) prefetch_one();
          __syncwarp();
          for (int c = 0; c < nchunks; ++c)
            fill(reinterpret_cast<const uint8_t*>(a.in), (uint32_t)a.cin * 2u, rows,
                 w_b + ((size_t)(k * nchunks + c) * a.cout + n0) * 128, c);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int o = tile * BM + r0 + 16 * i;
          rows[i] = o < a.n_out ? o : -1;
        }
        for (int c = 0; c < nchunks; ++c)
          fill(reinterpret_cast<const uint8_t*>(a.in), (uint32_t)a.cin * 2u, rows, w_b + ((size_t)c * a.cout + n0) * 128, c);
      }
      if (a.in2) {   // K extension: rows of the second source are the tile's own output rows
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int o = tile * BM + r0 + 16 * i;
          rows[i] = o < a.n_out ? (a.out_rows ? __ldg(a.out_rows + o) : o) : -1;
        }
        for (int c = 0; c < nchunks2; ++c)
          fill(reinterpret_cast<const uint8_t*>(a.in2), (uint32_t)a.cin2 * 2u, rows,
               reinterpret_cast<const uint8_t*>(a.w2) + ((size_t)c * a.cout + n0) * 128, c);
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    int li = 0, t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int iters = __popc(tile_mask_of(tile, t)) * nchunks + nchunks2;
      const int buf = t % NB;
      mbar_wait(bar_accempty + 8 * buf, (((uint32_t)(t / NB)) & 1u) ^ 1u);   // the epilogue drained this accumulator
      tc_fence_after();
      if (iters == 0) {
        if (lane == 0) mbar_arrive(bar_accfull + 8 * buf);
        __syncwarp();
        continue;
      }
      for (int it = 0; it < iters; ++it, ++li) {
        const int stage = li % S;
        mbar_wait(bar_full + 8 * stage, (uint32_t)(li / S) & 1u);
        fence_proxy_async();     // cp.async wrote the A tile through the generic proxy
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_s = base + stage * Cfg::kStageBytes;
          const uint64_t adesc = make_desc(a_s), bdesc = make_desc(a_s + kATileBytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_d + (uint32_t)(buf * BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), Cfg::kIdesc,
                     (it > 0 || kk > 0) ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);
          if (it == iters - 1) umma_commit(bar_accfull + 8 * buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    const bool has_aff = a.aff_scale != nullptr;
    int t = 0;
    auto out_row_of = [&](int tile) -> int {
      const int o = tile * BM + q * 32 + lane;
      return (tile < n_tiles && o < a.n_out) ? (a.out_rows ? __ldg(a.out_rows + o) : o) : -1;
    };
    int row_next = out_row_of(blockIdx.x);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int iters = __popc(tile_mask_of(tile, t)) * nchunks + nchunks2;
      const int buf = t % NB;
      const int row = row_next;
      row_next = out_row_of(tile + gridDim.x);      // the next tile's output row travels under this tile's epilogue
      const float* film_row = nullptr;
      if (a.film && row >= 0) film_row = a.film + (size_t)__ldg(a.coords + 4 * (size_t)row) * a.film_ld;
      mbar_wait(bar_accfull + 8 * buf, ((uint32_t)(t / NB)) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t r[32];
        if (iters > 0) {
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
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (row < 0) continue;
        const int c0 = n0 + cb * 32;
        const size_t off = (size_t)row * a.cout + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // four channels at a time
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (a.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c0) + j);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          if (a.residual) {
            const float4 b = *(reinterpret_cast<const float4*>(a.residual + off) + j);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          if (film_row) {
            const float4 sc = *(reinterpret_cast<const float4*>(film_row + c0) + j);
            const float4 sh = *(reinterpret_cast<const float4*>(film_row + a.cout + c0) + j);
            v.x = fmaf(1.f + sc.x, v.x, sh.x); v.y = fmaf(1.f + sc.y, v.y, sh.y);
            v.z = fmaf(1.f + sc.z, v.z, sh.z); v.w = fmaf(1.f + sc.w, v.w, sh.w);
          }
          if (a.out32) *(reinterpret_cast<float4*>(a.out32 + off) + j) = v;
          if (a.out16) {
            if (has_aff) {
              const float4 s = __ldg(reinterpret_cast<const float4*>(a.aff_scale + c0) + j);
              const float4 h = __ldg(reinterpret_cast<const float4*>(a.aff_shift + c0) + j);
              v.x = fmaf(v.x, s.x, h.x); v.y = fmaf(v.y, s.y, h.y); v.z = fmaf(v.z, s.z, h.z); v.w = fmaf(v.w, s.w, h.w);
            }
            if (a.act & 1) {
              v.x = sp_silu(v.x); v.y = sp_silu(v.y); v.z = sp_silu(v.z); v.w = sp_silu(v.w);
            }
            r[2 * j] = pack_half2_sat(v.x, v.y);        // (slots 2j, 2j+1 <= 4j: already consumed)
            r[2 * j + 1] = pack_half2_sat(v.z, v.w);
          }
        }
        if (a.out16) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(a.out16) + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
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

static int sp_num_sms() {
  static int n_sm[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (n_sm[dev] == 0 && cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n_sm[dev] = 148;
  return n_sm[dev];
}

template <int BN>
static int launch_sp(const spvd_conv_sp_args_t& a, int stages, cudaStream_t s) {
  using Cfg = SpCfg<BN>;
  auto kern = k_conv_sp_f16<BN>;
  constexpr int kMaxSmem = 200 * 1024;
  static bool configured[64] = {};
  int dev = 0;
  SPVD_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    SPVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured[dev] = true;
  }
  const int n_sm = sp_num_sms();
  const int n_tiles = cdiv(a.n_out, BM), slices = a.cout / BN;
  int ctas = (2 * n_sm) / slices;
  if (ctas < 1) ctas = 1;
  if (ctas > n_tiles) ctas = n_tiles;
  // few CTAs (coarse levels): one per SM with the whole shared memory as stage ring; otherwise two per SM
  const int budget = (ctas * slices <= n_sm ? kMaxSmem : 112 * 1024) - 1024 - kSpTailBytes;
  int S = budget / Cfg::kStageBytes;
  if (S > kSpMaxStages) S = kSpMaxStages;
  if (stages > 0 && stages < S) S = stages;
  SPVD_CHECK_ARG(S >= 2, "spvd_conv_sp_f16: no room for two pipeline stages at BN = %d", BN);
  const size_t smem = (size_t)S * Cfg::kStageBytes + 1024 + kSpTailBytes;
  SPVD_CUDA(launch_pdl(kern, dim3(ctas, slices), dim3(kSpThreads), smem, s, a, S));
  return SPVD_OK;
}

}  // namespace spvd

using namespace spvd;

extern "C" int spvd_conv_sp_f16(const spvd_conv_sp_args_t* args, int bn, int stages, spvd_stream_t stream) {
  SPVD_CHECK_ARG(args, "spvd_conv_sp_f16: NULL arguments");
  const spvd_conv_sp_args_t& a = *args;
  SPVD_CHECK_ARG(a.in && a.w && (a.out32 || a.out16) && a.n_out >= 0, "spvd_conv_sp_f16: bad arguments");
  SPVD_CHECK_ARG(spvd_conv_umma_supported(a.kvol, a.cin, a.cout) && a.cin % 8 == 0, "spvd_conv_sp_f16: unsupported shape %dx%dx%d",
                 a.kvol, a.cin, a.cout);
  SPVD_CHECK_ARG(a.map_t || (a.kvol == 1 && !a.out_rows), "spvd_conv_sp_f16: a NULL map needs kvol == 1 and no row permutation");
  SPVD_CHECK_ARG(!a.map_t || (a.ld % 128 == 0 && a.ld >= a.n_out), "spvd_conv_sp_f16: ld %d must be a multiple of 128 >= n_out", a.ld);
  SPVD_CHECK_ARG(!a.in2 || (a.w2 && a.cin2 > 0 && a.cin2 % 8 == 0), "spvd_conv_sp_f16: the second source needs weights and cin2 %% 8 == 0");
  SPVD_CHECK_ARG(!a.film || (a.coords && a.film_ld >= 2 * a.cout && a.film_ld % 4 == 0 && ((size_t)a.film & 15) == 0),
                 "spvd_conv_sp_f16: FiLM needs coords and a table of 2 * cout columns (16-byte aligned)");
  SPVD_CHECK_ARG((a.aff_scale == nullptr) == (a.aff_shift == nullptr), "spvd_conv_sp_f16: affine needs scale and shift");
  SPVD_CHECK_ARG((((size_t)a.in | (size_t)a.w | (size_t)a.in2 | (size_t)a.w2 | (size_t)a.bias | (size_t)a.residual | (size_t)a.out32 |
                   (size_t)a.out16 | (size_t)a.aff_scale | (size_t)a.aff_shift) & 15) == 0,
                 "spvd_conv_sp_f16: pointers must be 16-byte aligned");
  SPVD_CHECK_ARG(a.residual == nullptr || a.residual != a.out32, "spvd_conv_sp_f16: residual must not alias out32");
  if (a.n_out == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (bn <= 0) {
    // widest slice that divides cout (the A tile is gathered once per slice); narrower while the grid would leave SMs idle
    const int widths[] = {256, 192, 128, 96, 64, 32};
    const int n_tiles = cdiv(a.n_out, BM);
    for (int w : widths) {
      if (a.cout % w) continue;
      if (bn <= 0 || (n_tiles * (a.cout / bn) < sp_num_sms() && w >= 64)) bn = w;
    }
  }
  SPVD_CHECK_ARG(bn > 0 && a.cout % bn == 0, "spvd_conv_sp_f16: slice width %d does not divide cout %d", bn, a.cout);
  switch (bn) {
    case 32: return launch_sp<32>(a, stages, s);
    case 64: return launch_sp<64>(a, stages, s);
    case 96: return launch_sp<96>(a, stages, s);
    case 128: return launch_sp<128>(a, stages, s);
    case 192: return launch_sp<192>(a, stages, s);
    case 256: return launch_sp<256>(a, stages, s);
  }
  set_error("spvd_conv_sp_f16: slice width %d is not instantiated", bn);
  return SPVD_ERR_INVALID;
}
