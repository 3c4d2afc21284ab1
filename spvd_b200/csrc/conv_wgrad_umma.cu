// Weight gradient of the sparse convolution on the tensor cores (tcgen05, TF32 operands, fp32 accumulate):
//   gw[k][ci][co] = sum over the (input row i, output row o) pairs of offset k of  x[i][ci] * gout[o][co]
// Per offset this is a GEMM whose reduction dimension is the pair list, so both operands are gathered ROWS whose
// contiguous direction (channels) is the M / N direction of the MMA: they are fed as MN-major operands.  For 32-bit
// (TF32) elements the only MN-major shared-memory layout the tensor core accepts is SWIZZLE_128B_BASE32B: rows of 128
// bytes (32 channels of one pair), atoms of 4 rows (512 bytes), the four 32-byte segments of a row XOR-permuted by
// (row & 3); 32-channel groups LBO bytes apart, 4-pair groups 512 bytes apart.
// A CTA owns a chunk of consecutive pair tiles (spvd_kmap_pairs: tiles of 128 pairs, sorted by offset), a 128-channel
// block of cin and a BN-wide slice of cout; every run of tiles with the same offset accumulates in TMEM and is flushed
// into gw[k] with vector RED.
#include "common.cuh"

namespace spvd {
namespace wg {

constexpr int kKT = 32;               // pairs per pipeline stage (4 MMAs of K = 8)
constexpr int kMBlock = 128;          // cin channels per CTA (MMA M)
constexpr int kThreads = 160;         // 4 gather/epilogue warps + 1 MMA warp
constexpr int kTilesPerCta = 16;      // pair tiles (of 128 pairs) per CTA
constexpr int kSubBytes = kKT * 128;  // one 32-channel group of a stage: 32 rows x 128 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// MN-major SWIZZLE_128B_BASE32B operand: 32-channel groups `lbo` bytes apart, 4-row (4 pairs) groups 512 bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}

template <int BN>
struct Cfg {
  static constexpr int kStageBytes = (kMBlock / 32 + BN / 32) * kSubBytes;
  static constexpr int kStages = (4 * kStageBytes <= 160 * 1024) ? 4 : (3 * kStageBytes <= 160 * 1024) ? 3 : 2;
  static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kPairBytes = 2 * kTilesPerCta * 128 * 4 + kTilesPerCta * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kPairBytes;
  // D = f32, A = B = tf32, BOTH MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
  static constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                     ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kMBlock >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(kThreads) k_conv_wgrad_umma(const float* __restrict__ x, const float* __restrict__ g,
                                                              const int* __restrict__ pair_in,
                                                              const int* __restrict__ pair_out,
                                                              const int* __restrict__ tile_k, int n_tiles, int n_identity,
                                                              int cin, int cout, float* __restrict__ gw) {
  using C = Cfg<BN>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + kStages * C::kStageBytes;   // full[kStages], empty[kStages], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * C::kStageBytes + 8 * (2 * kStages + 1));
  int* s_in = reinterpret_cast<int*>(smem + kStages * C::kStageBytes + 256);
  int* s_out = s_in + kTilesPerCta * 128;
  int* s_tk = s_out + kTilesPerCta * 128;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * kTilesPerCta;
  const int nt = min(kTilesPerCta, n_tiles - t0);
  const int ci0 = blockIdx.y * kMBlock, n0 = blockIdx.z * BN;
  if (nt <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_base + 8 * s, 128);
      mbar_init(bar_base + 8 * (kStages + s), 1);
    }
    mbar_init(bar_base + 8 * (2 * kStages), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the chunk's pair rows (identity pairs when there is no map: kernel size 1)
  for (int e = tid; e < nt * 128; e += kThreads) {
    const long long p = (long long)t0 * 128 + e;
    int a, b;
    if (pair_in) {
      a = pair_in[p];
      b = pair_out[p];
    } else {
      a = b = p < n_identity ? (int)p : -1;
    }
    s_in[e] = a;
    s_out[e] = b;
  }
  for (int e = tid; e < nt; e += kThreads) s_tk[e] = tile_k ? tile_k[t0 + e] : 0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  int it = 0;          // pipeline iteration counter across runs (same sequence in every role)
  int run = 0;
  for (int ta = 0; ta < nt; ++run) {
    int tb = ta + 1;
    const int k = s_tk[ta];
    while (tb < nt && s_tk[tb] == k) ++tb;
    const int iters = (tb - ta) * (128 / kKT);
    if (warp < 4) {
      // ---------------------------------------------------------------- gather x rows (A) and gout rows (B)
      const int chunk16 = tid & 7, r0 = tid >> 3;    // rows r0 and r0 + 16 of the 32-pair slab
      for (int j = 0; j < iters; ++j, ++it) {
        const int stage = it % kStages;
        const uint32_t parity = (uint32_t)(it / kStages) & 1u;
        mbar_wait(bar_base + 8 * (kStages + stage), parity ^ 1u);
        const uint32_t a_s = base + stage * C::kStageBytes;
        const int p0 = ta * 128 + j * kKT;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = r0 + 16 * h;
          const int ri = s_in[p0 + r], ro = s_out[p0 + r];
          const uint32_t off = r * 128 + ((((chunk16 >> 1) ^ (r & 3)) << 5) | ((chunk16 & 1) << 4));
#pragma unroll
          for (int sub = 0; sub < kMBlock / 32; ++sub) {
            const int c = ci0 + sub * 32 + chunk16 * 4;
            const bool ok = ri >= 0 && c < cin;
            cp_async16(a_s + sub * kSubBytes + off, ok ? x + (size_t)ri * cin + c : x, ok ? 16u : 0u);
          }
#pragma unroll
          for (int sub = 0; sub < BN / 32; ++sub) {
            const bool ok = ro >= 0;
            cp_async16(a_s + (kMBlock / 32 + sub) * kSubBytes + off,
                       ok ? g + (size_t)ro * cout + n0 + sub * 32 + chunk16 * 4 : g, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(bar_base + 8 * stage);
      }
      // ---------------------------------------------------------------- flush the run into gw[k]
      mbar_wait(bar_base + 8 * (2 * kStages), (uint32_t)run & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int ci = ci0 + warp * 32 + lane;
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ci < cin) {
          float* dst = gw + ((size_t)k * cin + ci) * cout + n0 + cb * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            red_add_v4(dst + 4 * q, make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
        }
      }
    } else if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      for (int j = 0; j < iters; ++j, ++it) {
        const int stage = it % kStages;
        const uint32_t parity = (uint32_t)(it / kStages) & 1u;
        mbar_wait(bar_base + 8 * stage, parity);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_s = base + stage * C::kStageBytes;
        const uint64_t adesc = make_desc_mn(a_s, kSubBytes);
        const uint64_t bdesc = make_desc_mn(a_s + (kMBlock / 32) * kSubBytes, kSubBytes);
#pragma unroll
        for (int kk = 0; kk < kKT / 8; ++kk)   // 8 pairs per MMA = one 1024-byte row group
          umma_tf32(tmem_d, adesc + (uint64_t)(kk * 64), bdesc + (uint64_t)(kk * 64), C::kIdesc, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(bar_base + 8 * (kStages + stage));
        if (j == iters - 1) umma_commit(bar_base + 8 * (2 * kStages));
      }
    } else {
      it += iters;
    }
    // the accumulator is reused by the next run: every epilogue warp must have read it
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    ta = tb;
  }
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

template <int BN>
static int launch(const float* x, const float* g, const int32_t* pin, const int32_t* pout, const int32_t* tk, int n_tiles,
                  int n_identity, int cin, int cout, float* gw, cudaStream_t s) {
  using C = Cfg<BN>;
  static bool configured_dev[64] = {};   // the attribute is per device: one flag per device, not per process
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(k_conv_wgrad_umma<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  dim3 grid(cdiv(n_tiles, kTilesPerCta), cdiv(cin, kMBlock), cout / BN);
  k_conv_wgrad_umma<BN><<<grid, kThreads, C::kSmemBytes, s>>>(x, g, pin, pout, tk, n_tiles, n_identity, cin, cout, gw);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

}  // namespace wg
}  // namespace spvd

using namespace spvd;

extern "C" int spvd_conv_umma_supported(int kvol, int cin, int cout);

extern "C" int spvd_conv_wgrad_umma(const float* x, const float* gout, const int32_t* pair_in, const int32_t* pair_out,
                                    const int32_t* tile_k, int n_pair_tiles, int n_identity, int kvol, int cin, int cout,
                                    float* gw, spvd_stream_t stream) {
  SPVD_CHECK_ARG(x && gout && gw && kvol >= 1 && n_pair_tiles >= 0, "spvd_conv_wgrad_umma: bad arguments");
  SPVD_CHECK_ARG(spvd_conv_umma_supported(kvol, cin, cout), "spvd_conv_wgrad_umma: unsupported shape %dx%dx%d", kvol, cin, cout);
  SPVD_CHECK_ARG((pair_in && pair_out && tile_k) || (!pair_in && !pair_out && kvol == 1),
                 "spvd_conv_wgrad_umma: pair lists are required unless kvol == 1 (identity pairs)");
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(gw, 0, (size_t)kvol * cin * cout * 4, s));
  if (!pair_in) n_pair_tiles = cdiv(n_identity, 128);
  if (n_pair_tiles == 0) return SPVD_OK;
  const int widths[] = {256, 192, 128, 96, 64, 32};
  int bn = 32;
  for (int w : widths)
    if (cout % w == 0) {
      bn = w;
      break;
    }
  switch (bn) {
    case 32: return wg::launch<32>(x, gout, pair_in, pair_out, tile_k, n_pair_tiles, n_identity, cin, cout, gw, s);
    case 64: return wg::launch<64>(x, gout, pair_in, pair_out, tile_k, n_pair_tiles, n_identity, cin, cout, gw, s);
    case 96: return wg::launch<96>(x, gout, pair_in, pair_out, tile_k, n_pair_tiles, n_identity, cin, cout, gw, s);
    case 128: return wg::launch<128>(x, gout, pair_in, pair_out, tile_k, n_pair_tiles, n_identity, cin, cout, gw, s);
    case 192: return wg::launch<192>(x, gout, pair_in, pair_out, tile_k, n_pair_tiles, n_identity, cin, cout, gw, s);
    case 256: return wg::launch<256>(x, gout, pair_in, pair_out, tile_k, n_pair_tiles, n_identity, cin, cout, gw, s);
  }
  return SPVD_ERR_INVALID;
}
