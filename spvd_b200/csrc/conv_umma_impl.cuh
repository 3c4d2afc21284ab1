// Sparse convolution on the 5th-generation tensor cores (sm_100a tcgen05 + TMEM), C1-C3.
//
//   out[o, :] = sum_k  in[map_t[k][o], :] @ W[k]  (+ bias)          fp32 in/out, TF32 operands,
//                                                                    fp32 accumulation in TMEM
// Output-stationary implicit GEMM: one CTA owns a 128-row output tile x BN output channels and
// walks the (active kernel offset, 32-channel slab) pairs of that tile.  Per step
//   * 4 producer warps gather the 128 input rows of the slab (128 B each, 16-byte loads, rows with
//     map -1 are zero-filled without touching memory), round to TF32 (cvt.rna) and store them into
//     the canonical K-major SWIZZLE_128B shared-memory tile;
//   * one thread starts a bulk async copy (cp.async.bulk, SASS UBLKCP) of the matching BN x 32
//     weight tile, which spvd_conv_pack_weights laid out in global memory as the exact
//     shared-memory image (K-major, 128-byte swizzle, TF32-rounded), completing on the same mbarrier;
//   * one thread of the MMA warp issues 4 x tcgen05.mma (M=128, N=BN, K=8) into the TMEM accumulator
//     and commits the stage back to the producers.
// Offsets with no active row in the tile are skipped through the per-tile bit mask.  The epilogue
// reads the accumulator with tcgen05.ld (32 lanes x 32 columns per warp), adds the bias and writes
// 128-byte row segments.
#pragma once
#include "common.cuh"

namespace spvd {

// tools/conv_trace.cu compiles this header with -DSPVD_TRACE: per-CTA clock marks of the pipeline phases
#ifdef SPVD_TRACE
__device__ unsigned long long* g_spvd_trace = nullptr;     // [ctas][32]
#define SPVD_MARK(slot)                                                                                          \
  do {                                                                                                           \
    if (g_spvd_trace) g_spvd_trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = clock64(); \
  } while (0)
#define SPVD_ACC(slot, v)                                                                                        \
  do {                                                                                                           \
    if (g_spvd_trace) g_spvd_trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] += (v); \
  } while (0)
#define SPVD_CLK() clock64()
#define SPVD_GMARK(slot)                                                                                         \
  do {                                                                                                           \
    unsigned long long t_;                                                                                       \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                       \
    if (g_spvd_trace) g_spvd_trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = t_; \
  } while (0)
#define SPVD_SMID(slot)                                                                                          \
  do {                                                                                                           \
    unsigned s_;                                                                                                 \
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s_));                                                              \
    if (g_spvd_trace) g_spvd_trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = s_; \
  } while (0)
#else
#define SPVD_GMARK(slot) do {} while (0)
#define SPVD_SMID(slot) do {} while (0)
#define SPVD_MARK(slot) do {} while (0)
#define SPVD_ACC(slot, v) do {} while (0)
#define SPVD_CLK() 0ll
#endif

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int kStages = 4;
constexpr int kProducerThreads = 128;
constexpr int kUmmaThreads = kProducerThreads + 32;
constexpr int kATileBytes = BM * BK * 4;  // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32; // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;           // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

template <int BN, int MT, bool kPairs = false, bool kLite = false>
struct UmmaCfg {
  static constexpr int kCols = BN * MT;
  static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
  static constexpr int kBTileBytes = BN * BK * 4;
  static constexpr int kStageBytes = MT * kATileBytes + kBTileBytes;
  // pair-major tiles run only cin/32 iterations: keep the footprint small so that several CTAs share an SM and
  // one CTA's prologue / RED epilogue overlaps another's main loop
  static constexpr int kLiteStages = (3 * kStageBytes + MT * 27 * BM * 4 + 1280 <= 113 * 1024) ? 3 : 2;   // still two CTAs per SM
  static constexpr int kStages = kPairs ? 2 : kLite ? kLiteStages : (4 * kStageBytes <= 200 * 1024) ? 4 : (3 * kStageBytes <= 200 * 1024) ? 3 : 2;
  static constexpr int kRowsBytes = kPairs ? 2 * BM * 4 : MT * 27 * BM * 4;  // staged row indices
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kRowsBytes;
  // idesc: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
  static constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                     ((uint32_t)(BM >> 4) << 24);
  // the same with A = B = f16 (format 0): kind::f16, K = 16 per instruction
  static constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(kCols <= 512, "accumulators exceed TMEM");
};

__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// One CTA owns MT consecutive 128-row output tiles x BN output channels: every weight tile fetched from
// L2 feeds MT tensor-core tiles (the weight stream, not the gathered rows, dominates L2 traffic).
// kAsyncA: the input rows are already TF32-representable (the producing kernel rounded them), so the
// gather is a zero-register cp.async (LDGSTS) with up to kStages slabs in flight per CTA.  Otherwise the
// rows go through registers (cvt.rna.tf32) with the next slab's loads issued before the current stores
// (MT = 1 only).  gridDim.z > 1 = split over the (offset, slab) iterations; partial tiles are accumulated
// with vector RED into a zero-initialised output (bias / residual added by split 0).
//
// kPairs (pair-major, for sparse levels): a CTA owns one tile of 128 (input row, output row) pairs of ONE
// kernel offset (spvd_kmap_pairs): map_t = pair_in, tile_mask = tile_k (offset of each tile), pair_out =
// output rows; the K loop runs over the channel slabs only and the epilogue scatters the 128 product rows with
// vector RED into the zero-initialised output (bias / residual added by the centre-offset tiles, which cover
// every output row exactly once).  Issued tensor-core rows = real pairs, instead of 128 x (active offsets).
//
// kCluster (2 or 4): the CTAs of a cluster own consecutive output tiles and walk the same (offset, slab) sequence; each
// loads 1/kCluster of every weight tile and multicasts it into all of them (the weight stream is ~70% of the L2
// traffic of the dense levels), a stage is recycled when every CTA of the cluster has consumed it.
template <int BN, int MT, bool kAsyncA, bool kPairs = false, bool kLite = false, int kCluster = 1, bool kF16 = false>
__global__ void __launch_bounds__(kUmmaThreads) k_conv_fwd_umma(const float* __restrict__ in,
                                                               const float* __restrict__ w_packed,
                                                               const int* __restrict__ map_t, int ld,
                                                               const unsigned* __restrict__ tile_mask, int n_out,
                                                               int kvol, int cin, int cout,
                                                               const float* __restrict__ bias,
                                                               const float* __restrict__ residual,
                                                               float* __restrict__ out,
                                                               const int* __restrict__ pair_out, int center_k) {
  using Cfg = UmmaCfg<BN, MT, kPairs, kLite>;
  constexpr int kStages = Cfg::kStages;
  static_assert(kAsyncA || MT == 1, "the register gather path is instantiated for MT = 1 only");
  static_assert(!kPairs || MT == 1, "pair-major tiles: MT = 1");
  static_assert(kCluster == 1 || (MT == 1 && !kPairs && (BN / kCluster) % 8 == 0), "cluster multicast: MT = 1, slices of 8 rows");
  static_assert(!kF16 || kAsyncA || (MT == 1 && kCluster == 1), "f16 operands converted on the fly: one tile per CTA, no cluster");
  pdl_trigger();
  if (threadIdx.x == 0) {
    SPVD_MARK(0);
    SPVD_GMARK(10);
    SPVD_SMID(12);
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + kStages * Cfg::kStageBytes;  // full[kStages], empty[kStages], accum, tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * Cfg::kStageBytes + 8 * (2 * kStages + 1));
  int* active_k = reinterpret_cast<int*>(smem + kStages * Cfg::kStageBytes + 8 * (2 * kStages + 1) + 16);
  int* s_rows = reinterpret_cast<int*>(smem + kStages * Cfg::kStageBytes + 256);   // [offset slot][MT*128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * (BM * MT), n0 = blockIdx.y * BN;
  // a slab is one 128-byte row segment: 32 TF32 or 64 f16 input channels (f16: the last slab may be half empty)
  const int nchunks = kF16 ? (cin + 63) / 64 : cin / BK;
  const int n_tiles = (n_out + BM - 1) / BM;
  const uint8_t* in_b = reinterpret_cast<const uint8_t*>(in);
  const uint8_t* w_b = reinterpret_cast<const uint8_t*>(w_packed);
  const uint32_t row_bytes = (uint32_t)cin * ((kF16 && kAsyncA) ? 2u : 4u);

  unsigned mask = kvol >= 32 ? 0xffffffffu : ((1u << kvol) - 1u);
  if constexpr (kPairs) {
    mask = 1u << (reinterpret_cast<const int*>(tile_mask)[blockIdx.x] & 31);   // the tile's single offset
  } else if (tile_mask) {
    unsigned u = 0;
    const int t0 = kCluster > 1 ? ((int)blockIdx.x / kCluster) * kCluster : (int)blockIdx.x * MT;
#pragma unroll
    for (int t = 0; t < (kCluster > 1 ? kCluster : MT); ++t)
      if (t0 + t < n_tiles) u |= tile_mask[t0 + t];
    mask &= u;
  }
  const int total_iters = __popc(mask) * nchunks;
  const int per_split = (total_iters + (int)gridDim.z - 1) / (int)gridDim.z;
  const int it_begin = min(total_iters, (int)blockIdx.z * per_split);
  const int it_end = min(total_iters, it_begin + per_split);
  const int num_iters = it_end - it_begin;
  const bool split = kPairs || gridDim.z > 1;
  if (split && num_iters == 0) return;  // nothing to add (uniform across the CTA, before any barrier/alloc)

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_base + 8 * s, kProducerThreads + 1);
      mbar_init(bar_base + 8 * (kStages + s), kCluster);   // released by the MMA commit of every CTA of the cluster
    }
    mbar_init(bar_base + 8 * (2 * kStages), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int j = 0;
    for (int k = 0; k < kvol && k < 32; ++k)
      if ((mask >> k) & 1u) active_k[j++] = k;
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) cluster_sync_all();   // peers' barriers are initialised before anything remote arrives
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  if (threadIdx.x == 0) SPVD_MARK(1);

  // stage the input-row indices of every offset this CTA will visit (one coalesced pass; the per-iteration
  // critical path then has no dependent global load in front of the gather)
  const int ka_begin = num_iters > 0 ? it_begin / nchunks : 0;
  const int ka_end = num_iters > 0 ? (it_end - 1) / nchunks + 1 : 0;
  if constexpr (kPairs) {
    if (tid < BM) {
      s_rows[tid] = __ldg(map_t + (size_t)blockIdx.x * BM + tid);            // input rows of the pair tile
      s_rows[BM + tid] = __ldg(pair_out + (size_t)blockIdx.x * BM + tid);    // output rows
    }
  } else {
    if (map_t) {
      // 16-byte asynchronous copies (512 B of one offset's map row per 32 lanes), all in flight at once: the dependent
      // __ldg loop this replaces cost ~20k cycles per CTA with 27 active offsets (tools/_conv_trace.py)
      constexpr int kChunks = MT * BM / 4;
      for (int e = tid; e < (ka_end - ka_begin) * kChunks; e += kUmmaThreads) {
        const int slot = e / kChunks, q = e - slot * kChunks;
        const int o = m0 + 4 * q;
        int* dst = s_rows + slot * (MT * BM) + 4 * q;
        if (o < ld) cp_async16(smem_u32(dst), map_t + (size_t)active_k[ka_begin + slot] * ld + o, 16u);
        else dst[0] = dst[1] = dst[2] = dst[3] = -1;
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
      for (int e = tid; e < (ka_end - ka_begin) * (MT * BM); e += kUmmaThreads) {
        const int r = e % (MT * BM);
        s_rows[e] = (m0 + r < n_out) ? m0 + r : -1;
      }
    }
  }
  pdl_wait();   // everything above read geometry and weights only; activations follow
  __syncthreads();
  if (threadIdx.x == 0) {
    SPVD_MARK(2);
    SPVD_ACC(16, (unsigned long long)num_iters);
  }

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const int chunk16 = tid & 7;  // which 16-byte piece of the 128-byte row
    const int r0 = tid >> 3;      // rows r0 + 16*i of each 128-row tile
    int rows[MT * 8];
    auto load_rows = [&](int ka) {
      const int* sr = s_rows + (ka - ka_begin) * (MT * BM);
#pragma unroll
      for (int i = 0; i < MT * 8; ++i) rows[i] = sr[(i >> 3) * BM + r0 + 16 * (i & 7)];
    };
    if constexpr (kAsyncA) {
      for (int it = it_begin; it < it_end; ++it) {
        const int li = it - it_begin;
        const int stage = li % kStages;
        const uint32_t parity = (uint32_t)(li / kStages) & 1u;
        const int ka = it / nchunks, c = it - ka * nchunks;
        const int k = active_k[ka];
        if (c == 0 || li == 0) load_rows(ka);
        const long long tw0 = SPVD_CLK();
        mbar_wait(bar_base + 8 * (kStages + stage), parity ^ 1u);
        if (tid == 0) SPVD_ACC(17, (unsigned long long)(SPVD_CLK() - tw0));
        const uint32_t a_s = base + stage * Cfg::kStageBytes;
        if (tid == 0) {
          mbar_arrive_expect_tx(bar_base + 8 * stage, Cfg::kBTileBytes);
          const uint8_t* wsrc = w_b + ((size_t)(k * nchunks + c) * cout + n0) * 128;
          if constexpr (kCluster > 1) {
            constexpr uint32_t kSlice = Cfg::kBTileBytes / kCluster;
            const uint32_t rk = cluster_ctarank();
            bulk_g2s_multicast(a_s + MT * kATileBytes + rk * kSlice, wsrc + rk * kSlice,
                               kSlice, bar_base + 8 * stage, (uint16_t)((1u << kCluster) - 1u));
          } else {
            bulk_g2s(a_s + MT * kATileBytes, wsrc, Cfg::kBTileBytes, bar_base + 8 * stage);
          }
        }
#pragma unroll
        for (int i = 0; i < MT * 8; ++i) {
          const int r = r0 + 16 * (i & 7);
          const uint32_t dst = a_s + (i >> 3) * kATileBytes + r * 128 + ((chunk16 ^ (r & 7)) << 4);
          const uint32_t col_b = (uint32_t)c * 128u + (uint32_t)chunk16 * 16u;
          const bool ok = rows[i] >= 0 && (!kF16 || col_b < row_bytes);
          const uint8_t* src = ok ? in_b + (size_t)rows[i] * row_bytes + col_b : in_b;
          cp_async16(dst, src, ok ? 16u : 0u);
        }
        cp_async_mbar_arrive_noinc(bar_base + 8 * stage);
        if (tid == 0 && li == 0) SPVD_MARK(3);
      }
      if (tid == 0) SPVD_MARK(4);
    } else if constexpr (kF16) {
      // fp32 rows converted on the fly: this thread's 16-byte f16 piece of a row is 8 input channels = two float4
      float4 v[16];
      auto issue_loads = [&](int it) {
        const int ka = it / nchunks, c = it - ka * nchunks;
        if (c == 0 || it == it_begin) load_rows(ka);
        const int ch = c * 64 + chunk16 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[2 * i] = v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rows[i] >= 0 && ch < cin) {
            const float4* p = reinterpret_cast<const float4*>(in + (size_t)rows[i] * cin + ch);
            v[2 * i] = *p;          // (plain loads: the predecessor grid wrote these under PDL)
            v[2 * i + 1] = *(p + 1);
          }
        }
      };
      if (num_iters > 0) issue_loads(it_begin);
      for (int it = it_begin; it < it_end; ++it) {
        const int li = it - it_begin;
        const int stage = li % kStages;
        const uint32_t parity = (uint32_t)(li / kStages) & 1u;
        const int ka = it / nchunks, c = it - ka * nchunks;
        const int k = active_k[ka];
        uint32_t t[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i][0] = pack_half2_sat(v[2 * i].x, v[2 * i].y);
          t[i][1] = pack_half2_sat(v[2 * i].z, v[2 * i].w);
          t[i][2] = pack_half2_sat(v[2 * i + 1].x, v[2 * i + 1].y);
          t[i][3] = pack_half2_sat(v[2 * i + 1].z, v[2 * i + 1].w);
        }
        if (it + 1 < it_end) issue_loads(it + 1);   // next slab's loads fly while this one is stored
        mbar_wait(bar_base + 8 * (kStages + stage), parity ^ 1u);
        const uint32_t a_s = base + stage * Cfg::kStageBytes;
        if (tid == 0) {
          mbar_arrive_expect_tx(bar_base + 8 * stage, Cfg::kBTileBytes);
          bulk_g2s(a_s + MT * kATileBytes, w_b + ((size_t)(k * nchunks + c) * cout + n0) * 128, Cfg::kBTileBytes,
                   bar_base + 8 * stage);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const uint32_t dst = a_s + r * 128 + ((chunk16 ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(t[i][0]), "r"(t[i][1]), "r"(t[i][2]),
                       "r"(t[i][3])
                       : "memory");
        }
        fence_proxy_async();
        mbar_arrive(bar_base + 8 * stage);
      }
    } else {
      float4 v[8];
      auto issue_loads = [&](int it) {
        const int ka = it / nchunks, c = it - ka * nchunks;
        if (c == 0 || it == it_begin) load_rows(ka);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rows[i] >= 0) v[i] = *(reinterpret_cast<const float4*>(in + (size_t)rows[i] * cin + c * BK) + chunk16);
        }
      };
      if (num_iters > 0) issue_loads(it_begin);
      for (int it = it_begin; it < it_end; ++it) {
        const int li = it - it_begin;
        const int stage = li % kStages;
        const uint32_t parity = (uint32_t)(li / kStages) & 1u;
        const int ka = it / nchunks, c = it - ka * nchunks;
        const int k = active_k[ka];
        uint32_t t[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i][0] = to_tf32(v[i].x);
          t[i][1] = to_tf32(v[i].y);
          t[i][2] = to_tf32(v[i].z);
          t[i][3] = to_tf32(v[i].w);
        }
        if (it + 1 < it_end) issue_loads(it + 1);   // next slab's loads fly while this one is stored
        mbar_wait(bar_base + 8 * (kStages + stage), parity ^ 1u);
        const uint32_t a_s = base + stage * Cfg::kStageBytes;
        if (tid == 0) {
          mbar_arrive_expect_tx(bar_base + 8 * stage, Cfg::kBTileBytes);
          const uint8_t* wsrc = w_b + ((size_t)(k * nchunks + c) * cout + n0) * 128;
          if constexpr (kCluster > 1) {
            constexpr uint32_t kSlice = Cfg::kBTileBytes / kCluster;
            const uint32_t rk = cluster_ctarank();
            bulk_g2s_multicast(a_s + MT * kATileBytes + rk * kSlice, wsrc + rk * kSlice,
                               kSlice, bar_base + 8 * stage, (uint16_t)((1u << kCluster) - 1u));
          } else {
            bulk_g2s(a_s + MT * kATileBytes, wsrc, Cfg::kBTileBytes, bar_base + 8 * stage);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const uint32_t dst = a_s + r * 128 + ((chunk16 ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(t[i][0]), "r"(t[i][1]), "r"(t[i][2]),
                       "r"(t[i][3])
                       : "memory");
        }
        fence_proxy_async();
        mbar_arrive(bar_base + 8 * stage);
      }
    }
    // ------------------------------------------------------------------ epilogue
    if (num_iters > 0) {
      mbar_wait(bar_base + 8 * (2 * kStages), 0);
      tc_fence_after();
    }
    if (tid == 0) SPVD_MARK(7);
    const bool first = kPairs ? (active_k[0] == center_k) : (blockIdx.z == 0);
    const bool add_bias = bias != nullptr && first;
    const bool add_res = residual != nullptr && first;
    const bool wide = ((reinterpret_cast<size_t>(out) | reinterpret_cast<size_t>(residual)) & 31) == 0;   // 256-bit accesses
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      int row = m0 + mt * BM + warp * 32 + lane;
      if constexpr (kPairs) {
        row = s_rows[BM + warp * 32 + lane];
        if (row < 0) row = n_out;   // padding pair: nothing to write
      } else if (pair_out) {
        // mask-sorted tiles (spvd_kmap_sort): tile row j is output row pair_out[j]
        row = row < n_out ? __ldg(pair_out + row) : n_out;
      }
      // residual of the row, one 32-column chunk ahead of its use (the loads fly under tcgen05.ld and the stores)
      const bool pre_res = add_res && !split && wide && row < n_out;
      float q[32];
      auto load_res = [&](int cb) {
        const float* rp = residual + (size_t)row * cout + n0 + cb * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=f"(q[8 * j]), "=f"(q[8 * j + 1]), "=f"(q[8 * j + 2]), "=f"(q[8 * j + 3]), "=f"(q[8 * j + 4]),
                         "=f"(q[8 * j + 5]), "=f"(q[8 * j + 6]), "=f"(q[8 * j + 7])
                       : "l"(rp + 8 * j));
      };
      if (pre_res) load_res(0);
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t r[32];
        if (num_iters > 0) {
          const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * BN + cb * 32);
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
          const float* bp = add_bias ? bias + n0 + cb * 32 : nullptr;
          const float* rp = add_res ? residual + (size_t)row * cout + n0 + cb * 32 : nullptr;
          if (!split && wide) {
            // one full 32-byte sector per lane and instruction (LDG/STG.256)
            if (pre_res) {
#pragma unroll
              for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + q[e]);
              if (cb + 1 < BN / 32) load_res(cb + 1);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r[8 * j + e]);
              if (bp) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp) + 2 * j), b1 = __ldg(reinterpret_cast<const float4*>(bp) + 2 * j + 1);
                o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
              }
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dstf + 8 * j), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                           "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                           : "memory");
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
              if (bp) {
                float4 b = __ldg(reinterpret_cast<const float4*>(bp) + j);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              if (rp) {
                float4 b = *reinterpret_cast<const float4*>(rp + 4 * j);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              if (split) red_add_v4(dstf + 4 * j, o);
              else reinterpret_cast<float4*>(dstf)[j] = o;
            }
          }
        }
      }
    }
    if (tid == 0) SPVD_MARK(8);
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    for (int li = 0; li < num_iters; ++li) {
      const int stage = li % kStages;
      const uint32_t parity = (uint32_t)(li / kStages) & 1u;
      const long long tw0 = SPVD_CLK();
      mbar_wait(bar_base + 8 * stage, parity);
      if (lane == 0) {
        SPVD_ACC(18, (unsigned long long)(SPVD_CLK() - tw0));
        if (li == 0) SPVD_MARK(5);
      }
      if constexpr (kAsyncA) fence_proxy_async();   // cp.async wrote the A tiles through the generic proxy
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_s = base + stage * Cfg::kStageBytes;
        const uint64_t bdesc = make_desc(a_s + MT * kATileBytes);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint64_t adesc = make_desc(a_s + mt * kATileBytes);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk)
            if constexpr (kF16)   // 4 x (K = 16 halves); the 32-byte K step of the descriptors is the same as for TF32
              umma_f16(tmem_d + (uint32_t)(mt * BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), Cfg::kIdescF16,
                       (li > 0 || kk > 0) ? 1u : 0u);
            else
              umma_tf32(tmem_d + (uint32_t)(mt * BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), Cfg::kIdesc,
                        (li > 0 || kk > 0) ? 1u : 0u);
        }
        if constexpr (kCluster > 1) umma_commit_multicast(bar_base + 8 * (kStages + stage), (uint16_t)((1u << kCluster) - 1u));
        else umma_commit(bar_base + 8 * (kStages + stage));
        if (li == num_iters - 1) umma_commit(bar_base + 8 * (2 * kStages));
      }
      __syncwarp();
    }
    if (lane == 0) SPVD_MARK(6);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    SPVD_MARK(9);
    SPVD_GMARK(11);
  }
  if constexpr (kCluster > 1) cluster_sync_all();   // no CTA leaves while a peer may still write its smem / barriers
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

// weights [kvol][cin][cout] -> per (k, 32-channel slab) image of the K-major SWIZZLE_128B B tile:
// element (n, kk) at float offset ((k*nchunks + c)*cout + n)*32 + (((kk>>2) ^ (n&7))<<2) + (kk&3)
static __global__ void k_pack_weights(const float* __restrict__ w, int kvol, int cin, int cout, float* __restrict__ wp) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)kvol * cin * cout;
  if (t >= total) return;
  int n = (int)(t % cout);
  long long q = t / cout;
  int ci = (int)(q % cin);
  int k = (int)(q / cin);
  int c = ci / BK, kk = ci % BK;
  size_t dst = ((size_t)(k * (cin / BK) + c) * cout + n) * BK + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
  wp[dst] = __uint_as_float(to_tf32(w[t]));
}

template <int BN, int MT, bool kAsyncA, bool kLite = false, bool kF16 = false>
static int launch_umma(const float* in, const float* wp, const int32_t* map_t, int ld, const uint32_t* tile_mask,
                       int n_out, int kvol, int cin, int cout, const float* bias, const float* residual, float* out,
                       int nsplit, cudaStream_t s, const int32_t* out_rows = nullptr) {
  using Cfg = UmmaCfg<BN, MT, false, kLite>;
  static bool configured_dev[64] = {};   // the attribute is per device: one flag per device, not per process
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(k_conv_fwd_umma<BN, MT, kAsyncA, false, kLite, 1, kF16>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  if (nsplit > 1) SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * cout * 4, s));
  dim3 grid(cdiv(n_out, BM * MT), cout / BN, nsplit);
  SPVD_CUDA(launch_pdl(k_conv_fwd_umma<BN, MT, kAsyncA, false, kLite, 1, kF16>, grid, dim3(kUmmaThreads),
                       (size_t)Cfg::kSmemBytes, s, in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out,
                       (const int*)out_rows, 0));
  return SPVD_OK;
}

template <int BN, bool kAsyncA, bool kF16 = false>
static int launch_umma_pairs_a(const float* in, const float* wp, const int32_t* pair_in, const int32_t* pair_out,
                               const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int kvol, int cin, int cout,
                               const float* bias, const float* residual, float* out, cudaStream_t s) {
  using Cfg = UmmaCfg<BN, 1, true>;
  static bool configured_dev[64] = {};   // the attribute is per device: one flag per device, not per process
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(k_conv_fwd_umma<BN, 1, kAsyncA, true, false, 1, kF16>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * cout * 4, s));
  if (n_pair_tiles == 0) return SPVD_OK;
  dim3 grid(n_pair_tiles, cout / BN, 1);
  SPVD_CUDA(launch_pdl(k_conv_fwd_umma<BN, 1, kAsyncA, true, false, 1, kF16>, grid, dim3(kUmmaThreads),
                       (size_t)Cfg::kSmemBytes, s, in, wp, pair_in, 0, (const unsigned*)tile_k, n_out, kvol, cin, cout, bias,
                       residual, out, pair_out, center_k));
  return SPVD_OK;
}

template <int BN>
static int launch_umma_pairs(bool async_a, const float* in, const float* wp, const int32_t* pair_in,
                             const int32_t* pair_out, const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out,
                             int kvol, int cin, int cout, const float* bias, const float* residual, float* out,
                             cudaStream_t s) {
  return async_a ? launch_umma_pairs_a<BN, true>(in, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol,
                                                 cin, cout, bias, residual, out, s)
                 : launch_umma_pairs_a<BN, false>(in, wp, pair_in, pair_out, tile_k, n_pair_tiles, center_k, n_out, kvol,
                                                  cin, cout, bias, residual, out, s);
}

template <int BN, int kCluster, bool kF16 = false>
static int launch_umma_cluster(const float* in, const float* wp, const int32_t* map_t, int ld, const uint32_t* tile_mask,
                               int n_out, int kvol, int cin, int cout, const float* bias, const float* residual, float* out,
                               int nsplit, cudaStream_t s, const int32_t* out_rows = nullptr) {
  using Cfg = UmmaCfg<BN, 1, false, true>;
  auto kern = k_conv_fwd_umma<BN, 1, true, false, true, kCluster, kF16>;
  static bool configured_dev[64] = {};   // the attribute is per device: one flag per device, not per process
  bool& configured = configured_dev[current_device_slot()];
  if (!configured) {
    SPVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  if (nsplit > 1) SPVD_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * cout * 4, s));
  const int tiles = cdiv(cdiv(n_out, BM), kCluster) * kCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, cout / BN, nsplit);
  cfg.blockDim = dim3(kUmmaThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  SPVD_CUDA(cudaLaunchKernelEx(&cfg, kern, in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out,
                               (const int*)out_rows, 0));
  return SPVD_OK;
}

template <int BN>
static int launch_umma_bn(bool async_a, int mt, const float* in, const float* wp, const int32_t* map_t, int ld,
                          const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout, const float* bias,
                          const float* residual, float* out, int nsplit, cudaStream_t s, const int32_t* out_rows = nullptr) {
  if (!async_a) {
    if (mt == 3)
      return launch_umma<BN, 1, false, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
    return launch_umma<BN, 1, false>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  }
#ifdef SPVD_EXPERIMENTAL   // cluster multicast of the weight tile / two tiles per CTA: correct, measured no faster
  if (mt == 4)   // lite + weight tiles multicast across a cluster of 2 CTAs
    return launch_umma_cluster<BN, 2>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  if (mt == 5)
    return launch_umma_cluster<BN, 4>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  if (mt == 2)
    return launch_umma<BN, 2, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
#endif
  if (mt >= 2)   // one tile per CTA, two stages: small enough for two CTAs per SM
    return launch_umma<BN, 1, true, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
  return launch_umma<BN, 1, true>(in, wp, map_t, ld, tile_mask, n_out, kvol, cin, cout, bias, residual, out, nsplit, s, out_rows);
}

// widest instantiated slice that divides cout: the A tile is gathered once per slice, so fewer is better
static int pick_bn(int cout, int n_tiles = 1 << 30, int kvol = 27) {
  const int widths[] = {256, 192, 128, 96, 64, 32};
  if (kvol == 1) {
    // dense layer (nn.Linear): only cin/32 iterations per tile, nothing to split -> narrow the slices until the
    // grid covers the GPU (the A tile of a dense layer is cheap to re-read)
    for (int w : widths)
      if (cout % w == 0 && (long long)n_tiles * (cout / w) >= 148) return w;
    for (int i = 5; i >= 0; --i)
      if (cout % widths[i] == 0 && widths[i] >= 64) return widths[i];
  }
  if (n_tiles < 32) {
    // a handful of tiles (stride 16): slices of at most 128 columns and more split-K fill the GPU better than one wide
    // slice per tile (profiles/r02b_small_level_sweep.txt: 192->192 23.6 -> 19.2 us, 256->256 27.6 -> 23.6 us)
    for (int w : widths)
      if (w <= 128 && w >= 64 && cout % w == 0) return w;
  }
  for (int w : widths)
    if (cout % w == 0) return w;
  return 32;
}

// split the (offset, slab) iterations over gridDim.z when the tiles alone cannot fill the GPU
static int pick_split(int n_ctas, int max_iters) {
  const int target = 2 * 148;   // two CTAs per SM: stay within ONE wave (a second, partial wave costs a full tile time)
  if (n_ctas >= 148) return 1;
  int s = target / n_ctas;
  int cap = max_iters / 6;  // keep >= 6 pipeline iterations per CTA
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}

}  // namespace spvd
