// EXPERIMENT (not part of libspvd_b200.so; built and driven by tools/_halo_sweep.py on the GPU box): "halo" sparse
// convolution on tcgen05 (f16 operands, fp32 accumulate in TMEM) -- the gather of the input rows is taken out of the
// (offset, slab) loop.
//
// The one-tile kernels fetch 128 rows of 128 bytes from L2 with LDGSTS in every (tile, offset, slab) step although the 27
// offsets of 128 consecutive voxels touch only ~300-500 DISTINCT input rows (7-14x reuse at strides 4-16, 3-5x at stride 2).
// Here a CTA owns two consecutive 128-row tiles and, per 64-channel slab,
//   1. loads the distinct rows of its tiles ONCE into shared memory (the "halo", <= kHaloCap rows x 128 B; the row list and
//      the per-(offset, row) index into it are planned per level by spvd_kmap_halo),
//   2. per active offset re-slices the halo into the canonical K-major SWIZZLE_128B A tile with a shared -> shared copy,
//   3. issues the MMAs of both tiles against ONE weight tile (bulk-copied per (offset, slab)).
// RESULT (profiles/r02_halo_kernel_sweep.txt): bit-exact against the production kernel but 5-15 % SLOWER
// on the large levels and 2-3x slower on the small ones.  The phase clocks (-DSPVD_HALO_TRACE) show why: the L2 gather
// is replaced by 32 KB of shared-memory traffic per A tile (16 KB read + 16 KB written by the re-slicers) on top of the
// MMA's own operand reads and the weight stream, and one CTA per SM (the halo fills the shared memory) leaves the
// producer -> MMA ring 2-3 stages deep, so a tile still costs ~700 cycles per SM.  Kept as the record of that measurement.
#include <cuda_fp16.h>

#include "../../spvd_b200/csrc/conv_umma_impl.cuh"

namespace spvd {
thread_local char g_err[512];
void set_error(const char* fmt, ...) { (void)fmt; }

constexpr int kHaloCap = 640;                 // rows of the halo buffer (80 KB); a group with more distinct rows falls back
constexpr int kHashSlots = 4096;              // planning: open-addressing set of the group's rows

// ------------------------------------------------------------------------------------------------ planning
// One CTA per group of MT tiles: the distinct input rows the group references over all offsets (hash set in shared
// memory), and for every (offset, tile row) the position of its input row in that list (0xFFFF = no neighbour).
template <int MT>
__global__ void __launch_bounds__(256) k_halo_plan(const int* __restrict__ map_t, int ld, int kvol, const int* __restrict__ n_dev,
                                                   int n_cap, int* __restrict__ halo_rows, int* __restrict__ halo_count,
                                                   unsigned short* __restrict__ lidx, int* __restrict__ max_count) {
  __shared__ int keys[kHashSlots];
  __shared__ int pos[kHashSlots];
  __shared__ int wsum[8];
  __shared__ int total;
  __shared__ int overflow;
  const int n = resolve_n(n_dev, n_cap);
  const int g = blockIdx.x, row0 = g * MT * BM;
  if (row0 >= n) {
    if (threadIdx.x == 0) halo_count[g] = 0;
    return;
  }
  for (int i = threadIdx.x; i < kHashSlots; i += 256) keys[i] = -1;
  if (threadIdx.x == 0) overflow = 0;
  __syncthreads();
  const int entries = kvol * MT * BM;
  for (int e = threadIdx.x; e < entries; e += 256) {
    const int k = e / (MT * BM), r = e - k * (MT * BM);
    const int v = row0 + r < ld ? __ldg(map_t + (size_t)k * ld + row0 + r) : -1;
    if (v < 0) continue;
    unsigned s = ((unsigned)v * 2654435761u) >> 20;          // 12 bits
    int probes = 0;
    while (true) {
      const int old = atomicCAS(&keys[s], -1, v);
      if (old == -1 || old == v) break;
      s = (s + 1) & (kHashSlots - 1);
      if (++probes >= kHashSlots) {     // more distinct rows than slots (degenerate geometry): report, the caller falls back
        overflow = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (overflow) {
    if (threadIdx.x == 0) {
      halo_count[g] = 1 << 20;
      atomicMax(max_count, 1 << 20);
    }
    return;
  }
  // compact the occupied slots: position = exclusive prefix count (16 slots per thread)
  int cnt = 0;
  const int s0 = threadIdx.x * (kHashSlots / 256);
  for (int i = 0; i < kHashSlots / 256; ++i) cnt += keys[s0 + i] >= 0;
  int incl = cnt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += wsum[w];
  if (threadIdx.x == 255) total = base + incl;
  int p = base + incl - cnt;
  for (int i = 0; i < kHashSlots / 256; ++i) {
    const int v = keys[s0 + i];
    pos[s0 + i] = p;
    if (v >= 0) {
      if (p < kHaloCap) halo_rows[(size_t)g * kHaloCap + p] = v;
      ++p;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    halo_count[g] = total;
    atomicMax(max_count, total);
  }
  for (int e = threadIdx.x; e < entries; e += 256) {
    const int k = e / (MT * BM), r = e - k * (MT * BM);
    const int v = row0 + r < ld ? __ldg(map_t + (size_t)k * ld + row0 + r) : -1;
    unsigned short li = 0xFFFF;
    if (v >= 0) {
      unsigned s = ((unsigned)v * 2654435761u) >> 20;
      while (keys[s] != v) s = (s + 1) & (kHashSlots - 1);
      li = (unsigned short)min(pos[s], 0xFFFE);
    }
    lidx[((size_t)g * kvol + k) * (MT * BM) + r] = li;
  }
}

// ------------------------------------------------------------------------------------------------ the convolution
// A CTA owns a group of two consecutive 128-row tiles and runs TWO independent producer -> MMA chains, one per tile:
// measured (tools/_halo_sweep.py with phase clocks), one MMA-issuing warp spends ~600-700 cycles per A tile on its own
// serial chain (two barrier waits at 60-90 cycles each, the proxy fence, ~60 cycles per tcgen05.mma whatever N <= 256 is,
// two commits), and four re-slicing warps need about as long per tile; two chains keep the tensor pipe fed.
//   warps 0-3  re-slice tile 0, then run its epilogue           warp 8   issues the MMAs of tile 0
//   warps 4-7  re-slice tile 1, then run its epilogue           warp 9   issues the MMAs of tile 1
//   all of 0-7 load the halo of a slab                          warp 10  streams the weight tiles (one bulk copy per
//                                                                        (offset, slab), consumed by both chains)
constexpr int kHalo2Threads = 352;

template <int BN>
struct Halo2Cfg {
  static constexpr int SA = BN <= 64 ? 3 : 2;        // A stages per chain
  static constexpr int SW = BN <= 128 ? 3 : 2;       // weight stages (shared by the chains)
  static constexpr int kWBytes = BN * 128;
  static constexpr int kCols = BN * 2;
  static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
  static constexpr int kHaloBytes = kHaloCap * 128;
  static constexpr int kIdxBytes = 27 * 2 * BM * 2;
  static constexpr int kRowsBytes = kHaloCap * 4;
  static constexpr int kSmemBytes = 1024 + kHaloBytes + 2 * SA * kATileBytes + SW * kWBytes + 256 + kIdxBytes + kRowsBytes;
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(kCols <= 512, "accumulators exceed TMEM");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory");
};

#ifdef SPVD_HALO_TRACE
__device__ unsigned long long g_halo_trace[32];
#define HT(i) do { long long t_ = clock64(); acc_[i] += t_ - t0_; t0_ = t_; } while (0)
#define HT_DECL long long acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0_ = clock64()
#define HT_DUMP(base, n) do { if (blockIdx.x == 3) { for (int i_ = 0; i_ < 8; ++i_) g_halo_trace[(base) + i_] = (unsigned long long)acc_[i_]; g_halo_trace[24 + (base) / 8] = (unsigned long long)(n); } } while (0)
#else
#define HT(i) do {} while (0)
#define HT_DECL do {} while (0)
#define HT_DUMP(base, n) do {} while (0)
#endif

template <int BN>
__global__ void __launch_bounds__(kHalo2Threads, 1) k_conv_halo2_f16(
    const __half* __restrict__ in, const void* __restrict__ w_packed, const unsigned* __restrict__ tile_mask,
    const int* __restrict__ halo_rows, const int* __restrict__ halo_count, const unsigned short* __restrict__ lidx, int n_out,
    int kvol, int cin, int cout, const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ out) {
  using Cfg = Halo2Cfg<BN>;
  constexpr int SA = Cfg::SA, SW = Cfg::SW, MT = 2;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t halo_s = base;
  const uint32_t a_s0 = halo_s + Cfg::kHaloBytes;                 // chain c, stage s: a_s0 + (c * SA + s) * 16 KB
  const uint32_t w_s0 = a_s0 + 2 * SA * kATileBytes;
  const uint32_t bar0 = w_s0 + SW * Cfg::kWBytes;
  const uint32_t bar_afull = bar0, bar_aempty = bar0 + 8 * 2 * SA, bar_wfull = bar0 + 8 * 4 * SA, bar_wempty = bar_wfull + 8 * SW,
                 bar_acc = bar_wempty + 8 * SW;                   // bar_acc: two barriers, one per chain
  uint8_t* tail = smem + (bar0 - base);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 8 * (4 * SA + 2 * SW + 2));
  unsigned short* s_lidx = reinterpret_cast<unsigned short*>(tail + 256);
  int* s_rows = reinterpret_cast<int*>(tail + 256 + Cfg::kIdxBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.x, n0 = blockIdx.y * BN;
  const int n_tiles = (n_out + BM - 1) / BM;
  const int nchunks = (cin + 63) / 64;
  const unsigned kmask = kvol >= 32 ? 0xffffffffu : ((1u << kvol) - 1u);
  unsigned tmask[MT], umask = 0;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int tile = g * MT + mt;
    tmask[mt] = tile < n_tiles ? ((tile_mask ? tile_mask[tile] : kmask) & kmask) : 0u;
    umask |= tmask[mt];
  }
  const int n_halo = min(halo_count[g], kHaloCap);

  if (tid == 0) {
    for (int s = 0; s < 2 * SA; ++s) {
      mbar_init(bar_afull + 8 * s, 4);         // one arrival per re-slicing warp
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < SW; ++s) {
      mbar_init(bar_wfull + 8 * s, 1);
      mbar_init(bar_wempty + 8 * s, 2);        // both chains release a weight stage
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the group's plan: halo row list and (offset, row) -> halo position table (geometry: ready long before the activations)
  {
    const int idx_chunks = kvol * MT * BM * 2 / 16;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(lidx + (size_t)g * kvol * MT * BM);
    for (int e = tid; e < idx_chunks; e += kHalo2Threads) cp_async16(smem_u32(s_lidx) + 16 * e, src + 16 * e, 16u);
    const uint8_t* rsrc = reinterpret_cast<const uint8_t*>(halo_rows + (size_t)g * kHaloCap);
    for (int e = tid; e < kHaloCap / 4; e += kHalo2Threads) cp_async16(smem_u32(s_rows) + 16 * e, rsrc + 16 * e, 16u);
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_wait();

  const uint8_t* in_b = reinterpret_cast<const uint8_t*>(in);
  const uint32_t row_bytes = (uint32_t)cin * 2u;

  if (warp < 8) {
    // ------------------------------------------------------------------ halo loader + re-slicer (2 x 128 threads)
    const int ch = warp >> 2;                  // chain = tile of the group
    const int t128 = tid & 127;
    const int chunk16 = t128 & 7, r0 = t128 >> 3;
    const unsigned my_mask = tmask[ch];
    int ai = 0;
    HT_DECL;
    for (int c = 0; c < nchunks; ++c) {
      HT(0);
      // every re-slicer has finished reading the previous slab's halo before anybody overwrites it
      asm volatile("barrier.sync 1, 256;" ::: "memory");
      const uint32_t col_b = (uint32_t)c * 128u + (uint32_t)chunk16 * 16u;
      const bool ok = col_b < row_bytes;
      for (int u = tid >> 3; u < n_halo; u += 32)
        cp_async16(halo_s + u * 128 + chunk16 * 16, ok ? in_b + (size_t)s_rows[u] * row_bytes + col_b : in_b, ok ? 16u : 0u);
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("barrier.sync 1, 256;" ::: "memory");     // the slab's halo is complete and visible to all re-slicers
      HT(1);
      unsigned rem = my_mask;
      while (rem) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1;
        const int as = ai % SA;
        HT(0);
        mbar_wait(bar_aempty + 8 * (ch * SA + as), (((uint32_t)(ai / SA)) & 1u) ^ 1u);
        HT(2);
        const uint32_t a_s = a_s0 + (ch * SA + as) * kATileBytes;
        const unsigned short* li = s_lidx + (k * MT + ch) * BM;
        // all eight index loads, then all eight row-piece loads, then the stores: three dependent shared-memory round
        // trips per tile instead of twenty-four
        unsigned l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = li[r0 + 16 * i];
        uint32_t v[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0u;
          if (l[i] != 0xFFFFu)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3])
                         : "r"(halo_s + l[i] * 128 + chunk16 * 16));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_s + r * 128 + ((chunk16 ^ (r & 7)) << 4)), "r"(v[i][0]),
                       "r"(v[i][1]), "r"(v[i][2]), "r"(v[i][3])
                       : "memory");
        }
        HT(3);
        __syncwarp();                                     // one arrival per warp (the proxy fence is the consumer's)
        if (lane == 0) mbar_arrive(bar_afull + 8 * (ch * SA + as));
        HT(4);
        ++ai;
      }
    }
    // ------------------------------------------------------------------ epilogue of tile `ch`
    HT(0);
    mbar_wait(bar_acc + 8 * ch, 0);
    HT(5);
    if (tid == 0) HT_DUMP(0, ai);
    tc_fence_after();
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int row = (g * MT + ch) * BM + q * 32 + lane;
    const bool any = my_mask != 0u;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      uint32_t r[32];
      if (any) {
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * BN + cb * 32);
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
        const size_t off = (size_t)row * cout + n0 + cb * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n0 + cb * 32) + j);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          if (residual) {
            const float4 b = *(reinterpret_cast<const float4*>(residual + off) + j);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          *(reinterpret_cast<float4*>(out + off) + j) = v;
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ MMA issuer of chain `ch`
    const int ch = warp - 8;
    const unsigned my_mask = tmask[ch];
    int ai = 0, wi = 0;
    bool started = false;
    HT_DECL;
    for (int c = 0; c < nchunks; ++c) {
      unsigned rem = umask;
      while (rem) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1;
        const int ws = wi % SW;
        HT(0);
        mbar_wait(bar_wfull + 8 * ws, ((uint32_t)(wi / SW)) & 1u);
        HT(1);
        if ((my_mask >> k) & 1u) {
          const int as = ai % SA;
          mbar_wait(bar_afull + 8 * (ch * SA + as), ((uint32_t)(ai / SA)) & 1u);
          HT(2);
          fence_proxy_async();     // the re-slicers wrote the A tile through the generic proxy
          tc_fence_after();
          HT(3);
          if (lane == 0) {
            const uint64_t adesc = make_desc(a_s0 + (ch * SA + as) * kATileBytes);
            const uint64_t bdesc = make_desc(w_s0 + ws * Cfg::kWBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_d + (uint32_t)(ch * BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), Cfg::kIdesc,
                       (started || kk > 0) ? 1u : 0u);
            HT(4);
            umma_commit(bar_aempty + 8 * (ch * SA + as));
            umma_commit(bar_wempty + 8 * ws);
            HT(5);
          }
          __syncwarp();
          HT(6);
          started = true;
          ++ai;
        } else {
          // this tile has no neighbour at offset k: release the weight stage without using it (after it has filled, so that
          // the release can never run ahead into the stage's next use)
          if (lane == 0) mbar_arrive(bar_wempty + 8 * ws);
          __syncwarp();
        }
        ++wi;
      }
    }
    if (lane == 0) umma_commit(bar_acc + 8 * ch);
    __syncwarp();
    if (lane == 0 && ch == 0) HT_DUMP(8, ai);
  } else {
    // ------------------------------------------------------------------ weight streamer
    if (lane == 0) {
      const uint8_t* w_b = reinterpret_cast<const uint8_t*>(w_packed);
      int wi = 0;
      HT_DECL;
      for (int c = 0; c < nchunks; ++c) {
        unsigned rem = umask;
        while (rem) {
          const int k = __ffs(rem) - 1;
          rem &= rem - 1;
          const int ws = wi % SW;
          HT(0);
          mbar_wait(bar_wempty + 8 * ws, (((uint32_t)(wi / SW)) & 1u) ^ 1u);
          HT(1);
          mbar_arrive_expect_tx(bar_wfull + 8 * ws, Cfg::kWBytes);
          bulk_g2s(w_s0 + ws * Cfg::kWBytes, w_b + ((size_t)(k * nchunks + c) * cout + n0) * 128, Cfg::kWBytes, bar_wfull + 8 * ws);
          ++wi;
        }
      }
      HT_DUMP(16, wi);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

template <int BN>
static int launch_halo(const void* in, const void* wp, const uint32_t* tile_mask, const int32_t* halo_rows, const int32_t* halo_count,
                       const uint16_t* lidx, int n_out, int kvol, int cin, int cout, const float* bias, const float* residual,
                       float* out, cudaStream_t s) {
  using Cfg = Halo2Cfg<BN>;
  auto kern = k_conv_halo2_f16<BN>;
  static bool configured[64] = {};
  const int dev = current_device_slot();
  if (!configured[dev]) {
    SPVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured[dev] = true;
  }
  const int n_groups = cdiv(cdiv(n_out, BM), 2);
  SPVD_CUDA(launch_pdl(kern, dim3(n_groups, cout / BN), dim3(kHalo2Threads), (size_t)Cfg::kSmemBytes, s, (const __half*)in, wp, tile_mask,
                       halo_rows, halo_count, (const unsigned short*)lidx, n_out, kvol, cin, cout, bias, residual, out));
  return SPVD_OK;
}

}  // namespace spvd

using namespace spvd;

#ifdef SPVD_HALO_TRACE
extern "C" int spvd_halo_trace(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_halo_trace, sizeof(g_halo_trace)); }
#endif
extern "C" int spvd_halo_cap(void) { return kHaloCap; }

extern "C" int spvd_kmap_halo(const int32_t* map_t, int ld, int n_cap, const int32_t* n_dev, int kvol, int mt, int32_t* halo_rows,
                              int32_t* halo_count, uint16_t* lidx, int32_t* max_count, spvd_stream_t stream) {
  SPVD_CHECK_ARG(map_t && halo_rows && halo_count && lidx && max_count && n_cap > 0 && kvol >= 1 && kvol <= 27 && (mt == 1 || mt == 2),
                 "spvd_kmap_halo: bad arguments (kvol <= 27, mt = 1 or 2)");
  SPVD_CHECK_ARG(ld % 128 == 0 && ld >= n_cap, "spvd_kmap_halo: ld %d must be a multiple of 128 >= n_cap", ld);
  cudaStream_t s = (cudaStream_t)stream;
  SPVD_CUDA(cudaMemsetAsync(max_count, 0, 4, s));
  const int n_groups = cdiv(cdiv(n_cap, BM), mt);
  if (mt == 1) k_halo_plan<1><<<n_groups, 256, 0, s>>>(map_t, ld, kvol, n_dev, n_cap, halo_rows, halo_count, lidx, max_count);
  else k_halo_plan<2><<<n_groups, 256, 0, s>>>(map_t, ld, kvol, n_dev, n_cap, halo_rows, halo_count, lidx, max_count);
  SPVD_LAUNCH_CHECK();
  return SPVD_OK;
}

extern "C" int spvd_conv_halo_f16(const void* in, const void* w_packed, const uint32_t* tile_mask, const int32_t* halo_rows,
                                  const int32_t* halo_count, const uint16_t* lidx, int mt, int n_out, int kvol, int cin, int cout,
                                  const float* bias, const float* residual, float* out, spvd_stream_t stream) {
  SPVD_CHECK_ARG(in && w_packed && halo_rows && halo_count && lidx && out && n_out >= 0 && mt == 2,
                 "spvd_conv_halo_f16: bad arguments (the plan must be built with mt = 2)");
  SPVD_CHECK_ARG(cin % 8 == 0 && kvol <= 27 && cout <= 256 && cout % 32 == 0,
                 "spvd_conv_halo_f16: unsupported shape %dx%dx%d", kvol, cin, cout);
  SPVD_CHECK_ARG(residual == nullptr || residual != out, "spvd_conv_halo_f16: residual must not alias out");
  if (n_out == 0) return SPVD_OK;
  cudaStream_t s = (cudaStream_t)stream;
#define SPVD_HALO(BNV)                                                                                                                 \
  case BNV:                                                                                                                            \
    return launch_halo<BNV>(in, w_packed, tile_mask, halo_rows, halo_count, lidx, n_out, kvol, cin, cout, bias, residual, out, s);
  switch (cout) {
    SPVD_HALO(32)
    SPVD_HALO(64)
    SPVD_HALO(96)
    SPVD_HALO(128)
    SPVD_HALO(192)
    SPVD_HALO(256)
  }
#undef SPVD_HALO
  set_error("spvd_conv_halo_f16: cout %d is not instantiated (32, 64, 96, 128, 192, 256)", cout);
  return SPVD_ERR_INVALID;
}
