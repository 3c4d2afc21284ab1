// Micro-benchmark (not part of libspvd_b200.so; built and run by tools/_gather_bench.py): what does one SM sustain when
// it gathers 128-byte pieces of random rows of an L2-resident matrix into shared memory -- the access pattern of the
// sparse convolution's A operand -- as a function of the warps per SM, the copies in flight per thread and the
// instruction (cp.async.cg / cp.async.ca / ld.global + st.shared)?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: cp.async.cg, 1: cp.async.ca, 2: ld.global.v4 + st.shared.v4
template <int MODE, int DEPTH>
__global__ void k_gather(const uint8_t* __restrict__ mat, int row_bytes, const int* __restrict__ rows, int n_idx, int iters,
                         unsigned long long* __restrict__ clocks, int* __restrict__ sink) {
  extern __shared__ uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk16 = lane & 7, sub = lane >> 3;                 // lane -> 16-byte piece `chunk16` of row `sub` (of 4) per copy
  uint8_t* my = smem + (size_t)warp * DEPTH * 4096;              // per warp: DEPTH groups x (8 copies x 4 rows x 128 B)
  const long long t0 = clock64();
  unsigned acc = 0;
  int cursor = (blockIdx.x * (blockDim.x >> 5) + warp) * 32 * 7919 % n_idx;
  for (int it = 0; it < iters; ++it) {
    uint8_t* dst = my + (size_t)(it % DEPTH) * 4096;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // row indices computed in registers (no dependent index load in front of the copy): rows == nullptr -> pseudo-random
      // rows, else runs of consecutive rows starting at a pseudo-random base
      const unsigned h = ((unsigned)(cursor + (rows ? 0 : 4 * i + sub)) * 2654435761u) >> 7;
      int row = (int)(h % (unsigned)(n_idx - 160)) + (rows ? 4 * i + sub : 0);
      const uint8_t* src = mat + (size_t)row * row_bytes + ((it & 1) * 128 % row_bytes) + chunk16 * 16;
      const uint32_t d = smem_u32(dst + (4 * i + sub) * 128 + ((chunk16 ^ ((4 * i + sub) & 7)) << 4));
      if (MODE == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
      else if (MODE == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
      else {
        const uint4 v = *reinterpret_cast<const uint4*>(src);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    }
    if (MODE < 2) {
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    }
    cursor += 32;
    acc += my[(it * 37 + lane) & 4095];
  }
  if (MODE < 2) asm volatile("cp.async.wait_all;" ::: "memory");
  const long long t1 = clock64();
  if (lane == 0) clocks[blockIdx.x * (blockDim.x >> 5) + warp] = (unsigned long long)(t1 - t0);
  if (acc == 0xffffffffu) *sink = 1;
}

template <int MODE, int DEPTH>
static float run(const uint8_t* mat, int row_bytes, const int* rows, int n_idx, int iters, int ctas, int warps,
                 unsigned long long* clocks, int* sink) {
  const size_t smem = (size_t)warps * DEPTH * 4096;
  cudaFuncSetAttribute(k_gather<MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_gather<MODE, DEPTH><<<ctas, warps * 32, smem>>>(mat, row_bytes, rows, n_idx, 4, clocks, sink);
  cudaEventRecord(a);
  k_gather<MODE, DEPTH><<<ctas, warps * 32, smem>>>(mat, row_bytes, rows, n_idx, iters, clocks, sink);
  cudaEventRecord(b);
  if (cudaEventSynchronize(b) != cudaSuccess) return -1.f;
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// returns milliseconds of the timed launch (or < 0 on error)
extern "C" float gather_bench(const void* mat, int row_bytes, const int* rows, int n_idx, int iters, int ctas, int warps,
                              int mode, int depth, unsigned long long* clocks, int* sink) {
  const uint8_t* m = (const uint8_t*)mat;
#define GB(M, D) if (mode == M && depth == D) return run<M, D>(m, row_bytes, rows, n_idx, iters, ctas, warps, clocks, sink);
  GB(0, 1) GB(0, 2) GB(0, 4) GB(0, 8) GB(1, 1) GB(1, 2) GB(1, 4) GB(1, 8) GB(2, 1)
#undef GB
  return -2.f;
}
