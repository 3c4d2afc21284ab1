// Shared device/host helpers for libspvd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include <stdio.h>

#include "../../include/spvd_b200.h"

namespace spvd {

// ---------------------------------------------------------------------------------------------
// error reporting: thread-local string behind spvd_last_error()
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define SPVD_CHECK_ARG(cond, ...)                     \
  do {                                                \
    if (!(cond)) {                                    \
      spvd::set_error(__VA_ARGS__);                   \
      return SPVD_ERR_INVALID;                        \
    }                                                 \
  } while (0)

#define SPVD_CUDA(call)                                                                 \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      spvd::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                      cudaGetErrorString(e_));                                          \
      return SPVD_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define SPVD_LAUNCH_CHECK() SPVD_CUDA(cudaPeekAtLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// index of the current device for per-device one-time setup flags (cudaFuncSetAttribute is per device)
static inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return dev;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  The feature-path kernels (convolutions and the elementwise glue)
// are launched with programmaticStreamSerialization: each one lets its successor start early
// (pdl_trigger at the top) and itself waits for its predecessor's memory (pdl_wait) only right
// before it touches activations, so launch latency and prologues (barrier init, TMEM allocation,
// staging of geometry, which is complete long before: spvd_plan_build ends with a stream sync, and
// the op-by-op geometry kernels never trigger early) overlap the previous kernel's tail.
// Launched without the attribute, both instructions are no-ops.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// two floats -> packed f16x2 (round to nearest even, clamped to the largest finite half: no infinities reach the MMA);
// the low half is `lo`
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  lo = fminf(fmaxf(lo, -65504.f), 65504.f);
  hi = fminf(fmaxf(hi, -65504.f), 65504.f);
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// float4 number t of a row-major tensor stored as f16: 8 bytes at half offset 4*t
__device__ __forceinline__ void store_half4(void* y, long long t, float4 v) {
  reinterpret_cast<uint2*>(y)[t] = make_uint2(pack_half2_sat(v.x, v.y), pack_half2_sat(v.z, v.w));
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// coordinate keys: (b,x,y,z) -> 64 bit, 16 bits per field, xyz biased so that unsigned key order is
// the lexicographic order of (b,x,y,z) and a valid key is never 0 (0 = empty hash slot, the
// reference zero-initialises its tables: models/sparse_utils.py:37-42).
// ---------------------------------------------------------------------------------------------
constexpr int kCoordBias = 1 << 15;
constexpr int kCoordMax = kCoordBias - 1;
constexpr unsigned long long kKeySentinel = 0xFFFFFFFFFFFFFFFFull;  // sorts last, never a valid key

__host__ __device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  return b >= 0 && b <= 0xFFFE && x >= -kCoordMax && x <= kCoordMax && y >= -kCoordMax &&
         y <= kCoordMax && z >= -kCoordMax && z <= kCoordMax;
}

__host__ __device__ __forceinline__ unsigned long long pack_key(int b, int x, int y, int z) {
  return ((unsigned long long)(unsigned)(b & 0xFFFF) << 48) |
         ((unsigned long long)(unsigned)((x + kCoordBias) & 0xFFFF) << 32) |
         ((unsigned long long)(unsigned)((y + kCoordBias) & 0xFFFF) << 16) |
         (unsigned long long)(unsigned)((z + kCoordBias) & 0xFFFF);
}

__host__ __device__ __forceinline__ int4 unpack_key(unsigned long long k) {
  int4 c;
  c.x = (int)((k >> 48) & 0xFFFF);
  c.y = (int)((k >> 32) & 0xFFFF) - kCoordBias;
  c.z = (int)((k >> 16) & 0xFFFF) - kCoordBias;
  c.w = (int)(k & 0xFFFF) - kCoordBias;
  return c;
}

// murmur3 finaliser; slot by multiply-shift range reduction (capacity need not be a power of 2:
// the reference sizes its tables 2*N, models/sparse_utils.py:37-42)
__device__ __forceinline__ unsigned hash_key(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned)k;
}

__device__ __forceinline__ unsigned hash_slot(unsigned long long k, unsigned cap) {
  return (unsigned)(((unsigned long long)hash_key(k) * cap) >> 32);
}

// lookup: returns stored value (row + 1) or 0 when absent
__device__ __forceinline__ int hash_find(const unsigned long long* __restrict__ keys,
                                         const int* __restrict__ vals, unsigned cap,
                                         unsigned long long key) {
  unsigned slot = hash_slot(key, cap);
  for (unsigned probe = 0; probe < cap; ++probe) {
    unsigned long long k = __ldg(keys + slot);
    if (k == key) return __ldg(vals + slot);
    if (k == 0ull) return 0;
    if (++slot == cap) slot = 0;
  }
  return 0;
}

__device__ __forceinline__ int resolve_n(const int* n_dev, int n_cap) {
  if (n_dev == nullptr) return n_cap;
  int n = *n_dev;
  return n < n_cap ? n : n_cap;
}

}  // namespace spvd
