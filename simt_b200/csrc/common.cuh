// Shared helpers for the simt_b200 kernels (sm_100a only).
#pragma once
#ifdef SIMT_CPU_EMULATION
// tests/cpu_simt: the device code of this library compiled with g++ against a CUDA-on-CPU shim (test infrastructure).
// The shim supplies the CUDA built-ins and host versions of the inline-PTX helpers below; the host-side helpers that
// call the CUDA runtime are left out.
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <math.h>
#include <mutex>
#include "../../include/simt_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "simt_b200 is written for sm_100a (B200) only"
#endif

#ifndef SIMT_CPU_EMULATION
#define SIMT_CUDA_TRY(expr)                     \
  do {                                          \
    cudaError_t e__ = (expr);                   \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

namespace simt {

struct DeviceInfo {
  int sm_count;
  int smem_optin;  // max dynamic shared memory per CTA (bytes)
};

// One mutex for the library's small per-device caches (attribute queries, function attributes).
inline std::mutex& cache_mutex() {
  static std::mutex m;
  return m;
}

// Per-device cache (the only global state besides the tuning hooks and the launch profiler; all mutex-guarded).
inline int device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  static bool have[64] = {};
  int dev = 0;
  SIMT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SIMT_EUNSUPPORTED;
  std::lock_guard<std::mutex> lock(cache_mutex());
  if (!have[dev]) {
    DeviceInfo d;
    SIMT_CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    SIMT_CUDA_TRY(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev] = d;
    have[dev] = true;
  }
  *out = cache[dev];
  return 0;
}

// Opt a kernel in to `bytes` of dynamic shared memory once per device (function attributes are per context, so a
// process that drives several GPUs must set them on each).
template <typename Kernel>
inline int ensure_dynamic_smem(Kernel kernel, bool (&done)[64], int bytes) {
  int dev = 0;
  SIMT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SIMT_EUNSUPPORTED;
  std::lock_guard<std::mutex> lock(cache_mutex());
  if (!done[dev]) {
    SIMT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done[dev] = true;
  }
  return 0;
}

}  // namespace simt
#endif  // !SIMT_CPU_EMULATION

namespace simt {

// ---- pixel <-> cell mapping, identical float arithmetic on host and device ----------
__host__ __device__ __forceinline__ float src_index(float scale, int X) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(scale, (float)X);  // no FMA contraction: torch rounds the product
#else
  volatile float r = scale * (float)X;
  return r;
#endif
}
__host__ __device__ __forceinline__ int cell_of(int X, float scale, int ncell) {
  int c = (int)src_index(scale, X);  // src >= 0: truncation == floor
  return c < ncell - 1 ? c : ncell - 1;
}
__host__ __device__ __forceinline__ float lambda_of(int X, float scale, int cell) {
  float l = src_index(scale, X) - (float)cell;
  l = l < 0.f ? 0.f : l;
  return l > 1.f ? 1.f : l;
}
// first output index whose cell is >= c  (monotone in c; 0 for c<=0, OUT for c>=ncell)
__host__ __device__ inline int first_px_of_cell(int c, float scale, int ncell, int OUT) {
  if (c <= 0) return 0;
  if (c >= ncell || !(scale > 0.f)) return OUT;
  float guess = ceilf((float)c / scale);
  int X = guess >= (float)OUT ? OUT : (int)guess;
  if (X < 0) X = 0;
  while (X > 0 && cell_of(X - 1, scale, ncell) >= c) --X;
  while (X < OUT && cell_of(X, scale, ncell) < c) ++X;
  return X;
}

#ifndef SIMT_CPU_EMULATION
// launch profiler (capi.cu)
bool prof_enabled();
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st);

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 128-bit streaming load that does not allocate in L1 (data is touched once).
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
#endif  // !SIMT_CPU_EMULATION (the shim has host versions of ex2 / lg2 / rcp / ldg_stream)

}  // namespace simt
