// Pipe-rate micro-benchmarks for the head kernel's design decisions (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench.bin scripts/microbench.cu
// Prints warp-instructions per clock per SM for: FFMA, FFMA2 (fma.rn.f32x2), MUFU.EX2, SHFL, LDS.64, and mixes.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void ffma2(float& x, float& y, float a, float b, float c, float d) {
  asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rc, {%4, %5};\n\t"
               "fma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%0, %1}, ra;\n\t}"
               : "+f"(x), "+f"(y) : "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int NCH = 8;     // independent chains per thread
constexpr int ITER = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int sel) {
  __shared__ float2 sm[256 * 2];
  float x[NCH], y[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = x[i] + 0.5f; }
  sm[threadIdx.x] = make_float2(a, b);
  sm[threadIdx.x + 256] = make_float2(b, a);
  __syncthreads();
  int iacc = threadIdx.x;
  const long long t_begin = clock64();
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {  // scalar FFMA: 2 * NCH per iteration
#pragma unroll
      for (int i = 0; i < NCH; ++i) { x[i] = fmaf(x[i], a, b); y[i] = fmaf(y[i], a, b); }
    } else if (MODE == 1) {  // FFMA2: NCH per iteration (2 * NCH fmas)
#pragma unroll
      for (int i = 0; i < NCH; ++i) ffma2(x[i], y[i], a, a, b, b);
    } else if (MODE == 2) {  // MUFU.EX2: NCH per iteration
#pragma unroll
      for (int i = 0; i < NCH; ++i) x[i] = ex2(x[i]);
    } else if (MODE == 3) {  // payload mix: 7 FFMA2 : 2 MUFU (per channel pair per pixel)
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        ffma2(x[i], y[i], a, a, b, b);
        float e0 = ex2(x[i]), e1 = ex2(y[i]);
        ffma2(x[i], y[i], e0, e1, b, b);
        ffma2(x[i], y[i], a, a, b, b);
        ffma2(x[i], y[i], a, a, e0, e1);
        ffma2(x[i], y[i], a, a, b, b);
        ffma2(x[i], y[i], a, a, b, b);
        ffma2(x[i], y[i], a, a, b, b);
      }
    } else if (MODE == 4) {  // SHFL: NCH per iteration
#pragma unroll
      for (int i = 0; i < NCH; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1) + a;
    } else if (MODE == 5) {  // LDS.64: NCH per iteration
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        float2 v = sm[(threadIdx.x + (i & 1) * 256 + it * sel + iacc) & 511];
        iacc ^= __float_as_int(v.x) & sel;
        x[i] += v.x; y[i] += v.y;
      }
    } else if (MODE == 6) {  // FFMA2 + integer ALU side work (1:1)
#pragma unroll
      for (int i = 0; i < NCH; ++i) { ffma2(x[i], y[i], a, a, b, b); iacc = (iacc ^ (iacc >> 3)) + sel; }
    } else if (MODE == 7) {  // FFMA2 + scalar FADD side work (1:1)
#pragma unroll
      for (int i = 0; i < NCH; ++i) { ffma2(x[i], y[i], a, a, b, b); }
#pragma unroll
      for (int i = 0; i < NCH; ++i) { x[i] += b; }
    } else if (MODE == 8) {  // FMUL2-like geometric recurrence: x *= a (packed), NCH per iteration
#pragma unroll
      for (int i = 0; i < NCH; ++i) ffma2(x[i], y[i], a, a, 0.f, 0.f);
    } else if (MODE == 9) {  // MUFU + FFMA2 1:1 (is MUFU issue free next to the FMA pipe?)
#pragma unroll
      for (int i = 0; i < NCH; ++i) { ffma2(x[i], y[i], a, a, b, b); }
#pragma unroll
      for (int i = 0; i < NCH / 2; ++i) { x[i] = ex2(x[i]); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += x[i] + y[i];
  if (s == 12345.678f || iacc == 0x7fffffff) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<long long*>(out)[1] = clock64() - t_begin;
}

template <int MODE>
int run(const char* name, double inst_per_iter, int sm, double mhz, int warps_per_sm) {
  float* out;
  CK(cudaMalloc(&out, 16));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = 256;
  const int ctas = sm * warps_per_sm * 32 / threads;
  k<MODE><<<ctas, threads>>>(out, 1.0001f, 1e-7f, 0);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<ctas, threads>>>(out, 1.0001f, 1e-7f, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double winst = (double)ctas * (threads / 32) * ITER * inst_per_iter;
  long long cyc = 0;
  cudaMemcpy(&cyc, reinterpret_cast<long long*>(out) + 1, 8, cudaMemcpyDeviceToHost);
  const double clk = (double)cyc;   // SM cycles of CTA 0 (the grid is one wave: every CTA runs the whole time)
  (void)mhz;
  printf("%-34s warps/SM %2d  %8.3f ms %9lld cyc  %6.3f warp-inst/clk/SM (%.3f per scheduler)\n", name, warps_per_sm, best, cyc,
         winst / clk / sm, winst / clk / sm / 4);
  cudaFree(out);
  return 0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  printf("%s, %d SMs, clock %0.f MHz (rates assume this clock)\n", p.name, p.multiProcessorCount, mhz);
  for (int w : {8, 16, 32}) {
    run<0>("FFMA scalar", 2.0 * NCH, p.multiProcessorCount, mhz, w);
    run<1>("FFMA2 (f32x2)", 1.0 * NCH, p.multiProcessorCount, mhz, w);
    run<2>("MUFU.EX2", 1.0 * NCH, p.multiProcessorCount, mhz, w);
    run<3>("mix 7 FFMA2 : 2 MUFU", 9.0 * NCH, p.multiProcessorCount, mhz, w);
    run<4>("SHFL (+FADD)", 2.0 * NCH, p.multiProcessorCount, mhz, w);
    run<5>("LDS.64 (+2 FADD)", 3.0 * NCH, p.multiProcessorCount, mhz, w);
    run<6>("FFMA2 + 2 int ALU", 3.0 * NCH, p.multiProcessorCount, mhz, w);
    run<7>("FFMA2 + FADD 1:1", 2.0 * NCH, p.multiProcessorCount, mhz, w);
    run<8>("FFMA2 as FMUL2", 1.0 * NCH, p.multiProcessorCount, mhz, w);
    run<9>("FFMA2 + MUFU 2:1", 1.5 * NCH, p.multiProcessorCount, mhz, w);
  }
  return 0;
}
