// CUDA-on-CPU shim for the CPU SIMT emulation of the sharded step's exchange code (TEST INFRASTRUCTURE ONLY).
//
// csrc/xchg.cuh includes this file instead of common.cuh when SIMT_CPU_EMULATION is defined, so that g++ can compile
// the UNMODIFIED device code of csrc/xchg.cuh, step_xchg.cuh, step_kernels.cuh and stepx_*.inc.  Execution model:
//   * one OS thread per emulated rank (GPU); kernels of a rank run one after the other (one stream), the blocks of a
//     kernel one after the other (legal: none of these kernels lets a block wait for another block of its grid);
//   * every CUDA thread of a block is a ucontext fiber; __syncthreads / warp shuffles / votes are barriers among the
//     fibers; a fiber yields at every barrier, system-scope load and __nanosleep, and the scheduler picks the next
//     runnable fiber at random -- warps and lanes make progress in arbitrary order;
//   * st.relaxed.sys stores (the peer stores over NVLink) do not take effect at once: they sit in the rank's outbox and
//     are delivered later, in random order (program order is kept only between stores to the SAME address), some of them
//     only after the kernel that issued them has ended -- the receiver may see a step's words in any order and long
//     after words issued later.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <ucontext.h>

#include <cstdio>
#include <random>
#include <vector>

#include "../../include/simt_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { unsigned x, y, z, w; };
struct float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

namespace cpusimt {

struct Dim { unsigned x, y, z; };

struct Fiber {
  ucontext_t uc;
  bool done = false;
  Dim tid{0, 0, 0};
  int wait_kind = 0;            // 0 runnable, 1 waits at the block barrier, 2 waits at its warp's barrier
  unsigned wait_gen = 0;
};

struct Warp {
  unsigned long long buf[32];
  const void* site[32];         // call site of the collective every lane is in (divergent collectives are reported)
  int arrived = 0;
  unsigned gen = 0;
};

struct PendingStore { unsigned long long* p; unsigned long long v; };

// Per-rank (= per OS thread) emulator state.
struct Rank {
  std::mt19937_64 rng;
  ucontext_t sched;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::vector<char*> stacks;    // reused from launch to launch
  std::vector<unsigned char> dyn_smem;   // the running block's dynamic shared memory
  int cur = -1;                 // running fiber
  int live = 0;                 // fibers of the block that have not returned
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  Dim bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
  std::vector<PendingStore> outbox;   // peer stores in flight
  double p_deliver = 0.3;       // chance per scheduling point that some stores in flight arrive
  double p_flush_at_kernel_end = 0.5;
  unsigned long long drop_store_to = 0;   // fault injection: stores to this address never arrive
  unsigned long long n_switch = 0;
  void (*body)(void*) = nullptr;
  void* body_arg = nullptr;

  void deliver_one() {
    if (outbox.empty()) return;
    size_t i = (size_t)(rng() % outbox.size());
    for (size_t j = 0; j < i; ++j)          // program order between stores to the same address
      if (outbox[j].p == outbox[i].p) { i = j; break; }
    PendingStore s = outbox[i];
    outbox.erase(outbox.begin() + (long)i);
    if ((unsigned long long)(uintptr_t)s.p == drop_store_to) return;
    __atomic_store_n(s.p, s.v, __ATOMIC_RELAXED);
  }
  void maybe_deliver() {
    if (outbox.empty()) return;
    if ((double)(rng() % 1000) < p_deliver * 1000.0) {
      size_t n = 1 + (size_t)(rng() % 8);
      while (n-- && !outbox.empty()) deliver_one();
    }
  }
  void flush() { while (!outbox.empty()) deliver_one(); }
};

extern thread_local Rank* R;     // the emulator of the calling OS thread

// debugging aid: when set (by a harness' SIGALRM handler) the next fiber that yields prints its backtrace and exits
inline volatile int& debug_backtrace_flag() { static volatile int v = 0; return v; }
void debug_backtrace_and_exit();   // harness

inline void yield() {
  Rank* r = R;
  if (debug_backtrace_flag()) debug_backtrace_and_exit();
  ++r->n_switch;
  swapcontext(&r->fibers[(size_t)r->cur].uc, &r->sched);
}

inline void fiber_main() {
  Rank* r = R;
  r->body(r->body_arg);
  Fiber& f = r->fibers[(size_t)r->cur];
  f.done = true;
  --r->live;
  // a thread that has returned no longer takes part in the block's barriers
  if (r->live > 0 && r->bar_arrived == r->live) { r->bar_arrived = 0; ++r->bar_gen; }
  swapcontext(&f.uc, &r->sched);
}

// Run `body` for every thread of every block of a grid, blocks one after the other.
[[noreturn]] void die(const char* what);   // harness: report and _exit

inline bool runnable(Rank* r, const Fiber& f) {
  if (f.wait_kind == 1) return r->bar_gen != f.wait_gen;
  if (f.wait_kind == 2) return r->warps[f.tid.x >> 5].gen != f.wait_gen;
  return true;
}

static constexpr size_t kStackBytes = 32 * 1024;

inline unsigned char* dynamic_smem() {
  // 16-byte aligned start inside the block's buffer
  unsigned char* p = R->dyn_smem.data();
  return p + ((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15);
}

inline void launch(unsigned grid, unsigned block, void (*body)(void*), void* arg, size_t dyn_smem_bytes = 0) {
  Rank* r = R;
  while (r->stacks.size() < block) r->stacks.push_back(new char[kStackBytes]);
  r->body = body; r->body_arg = arg;
  r->gdim = Dim{grid, 1, 1}; r->bdim = Dim{block, 1, 1};
  for (unsigned b = 0; b < grid; ++b) {
    r->bid = Dim{b, 0, 0};
    // shared memory starts out as garbage on the GPU: fill it with a NaN pattern so that a read of an unwritten word
    // shows up in the results
    r->dyn_smem.assign(dyn_smem_bytes + 16, (unsigned char)0xFF);
    r->fibers.assign(block, Fiber{});
    r->warps.assign((block + 31) / 32, Warp{});
    r->live = (int)block; r->bar_arrived = 0;
    for (unsigned t = 0; t < block; ++t) {
      Fiber& f = r->fibers[t];
      f.tid = Dim{t, 0, 0};
      getcontext(&f.uc);
      f.uc.uc_stack.ss_sp = r->stacks[t];
      f.uc.uc_stack.ss_size = kStackBytes;
      f.uc.uc_link = nullptr;
      makecontext(&f.uc, (void (*)())fiber_main, 0);
    }
    std::vector<int> alive(block);
    for (unsigned t = 0; t < block; ++t) alive[t] = (int)t;
    while (!alive.empty()) {
      // random scheduling: a random runnable fiber runs until its next yield
      size_t k = (size_t)(r->rng() % alive.size());
      size_t probes = 0;
      while (!runnable(r, r->fibers[(size_t)alive[k]])) {
        k = (k + 1) % alive.size();
        if (++probes > alive.size()) die("every live thread of the block waits at a barrier (intra-block deadlock)");
      }
      r->cur = alive[k];
      swapcontext(&r->sched, &r->fibers[(size_t)r->cur].uc);
      if (r->fibers[(size_t)r->cur].done) { alive[k] = alive.back(); alive.pop_back(); }
      r->maybe_deliver();
    }
  }
  if ((double)(r->rng() % 1000) < r->p_flush_at_kernel_end * 1000.0) r->flush();
}

inline void block_barrier() {
  Rank* r = R;
  const unsigned gen = r->bar_gen;
  if (++r->bar_arrived == r->live) { r->bar_arrived = 0; ++r->bar_gen; return; }
  Fiber& f = r->fibers[(size_t)r->cur];
  f.wait_kind = 1; f.wait_gen = gen;
  while (r->bar_gen == gen) yield();
  r->fibers[(size_t)r->cur].wait_kind = 0;
}

inline Warp& my_warp() { Rank* r = R; return r->warps[(size_t)(r->fibers[(size_t)r->cur].tid.x >> 5)]; }
inline int my_lane() { Rank* r = R; return (int)(r->fibers[(size_t)r->cur].tid.x & 31); }
inline int warp_width() {
  Rank* r = R;
  const unsigned w = r->fibers[(size_t)r->cur].tid.x >> 5;
  const unsigned left = r->bdim.x - w * 32;
  return (int)(left < 32 ? left : 32);
}
// all lanes of the warp (full mask: the emulated kernels only use 0xffffffff collectives in converged code).
// `site` != null: the first barrier of a collective -- every lane must have come from the SAME call site, otherwise the
// kernel executes a full-mask collective in divergent code (undefined behaviour in CUDA) and the emulation stops.
inline void warp_barrier(const void* site = nullptr) {
  const unsigned w = R->fibers[(size_t)R->cur].tid.x >> 5;
  const unsigned gen = R->warps[w].gen;
  if (site) R->warps[w].site[my_lane()] = site;
  if (++R->warps[w].arrived == warp_width()) {
    if (site)
      for (int l = 1; l < warp_width(); ++l)
        if (R->warps[w].site[l] != R->warps[w].site[0]) {
          fprintf(stderr, "divergent warp collective: block %u warp %u lane 0 at %p, lane %d at %p\n", R->bid.x, w,
                  R->warps[w].site[0], l, R->warps[w].site[l]);
          die("lanes of one warp are in different full-mask collectives");
        }
    R->warps[w].arrived = 0; ++R->warps[w].gen; return;
  }
  Fiber& f = R->fibers[(size_t)R->cur];
  f.wait_kind = 2; f.wait_gen = gen;
  while (R->warps[w].gen == gen) yield();
  R->fibers[(size_t)R->cur].wait_kind = 0;
}

}  // namespace cpusimt

#define threadIdx (cpusimt::R->fibers[(size_t)cpusimt::R->cur].tid)
#define blockIdx (cpusimt::R->bid)
#define blockDim (cpusimt::R->bdim)
#define gridDim (cpusimt::R->gdim)

static inline void __syncthreads() { cpusimt::block_barrier(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { cpusimt::yield(); }

#define SIMT_COLLECTIVE __attribute__((noinline)) static
template <typename T>
SIMT_COLLECTIVE T __shfl_xor_sync(unsigned, T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  const int lane = cpusimt::my_lane();
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  { cpusimt::Warp& w = cpusimt::my_warp(); w.buf[lane] = bits; }
  cpusimt::warp_barrier(__builtin_return_address(0));
  { cpusimt::Warp& w = cpusimt::my_warp(); bits = w.buf[(lane ^ o) & 31]; }
  cpusimt::warp_barrier();
  T r;
  memcpy(&r, &bits, sizeof(T));
  return r;
}
template <typename T>
SIMT_COLLECTIVE T __shfl_sync(unsigned, T v, int src) {
  const int lane = cpusimt::my_lane();
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  { cpusimt::Warp& w = cpusimt::my_warp(); w.buf[lane] = bits; }
  cpusimt::warp_barrier(__builtin_return_address(0));
  { cpusimt::Warp& w = cpusimt::my_warp(); bits = w.buf[src & 31]; }
  cpusimt::warp_barrier();
  T r;
  memcpy(&r, &bits, sizeof(T));
  return r;
}
template <typename T>
SIMT_COLLECTIVE T __shfl_up_sync(unsigned, T v, int delta) {
  const int lane = cpusimt::my_lane();
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  { cpusimt::Warp& w = cpusimt::my_warp(); w.buf[lane] = bits; }
  cpusimt::warp_barrier(__builtin_return_address(0));
  { cpusimt::Warp& w = cpusimt::my_warp(); bits = w.buf[lane >= delta ? lane - delta : lane]; }
  cpusimt::warp_barrier();
  T r;
  memcpy(&r, &bits, sizeof(T));
  return r;
}
SIMT_COLLECTIVE void __syncwarp(unsigned = 0xffffffffu) { cpusimt::warp_barrier(__builtin_return_address(0)); }
SIMT_COLLECTIVE bool __any_sync(unsigned, bool p) {
  const int lane = cpusimt::my_lane();
  { cpusimt::Warp& w = cpusimt::my_warp(); w.buf[lane] = p ? 1ULL : 0ULL; }
  cpusimt::warp_barrier(__builtin_return_address(0));
  bool any = false;
  { cpusimt::Warp& w = cpusimt::my_warp(); for (int l = 0; l < cpusimt::warp_width(); ++l) any = any || (w.buf[l] != 0ULL); }
  cpusimt::warp_barrier();
  return any;
}
SIMT_COLLECTIVE int __reduce_max_sync(unsigned, int v) {
  const int lane = cpusimt::my_lane();
  { cpusimt::Warp& w = cpusimt::my_warp(); w.buf[lane] = (unsigned long long)(long long)v; }
  cpusimt::warp_barrier(__builtin_return_address(0));
  int m = v;
  { cpusimt::Warp& w = cpusimt::my_warp(); for (int l = 0; l < cpusimt::warp_width(); ++l) { const int x = (int)(long long)w.buf[l]; m = x > m ? x : m; } }
  cpusimt::warp_barrier();
  return m;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (shift & 31u));
}
static inline float atomicAdd(float* p, float v) { const float old = *p; *p = old + v; return old; }   // one OS thread per rank
SIMT_COLLECTIVE bool __all_sync(unsigned, bool p) {
  const int lane = cpusimt::my_lane();
  { cpusimt::Warp& w = cpusimt::my_warp(); w.buf[lane] = p ? 1ULL : 0ULL; }
  cpusimt::warp_barrier(__builtin_return_address(0));
  bool all = true;
  { cpusimt::Warp& w = cpusimt::my_warp(); for (int l = 0; l < cpusimt::warp_width(); ++l) all = all && (w.buf[l] != 0ULL); }
  cpusimt::warp_barrier();
  return all;
}

static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }

namespace simt {
static inline uint4 ldg_stream_u4(const uint4* p) { return *p; }
// host stand-ins for MUFU (ex2 / lg2 / rcp .approx.ftz): correctly rounded libm values instead of the ~2 ulp approximations
static inline float ftz(float x) { return fabsf(x) < 1.17549435e-38f ? copysignf(0.f, x) : x; }
static inline float ex2_approx(float x) { return ftz(exp2f(ftz(x))); }
static inline float lg2_approx(float x) { x = ftz(x); return x == 0.f ? -INFINITY : log2f(x); }
static inline float rcp_approx(float x) { return ftz(1.0f / ftz(x)); }
}  // namespace simt
