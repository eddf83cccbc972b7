// Mailbox layout and system-scope load/store helpers of the peer-memory exchange (csrc/xchg.cu).
//
// One mailbox per rank (cudaMalloc'd, CUDA-IPC mapped into every peer):
//   u64 hdr[32]:  [0] step counter   [1] block ticket of the scale kernel
//                 [8 + parity * 8 + rank]  stats flags   (step number once rank's stats have landed)
//   double data[2][8][n_stats]: stats slots
// Slots are double-buffered by step parity; see xchg.cu for why that is enough.
#pragma once
#include "common.cuh"

namespace simt {

static constexpr int kMaxPeers = 8;
static constexpr int kHdrWords = 32;
static constexpr size_t kHdrBytes = kHdrWords * sizeof(unsigned long long);
static constexpr int kHdrStatFlag = 8;
static constexpr long long kXchgMaxSpins = 1LL << 24;  // seconds, not the microseconds an exchange takes

struct XchgArgs {
  unsigned char* mail[kMaxPeers];  // mailbox base of every rank (mail[rank] is local memory); all null when world == 1
  int rank, world, n_stats;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long* hdr_of(unsigned char* mailbox) {
  return reinterpret_cast<unsigned long long*>(mailbox);
}
__device__ __forceinline__ double* slot_of(unsigned char* mailbox, int parity, int rank, int n_stats) {
  return reinterpret_cast<double*>(mailbox + kHdrBytes) + ((size_t)parity * kMaxPeers + rank) * n_stats;
}
// this step's number: the counter is advanced by the scale kernel, the last kernel of a step
__device__ __forceinline__ unsigned long long step_seq(unsigned char* own_mailbox) {
  return *reinterpret_cast<volatile unsigned long long*>(own_mailbox) + 1ULL;
}

// Bounded wait for flag >= seq; returns false on timeout.
__device__ __forceinline__ bool wait_flag(const unsigned long long* f, unsigned long long seq) {
  long long spins = 0;
  while (ld_acquire_sys(f) < seq) {
    if (++spins > kXchgMaxSpins) return false;
    __nanosleep(64);
  }
  return true;
}

}  // namespace simt
