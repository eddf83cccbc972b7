// Mailbox layout and system-scope load/store helpers of the peer-memory exchange of the sharded head step.
//
// One mailbox per rank (cudaMalloc'd, CUDA-IPC mapped into every peer):
//   u64 hdr[64]:  [0] step counter (advanced by the last kernel of a step)
//   u64 counts[4][8]               valid-pixel counts of step s in row s % 4, pushed by the first kernel of step s (or,
//                                  when the caller knows the next step's labels, of step s - 1) as ONE tagged word
//                                  each: (step number mod 2^24) << 40 | count -- a single 8-byte store, no fence
//   u64 stats [2][8][2 * slot]     {loss sum, valid count, raw dT tile [C][CKP]} as doubles, pushed by the step's last
//                                  kernel as TWO tagged words per value: (step number mod 2^32) << 32 | 32 bits of the
//                                  double; slot = 2 + C * 64 entries (the dT tile keeps the kernel's [y][k] order, so a
//                                  warp's 32 values are one contiguous 512-byte peer store)
// Every word carries its own step tag (the scheme of NCCL's LL protocol): an 8-byte store is atomic, so the receiver
// polls the word itself -- no flag, no fence, one NVLink store latency per exchange.
// Stats slots are double-buffered by step parity, count rows by step mod 4; see xchg.cu for why that is enough.
#pragma once
#include "common.cuh"   // (with SIMT_CPU_EMULATION: the CUDA-on-CPU shim of tests/cpu_simt)

namespace simt {

static constexpr int kMaxPeers = 8;
static constexpr int kHdrWords = 64;
static constexpr size_t kHdrBytes = kHdrWords * sizeof(unsigned long long);
static constexpr int kCountRows = 4;
static constexpr int kXchgMaxCKP = 64;   // widest padded channel count of the head kernel's dT tile
static constexpr size_t kCountBytes = kCountRows * kMaxPeers * sizeof(unsigned long long);

struct XchgArgs {
  unsigned char* mail[kMaxPeers];  // mailbox base of every rank (mail[rank] is local memory); all null when world <= 1
  int rank, world, n_stats;        // n_stats = 2 + CK*C values of the caller's stats buffer
  int slot_entries;                // capacity of one stats slot (values): 2 + C * kXchgMaxCKP
  long long max_spins;             // bound of every flag wait (<= 0: wait for ever); see simt_xchg_set_timeout
};

__device__ __forceinline__ unsigned long long* hdr_of(unsigned char* mailbox) {
  return reinterpret_cast<unsigned long long*>(mailbox);
}
__device__ __forceinline__ unsigned long long* count_slot_of(unsigned char* mailbox, unsigned long long seq, int rank) {
  return reinterpret_cast<unsigned long long*>(mailbox + kHdrBytes) + (int)(seq & (kCountRows - 1)) * kMaxPeers + rank;
}
// tagged count word: the step number travels with the value, so one relaxed 8-byte store publishes both
static constexpr unsigned long long kCountMask = (1ULL << 40) - 1ULL;
__device__ __forceinline__ unsigned long long count_word(unsigned long long seq, unsigned long long count) {
  return ((seq & 0xffffffULL) << 40) | (count & kCountMask);
}
#ifdef SIMT_CPU_EMULATION
// emulator: the store joins the rank's outbox and arrives later, out of order; a load is a scheduling point
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  cpusimt::R->outbox.push_back(cpusimt::PendingStore{p, v});
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  const unsigned long long v = __atomic_load_n(p, __ATOMIC_RELAXED);
  cpusimt::yield();
  return v;
}
#else
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#endif
// Wait for the count word of step `seq`; false when the bound expires.
__device__ __forceinline__ bool wait_count(const unsigned long long* p, unsigned long long seq, long long max_spins,
                                           unsigned long long* count) {
  long long spins = 0;
  unsigned long long v;
  while (((v = ld_relaxed_sys(p)) >> 40) != (seq & 0xffffffULL)) {
    if (max_spins > 0 && ++spins > max_spins) return false;
    if (spins > 64) __nanosleep(32);
  }
  *count = v & kCountMask;
  return true;
}
__device__ __forceinline__ unsigned long long* slot_of(unsigned char* mailbox, int parity, int rank, int slot_entries) {
  return reinterpret_cast<unsigned long long*>(mailbox + kHdrBytes + kCountBytes) +
         ((size_t)parity * kMaxPeers + rank) * (size_t)(2 * slot_entries);
}
// this step's number: the counter is advanced by the last kernel of a step
__device__ __forceinline__ unsigned long long step_seq(unsigned char* own_mailbox) {
  return *reinterpret_cast<volatile unsigned long long*>(own_mailbox) + 1ULL;
}

// ---- stats values as two tagged words ------------------------------------------------------------------------------
__device__ __forceinline__ void ll_push_f64(unsigned long long* slot, int i, unsigned long long seq, double v) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long tag = (seq & 0xffffffffULL) << 32;
  st_relaxed_sys(slot + 2 * i, tag | (bits & 0xffffffffULL));
  st_relaxed_sys(slot + 2 * i + 1, tag | (bits >> 32));
}
// Wait for both words of value i of step `seq`; false when the bound expires (the caller poisons its outputs and
// raises the error bit: a rank must never continue with a partial sum).
__device__ __forceinline__ bool ll_wait_f64(const unsigned long long* slot, int i, unsigned long long seq,
                                            long long max_spins, double* v) {
  const unsigned long long tag = seq & 0xffffffffULL;
  long long spins = 0;
  unsigned long long lo, hi;
  while (((lo = ld_relaxed_sys(slot + 2 * i)) >> 32) != tag || ((hi = ld_relaxed_sys(slot + 2 * i + 1)) >> 32) != tag) {
    if (max_spins > 0 && ++spins > max_spins) return false;
    if (spins > 64) __nanosleep(32);
  }
  *v = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffULL)));
  return true;
}

}  // namespace simt
