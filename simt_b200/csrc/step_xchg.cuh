// Device-side building blocks of the head STEP outside the fused kernel's pixel loop: modes, the workspace header, and
// everything the sharded step exchanges with its peers (count acquire, stats push / reduce).  The kernels that use
// them are in step_kernels.cuh (prologue, finalize, finish) and in head_kernel.cuh (stepx_*.inc).  Included by
// head_kernel.cuh (product build) and, with SIMT_CPU_EMULATION defined, by the CPU SIMT emulation harness of
// tests/cpu_simt, which runs THIS code for several emulated ranks with delayed, reordered peer stores.
#pragma once
#include "xchg.cuh"

namespace simt {

// MODE_STEP: forward + backward with the 1/N_valid scale known ON THE DEVICE before the kernel starts (a label-only
// count pass), so dLogits leave the kernel final and there is no scale pass.  Single GPU only: sharded, the count
// would be a second rendezvous per step on top of the stats exchange (measured slower than scaling after ONE exchange).
// MODE_STEPX: MODE_STEP on one rank's shard of the batch (the prologue talks to the peers' mailboxes); its own
// instantiation, so that the single-GPU kernel carries none of that code.
enum { MODE_FWD = 0, MODE_FWDBWD = 1, MODE_BWD = 2, MODE_PLACE = 3, MODE_STEP = 4, MODE_STEPX = 5 };

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr double kLn2 = 0.6931471805599453094;

// What the kernels that FINISH a sharded step write (the all-reduced results of that step).
struct FinishArgs {
  double* stats;        // [2 + CK*C]
  float* loss_mean;
  float* dT;            // [CK*C] or null
  const float* grad_out;
  int* err;
  int CK, C, CKP;       // CKP: row length of the kernel's dT tile (the order the values travel in)
};

// Workspace header (the first kWsHeader bytes of the caller's workspace), u64 words:
//   [0] unit scheduler counter  [1] count accumulator  [2] prep ticket  [3] count_local (f64)  [4] finalize ticket
//   [5] count_global (f64, sharded)  [6], [10] pending steps of even / odd parity: stats pushed, reduction outstanding
//   (sharded, deferred mode)  [7] count accumulator of next_labels  [8] unsent step: local stats not pushed yet
//   [9] tagged count word of the next step, to be pushed by the fused kernel
enum { WS_COUNTER = 0, WS_ACCUM = 1, WS_TICKET = 2, WS_COUNT_LOCAL = 3, WS_FIN_TICKET = 4, WS_COUNT_GLOBAL = 5,
       WS_PENDING = 6, WS_ACCUM_NEXT = 7, WS_UNSENT = 8, WS_COUNT_NEXT = 9, WS_PENDING_ODD = 10 };
__host__ __device__ __forceinline__ int ws_pending_word(unsigned long long step) { return (step & 1ULL) ? WS_PENDING_ODD : WS_PENDING; }
static constexpr size_t kWsHeader = 128;

// slot entry of value i of the caller's stats buffer: {loss, count} first, then dT[k][y] at 2 + y * CKP + k (the
// kernel's tile order: a warp's 32 values are contiguous)
__device__ __forceinline__ int stats_slot_entry(int i, int C, int CKP) {
  if (i < 2) return i;
  const int k = (i - 2) / C, y = (i - 2) - k * C;
  return 2 + y * CKP + k;
}
// push value i of this rank's local stats (step `seq`) into slot [parity][rank] of every mailbox
__device__ __forceinline__ void push_stats_value(const XchgArgs& X, const double* stats, int C, int CKP,
                                                 unsigned long long seq, int i) {
  if (i >= X.n_stats) return;
  const double v = stats[i];
  const int e = stats_slot_entry(i, C, CKP);
  for (int r = 0; r < X.world; ++r) ll_push_f64(slot_of(X.mail[r], (int)(seq & 1ULL), X.rank, X.slot_entries), e, seq, v);
}

// Sharded, deferred mode: the stats of step `pend` were pushed by every rank (from the prologue of its next fused
// kernel, or by its head_finish_kernel); wait for them in this rank's mailbox, sum in rank order (bitwise identical on
// every rank), write loss / dT / stats.  Value i of the caller's stats buffer.
// A peer that never arrives poisons the outputs with NaN and raises SIMT_ERRBIT_XCHG_TIMEOUT.
__device__ __forceinline__ void finish_pending(const XchgArgs& X, const FinishArgs& F, unsigned long long pend, int i) {
  if (i >= X.n_stats) return;
  unsigned char* own = X.mail[X.rank];
  const int par = (int)(pend & 1ULL);
  const int e = stats_slot_entry(i, F.C, F.CKP);
  // first look: every word of every rank in flight at once (they arrived long ago in the pipelined schedule)
  unsigned long long wv[kMaxPeers][2], wc[kMaxPeers][2];
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    if (r < X.world) {
      const unsigned long long* sl = slot_of(own, par, r, X.slot_entries);
      wv[r][0] = ld_relaxed_sys(sl + 2 * e); wv[r][1] = ld_relaxed_sys(sl + 2 * e + 1);
      wc[r][0] = ld_relaxed_sys(sl + 2); wc[r][1] = ld_relaxed_sys(sl + 3);
    }
  const unsigned long long tag = pend & 0xffffffffULL;
  double t = 0.0, cnt = 0.0;
  bool ok = true;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    if (r < X.world) {
      double v, c;
      if ((wv[r][0] >> 32) == tag && (wv[r][1] >> 32) == tag && (wc[r][0] >> 32) == tag && (wc[r][1] >> 32) == tag) {
        v = __longlong_as_double((long long)((wv[r][1] << 32) | (wv[r][0] & 0xffffffffULL)));
        c = __longlong_as_double((long long)((wc[r][1] << 32) | (wc[r][0] & 0xffffffffULL)));
      } else {   // not there yet: poll
        v = c = 0.0;
        ok = ll_wait_f64(slot_of(own, par, r, X.slot_entries), e, pend, X.max_spins, &v) && ok;
        ok = ll_wait_f64(slot_of(own, par, r, X.slot_entries), 1, pend, X.max_spins, &c) && ok;
      }
      t += v;
      cnt += c;
    }
  const float poison = nanf("");
  if (!ok && F.err) atomicOr(F.err, SIMT_ERRBIT_XCHG_TIMEOUT);
  if (F.stats) F.stats[i] = ok ? t : (double)poison;
  if (i >= 2 && F.dT && i - 2 < F.CK * F.C)
    F.dT[i - 2] = ok ? (float)(t * ((F.grad_out ? (double)__ldg(F.grad_out) : 1.0) / cnt)) : poison;
  if (i == 0 && F.loss_mean) {
    float m = (float)(t / cnt);   // 0/0 -> NaN like the reference's mean over nothing
    if (!ok || (F.err && (*F.err & SIMT_ERRBIT_LABEL_RANGE))) m = poison;
    *F.loss_mean = m;
  }
}

// Sharded step: wait (bounded) for every rank's valid-pixel count in this rank's mailbox and return
// grad_out / sum(counts); NaN + SIMT_ERRBIT_XCHG_TIMEOUT when a peer never arrives.  One thread per CTA, in the prologue.
static __device__ __forceinline__ float acquire_global_scale(const XchgArgs& X, const float* grad_out, int* err,
                                                             double* count_global, int lane) {
  unsigned char* own = X.mail[X.rank];
  const unsigned long long seq = step_seq(own);
  unsigned long long n = 0ULL;
  bool ok = true;
  if (lane < X.world) ok = wait_count(count_slot_of(own, seq, lane), seq, X.max_spins, &n);   // one lane per rank
  double c = (double)n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);   // integers: any order is exact
  ok = __all_sync(0xffffffffu, ok);
  if (!ok && err && lane == 0) atomicOr(err, SIMT_ERRBIT_XCHG_TIMEOUT);
  if (blockIdx.x == 0 && lane == 0 && count_global) *count_global = ok ? c : (double)nanf("");
  return ok ? (float)((grad_out ? (double)__ldg(grad_out) : 1.0) / c) : nanf("");
}

}  // namespace simt
