// The small kernels around the fused head kernel: label-count prologue, finalize (with the sharded step's stats
// all-reduce over peer memory) and the deferred-mode finish kernel.  Included by head.cu only (product build) and by
// the CPU SIMT emulation harness (tests/cpu_simt, SIMT_CPU_EMULATION).
#pragma once
#include "step_xchg.cuh"

namespace simt {

// valid = a class id below C that is not the ignore label (the main kernel's rule exactly)
template <typename LabelT>
__device__ __forceinline__ unsigned long long count_valid(const LabelT* __restrict__ labels, long long npix, int C,
                                                          int ignore, long long i0, long long stride) {
  unsigned long long cnt = 0;
  if (sizeof(LabelT) == 1) {
    const int ign8 = (ignore >= 0 && ignore <= 255) ? ignore : 256;
    const uint8_t* lb = reinterpret_cast<const uint8_t*>(labels);
    const long long n16 = ((reinterpret_cast<uintptr_t>(lb) & 15) == 0) ? (npix >> 4) : 0;
    const uint4* l4 = reinterpret_cast<const uint4*>(lb);
    for (long long i = i0; i < n16; i += stride) {
      const uint4 v = ldg_stream_u4(l4 + i);
      const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = (int)((wds[k] >> (8 * q)) & 0xffu);
          cnt += (unsigned)((c < C) & (c != ign8));
        }
    }
    for (long long i = n16 * 16 + i0; i < npix; i += stride) {
      const int c = (int)__ldg(lb + i);
      cnt += (unsigned)((c < C) & (c != ign8));
    }
  } else {
    const long long* lb = reinterpret_cast<const long long*>(labels);
    for (long long i = i0; i < npix; i += stride) {
      const long long y = __ldg(lb + i);
      cnt += (unsigned)((y >= 0) & (y < (long long)C) & (y != (long long)ignore));
    }
  }
  return cnt;
}

// Step prologue (MODE_STEP): zero dLogits and count this rank's valid pixels in ONE pass over the labels, so that the
// main kernel can apply grad_out / N_valid itself.  The last block to finish publishes the count.
// Sharded step, additionally:
//   * the last block pushes this rank's count for this step into every mailbox as one tagged word -- unless the
//     previous step already did (`next_labels`), in which case it only checks that the labels are the announced ones;
//   * with `next_labels` it also counts the NEXT step's labels; the fused kernel pushes that count one step early, so
//     that no rank ever waits for a count.
template <typename LabelT>
__global__ void __launch_bounds__(256) head_prep_kernel(float* __restrict__ dlogits, long long n_dl,
                                                         const LabelT* __restrict__ labels,
                                                         const LabelT* __restrict__ next_labels, long long npix, int C,
                                                         int ignore, unsigned long long* __restrict__ ws,
                                                         const XchgArgs X, const FinishArgs F) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // ---- zero dLogits ----
  const long long n4 = ((reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) ? (n_dl >> 2) : 0;
  float4* d4 = reinterpret_cast<float4*>(dlogits);
  for (long long i = i0; i < n4; i += stride) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = n4 * 4 + i0; i < n_dl; i += stride) dlogits[i] = 0.f;
  // ---- count valid labels (of this step and, when announced, of the next) ----
  unsigned long long cnt = count_valid<LabelT>(labels, npix, C, ignore, i0, stride);
  unsigned long long cnt2 = next_labels ? count_valid<LabelT>(next_labels, npix, C, ignore, i0, stride) : 0ULL;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    cnt2 += __shfl_xor_sync(0xffffffffu, cnt2, o);
  }
  __shared__ unsigned long long s_w[8], s_w2[8];
  if ((threadIdx.x & 31) == 0) { s_w[threadIdx.x >> 5] = cnt; s_w2[threadIdx.x >> 5] = cnt2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long b = 0, b2 = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { b += s_w[k]; b2 += s_w2[k]; }
    atomicAdd(ws + WS_ACCUM, b);
    if (next_labels) atomicAdd(ws + WS_ACCUM_NEXT, b2);
    __threadfence();
    const unsigned long long t = atomicAdd(ws + WS_TICKET, 1ULL);
    if (t == (unsigned long long)gridDim.x - 1ULL) {      // last block: every partial is in, every block is done
      const unsigned long long total = atomicAdd(ws + WS_ACCUM, 0ULL);
      const unsigned long long total_next = atomicAdd(ws + WS_ACCUM_NEXT, 0ULL);
      ws[WS_ACCUM] = 0ULL;
      ws[WS_ACCUM_NEXT] = 0ULL;
      ws[WS_TICKET] = 0ULL;
      *reinterpret_cast<double*>(ws + WS_COUNT_LOCAL) = (double)total;
      if (X.world > 1) {
        unsigned char* own = X.mail[X.rank];
        const unsigned long long seq = step_seq(own);
        const unsigned long long mine = ld_relaxed_sys(count_slot_of(own, seq, X.rank));
        if ((mine >> 40) == (seq & 0xffffffULL)) {
          // announced one step ago: the labels must be the ones that were counted then
          if ((mine & kCountMask) != (total & kCountMask) && F.err) atomicOr(F.err, SIMT_ERRBIT_NEXT_LABELS);
        } else {
          const unsigned long long word = count_word(seq, total);
          for (int r = 0; r < X.world; ++r) st_relaxed_sys(count_slot_of(X.mail[r], seq, X.rank), word);
        }
        // the NEXT step's count is pushed by the fused kernel's prologue (its peer stores then complete behind ~90 us
        // of arithmetic instead of holding up this short kernel's retirement)
        ws[WS_COUNT_NEXT] = next_labels ? count_word(seq + 1ULL, total_next) : 0ULL;
      }
    }
  }
}

// Sharded, deferred mode: finish everything that is outstanding now (HeadRunner.finish()).  ORDER: first the pending
// steps (oldest first), THEN push the last step's stats (which may not have left this rank yet: no fused kernel ran
// since) and reduce it.  A rank may overwrite its peers' copies of stats(u - 2) with stats(u) only after it has read
// every peer's stats(u - 1) -- a peer pushes those only after it has finished reading stats(u - 2)
// (tests/test_xchg_protocol_cpu.py checks this order, and that the opposite one loses words).  Collective in spirit:
// the peers' stats of the last step only arrive once they run their next step or this kernel.
__global__ void __launch_bounds__(256) head_finish_kernel(unsigned long long* __restrict__ ws, const XchgArgs X,
                                                           const FinishArgs F) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const unsigned long long p0 = *reinterpret_cast<volatile unsigned long long*>(ws + WS_PENDING);
  const unsigned long long p1 = *reinterpret_cast<volatile unsigned long long*>(ws + WS_PENDING_ODD);
  const unsigned long long unsent = *reinterpret_cast<volatile unsigned long long*>(ws + WS_UNSENT);
  unsigned long long todo[2] = {p0, p1};
  if (todo[0] > todo[1]) { const unsigned long long t = todo[0]; todo[0] = todo[1]; todo[1] = t; }
  for (int q = 0; q < 2; ++q)
    if (todo[q] != 0ULL) finish_pending(X, F, todo[q], i);
  if (unsent != 0ULL) {
    push_stats_value(X, reinterpret_cast<const double*>(ws + kWsHeader / 8), F.C, F.CKP, unsent, i);
    finish_pending(X, F, unsent, i);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(ws + WS_TICKET, 1ULL);
    if (t == (unsigned long long)gridDim.x - 1ULL) {
      ws[WS_TICKET] = 0ULL;
      ws[WS_PENDING] = 0ULL;
      ws[WS_PENDING_ODD] = 0ULL;
      ws[WS_UNSENT] = 0ULL;
    }
  }
}

// Fixed-order reduction of the per-CTA partials.  blockDim = (32 outputs, 32 slices of the CTA range):
// consecutive threads read consecutive tile entries (coalesced); every slice first issues ALL its loads
// (independent, many in flight), sums them in order, then re-zeroes the entries for the next call; the
// 32 slice sums are added in order.  The last block reduces loss / count and re-arms the unit scheduler.
static constexpr int kFinSlices = 32;
static constexpr int kFinMaxPer = 8;   // tiles per slice held in registers: ntiles (= SM count) <= 32 * 8 = 256

__global__ void __launch_bounds__(1024) head_finalize_kernel(
    float* __restrict__ part_dT, const double* __restrict__ part_loss, const long long* __restrict__ part_cnt,
    int nparts, int ntiles, int CK, int CKP, int C, int mode, float gscale, unsigned long long* __restrict__ counter,
    double* __restrict__ stats, float* __restrict__ loss_mean, float* __restrict__ dT_out, int* __restrict__ err,
    const float* __restrict__ grad_out, const double* __restrict__ count_dev, unsigned long long* __restrict__ ws,
    const XchgArgs X, int defer) {
  const int ndt = C * CKP;
  const bool sharded = X.world > 1;   // loss / dT are final only after the exchange at the end of this kernel
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  __shared__ double sm[kFinSlices][33];
  __shared__ long long smi[kFinSlices];
  if ((int)blockIdx.x < (int)gridDim.x - 1) {
    const int o = blockIdx.x * 32 + tx;  // output index in the [y][k] layout of the tiles
    double s = 0.0;
    if (o < ndt && (mode == MODE_FWDBWD || mode == MODE_BWD || mode == MODE_STEP)) {
      float v[kFinMaxPer];
#pragma unroll
      for (int q = 0; q < kFinMaxPer; ++q) {
        const int g = ty + q * kFinSlices;
        v[q] = (g < ntiles) ? part_dT[(size_t)g * ndt + o] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < kFinMaxPer; ++q) s += (double)v[q];
#pragma unroll
      for (int q = 0; q < kFinMaxPer; ++q) {
        const int g = ty + q * kFinSlices;
        if (g < ntiles) part_dT[(size_t)g * ndt + o] = 0.f;
      }
    }
    sm[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && o < ndt) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < kFinSlices; ++q) t += sm[q][tx];
      const int y = o / CKP, k = o - y * CKP;
      sm[0][tx] = -t;   // column tx is read by this thread only: the block's other warps push this value below
      if (k < CK) {
        if (stats) stats[2 + k * C + y] = -t;
        if (defer) reinterpret_cast<double*>(ws + kWsHeader / 8)[2 + k * C + y] = -t;   // staging: awaits its push
        // MODE_STEP on one GPU: grad_out / N_valid is already known on the device (count pass)
        const double sc = count_dev ? (grad_out ? (double)__ldg(grad_out) : 1.0) / *count_dev : (double)gscale;
        if (dT_out && !sharded) dT_out[k * C + y] = (float)(-t * sc);
      }
    }
  } else {
    double l = 0.0;
    long long c = 0;
    for (int g = threadIdx.x; g < nparts; g += blockDim.x) { l += part_loss[g]; c += part_cnt[g]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (tx == 0) { sm[ty][0] = l; smi[ty] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      l = 0.0; c = 0;
      for (int q = 0; q < kFinSlices; ++q) { l += sm[q][0]; c += smi[q]; }
      *counter = 0ULL;  // the main kernel of this call has finished: re-arm the unit scheduler
      const double ls = -kLn2 * l;
      if (stats) { stats[0] = ls; stats[1] = (double)c; }
      if (defer) { double* stg = reinterpret_cast<double*>(ws + kWsHeader / 8); stg[0] = ls; stg[1] = (double)c; }
      if (loss_mean && !sharded) {
        float m = (float)(ls / (double)c);  // 0/0 -> NaN like the reference's mean over nothing
        if (err && (*err & SIMT_ERRBIT_LABEL_RANGE)) m = nanf("");
        *loss_mean = m;
      }
    }
  }
  if (!sharded) return;
  // ---- sharded step: all-reduce of `stats` over peer memory ----------------------------------------------------
  // Every block pushes the stats entries it has just written into slot [parity][rank] of every mailbox as tagged
  // words (peer stores over NVLink), then polls the same entries of all `world` slots of its OWN mailbox, sums them
  // in rank order (so the reduced values are bitwise identical on every rank) and writes the final stats / dT / loss.
  // No block waits for another block; the last one to finish advances the step counter.
  // Deferred mode: nothing crosses the ranks here (peer stores would hold up this short kernel's retirement); the
  // local stats stay in the caller's buffer, the next fused kernel's prologue pushes them and the prologue of the step
  // after that (or head_finish_kernel) reduces them -- no rank ever waits, no short kernel ever issues a peer store.
  __shared__ int s_bad;
  unsigned char* own = X.mail[X.rank];
  const unsigned long long seq = step_seq(own);
  const int par = (int)(seq & 1ULL);
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  __syncthreads();   // this block's local stats entries are written (and s_bad is initialised)
  const float poison = nanf("");
  // the global valid count: summed by the main kernel from the ranks' count words
  const double cnt = *reinterpret_cast<const double*>(ws + WS_COUNT_GLOBAL);
  const double sc = (grad_out ? (double)__ldg(grad_out) : 1.0) / cnt;
  if ((int)blockIdx.x < (int)gridDim.x - 1) {
    const int o = blockIdx.x * 32 + tx;
    const int y = o / CKP, k = o - y * CKP;
    const bool mine = o < ndt && k < CK;
    const int i = 2 + k * C + y;     // index in the caller's stats buffer; the slot keeps the tile order (entry 2 + o)
    // warp r (ty = r < world) handles rank r: it pushes this block's entries into rank r's mailbox and collects rank
    // r's entries from this rank's own mailbox, so the `world` polls run concurrently in different warps (one after
    // the other they cost ~0.7 us per system-scope load and rank)
    if (mine && ty < X.world && !defer) {
      ll_push_f64(slot_of(X.mail[ty], par, X.rank, X.slot_entries), 2 + o, seq, sm[0][tx]);   // (not stats[i]: warp 0 overwrites it)
      double v = 0.0;
      if (!ll_wait_f64(slot_of(own, par, ty, X.slot_entries), 2 + o, seq, X.max_spins, &v)) s_bad = 1;
      sm[1 + ty][tx] = v;
    }
    __syncthreads();
    if (mine && ty == 0 && !defer) {
      double t = 0.0;
      for (int r = 0; r < X.world; ++r) t += sm[1 + r][tx];   // rank order: bitwise identical on every rank
      const bool ok = s_bad == 0;
      stats[i] = ok ? t : (double)poison;
      if (dT_out) dT_out[i - 2] = ok ? (float)(t * sc) : poison;
    }
  } else {
    // {loss sum, valid count}: thread 2 r + e handles entry e of rank r
    if (tid < 2 * X.world && !defer) {
      const int e = tid & 1, r = tid >> 1;
      ll_push_f64(slot_of(X.mail[r], par, X.rank, X.slot_entries), e, seq, stats[e]);
      double v = 0.0;
      if (!ll_wait_f64(slot_of(own, par, r, X.slot_entries), e, seq, X.max_spins, &v)) s_bad = 1;
      sm[1 + r][e] = v;
    }
    __syncthreads();
    if (tid == 0 && !defer) {
      double t[2] = {0.0, 0.0};
      for (int e = 0; e < 2; ++e)
        for (int r = 0; r < X.world; ++r) t[e] += sm[1 + r][e];
      const bool ok = s_bad == 0;
      stats[0] = ok ? t[0] : (double)poison;
      stats[1] = ok ? t[1] : (double)poison;
      if (loss_mean) {
        float m = (float)(t[0] / t[1]);   // 0/0 -> NaN like the reference's mean over nothing
        if (!ok || (err && (*err & SIMT_ERRBIT_LABEL_RANGE))) m = poison;
        *loss_mean = m;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (s_bad && err) atomicOr(err, SIMT_ERRBIT_XCHG_TIMEOUT);
    __threadfence();
    const unsigned long long t = atomicAdd(ws + WS_FIN_TICKET, 1ULL);
    if (t == (unsigned long long)gridDim.x - 1ULL) {   // every block is done with this step's slots
      ws[WS_FIN_TICKET] = 0ULL;
      if (defer) ws[WS_UNSENT] = seq;   // the next fused kernel's prologue (or head_finish_kernel) pushes the stats
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(own) = seq;   // the step is over: advance the counter
    }
  }
}

}  // namespace simt
