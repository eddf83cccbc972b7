// Host side of the fused SimT head (entry points, plans, the small kernels around the fused one).  The fused kernel
// itself is in head_kernel.cuh; its T = NULL (plain CE) instantiations are compiled in head_ident.cu.
#include <cstdlib>
#include "head_kernel.cuh"

namespace simt {

long long xchg_max_spins();   // xchg.cu

// workspace: [header kWsHeader][staging: 2 + CK*C doubles (sharded, deferred mode: local stats awaiting their push)]
//            [part_loss f64 x G][part_cnt i64 x G][part_dT f32 x ntiles*C*CKP (sized for G tiles)]
static size_t ws_staging_bytes(int CK, int C) { return (((size_t)(2 + (size_t)CK * C) * 8) + 127) / 128 * 128; }

// head_ident.cu: the IDENT = true instantiations
int dispatch_modes_ident(int mode, int label_bytes, const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out);

static int dispatch_all(int mode, int label_bytes, const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
  if (mode != MODE_PLACE && A.T == nullptr) return dispatch_modes_ident(mode, label_bytes, A, P, st, grid_out);
  return dispatch_modes<false>(mode, label_bytes, A, P, st, grid_out);
}

// (workspace header words WS_*: head_kernel.cuh)

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
    if (mine && ty < X.world && !defer)
      ll_push_f64(slot_of(X.mail[ty], par, X.rank, X.slot_entries), 2 + o, seq, sm[0][tx]);   // (not stats[i]: warp 0 overwrites it)
    if (mine && ty == 0 && !defer) {
      double t = 0.0;
      bool ok = true;
      for (int r = 0; r < X.world; ++r) {
        double v = 0.0;
        ok = ll_wait_f64(slot_of(own, par, r, X.slot_entries), 2 + o, seq, X.max_spins, &v) && ok;
        t += v;
      }
      if (!ok) s_bad = 1;
      stats[i] = ok ? t : (double)poison;
      if (dT_out) dT_out[i - 2] = ok ? (float)(t * sc) : poison;
    }
  } else {
    if (tid < 2 * X.world && !defer)
      ll_push_f64(slot_of(X.mail[tid >> 1], par, X.rank, X.slot_entries), tid & 1, seq, stats[tid & 1]);
    __syncthreads();
    if (tid == 0 && !defer) {
      double t[2] = {0.0, 0.0};
      bool ok = true;
      for (int i = 0; i < 2; ++i)
        for (int r = 0; r < X.world; ++r) {
          double v = 0.0;
          ok = ll_wait_f64(slot_of(own, par, r, X.slot_entries), i, seq, X.max_spins, &v) && ok;
          t[i] += v;
        }
      if (!ok) s_bad = 1;
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

__global__ void head_scale_kernel(float* __restrict__ dlogits, long long n, const double* __restrict__ stats,
                                  int nT, const float* __restrict__ grad_out, float* __restrict__ dT) {
  const float s = (float)((grad_out ? (double)__ldg(grad_out) : 1.0) / stats[1]);
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = ((reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) ? (n >> 2) : 0;
  float4* d4 = reinterpret_cast<float4*>(dlogits);
  for (long long i = i0; i < n4; i += stride) {
    float4 v = d4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    d4[i] = v;
  }
  for (long long i = n4 * 4 + i0; i < n; i += stride) dlogits[i] *= s;
  if (dT)
    for (long long i = i0; i < nT; i += stride) dT[i] = (float)(stats[2 + i] * (double)s);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// workspace: [counter u64 (+pad to 64 B)][part_loss f64 x G][part_cnt i64 x G][part_dT f32 x ntiles*C*CKP] (sized for G tiles)

// Benchmark tuning (simt_head_set_tuning) and the launch caches are process-global; one mutex guards both, so the
// entry points may be called from several host threads (one stream / device each).
struct Tuning { int ur, unused, threads, lpr; };
static Tuning g_tuning = {0, 0, 0, 0};
std::mutex g_head_mutex;
static Tuning current_tuning() {
  std::lock_guard<std::mutex> lock(g_head_mutex);
  return g_tuning;
}

static int choose_config(int CK, int lpr_req, Plan* P) {
  struct Cfg { int cpl, lpr, nt, minb; };
  static const Cfg cfgs[] = {
#define X(cpl, lpr, nt, minb_fwd, minb_bwd) {cpl, lpr, nt, minb_fwd},
      SIMT_HEAD_CONFIGS(X)
#undef X
  };
  const Cfg* best = nullptr;
  for (const Cfg& c : cfgs) {
    if (c.cpl * c.lpr < CK) continue;
    if (lpr_req > 0 && c.lpr != lpr_req) continue;
    // prefer the fewest lanes per cell, then the least channel padding
    if (!best || c.lpr < best->lpr || (c.lpr == best->lpr && c.cpl * c.lpr < best->cpl * best->lpr)) best = &c;
  }
  if (!best && lpr_req > 0) return choose_config(CK, 0, P);
  if (!best) return SIMT_EUNSUPPORTED;
  P->CPL = best->cpl; P->LPR = best->lpr; P->NT = best->nt; P->MINB = best->minb;
  P->CKP = best->cpl * best->lpr;
  return 0;
}

static int make_plan(int mode, int B, int CK, int C, int h, int w, int H, int W, HeadArgs* A, Plan* P) {
  const Tuning tune = current_tuning();
  int rc = choose_config(CK, tune.lpr, P);
  if (rc) return rc;
  A->sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  A->sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  A->ncy = h > 1 ? h - 1 : 1;
  A->ncx = w > 1 ? w - 1 : 1;
  // cell-rows per unit: ~8 pixel rows per unit keeps the per-unit overhead amortised
  int ur = tune.ur;
  if (ur <= 0) {
    const double rows_per_cell = (double)H / (double)A->ncy;
    ur = (int)(8.0 / rows_per_cell + 0.5);
    if (ur < 1) ur = 1;
    if (ur > 32) ur = 32;
  }
  const int cpw = 32 / P->LPR;
  A->ur = ur;
  // split each cell-row over rs units when the grid would otherwise see only a few units per warp
  int rs = tune.unused > 0 ? tune.unused : 0;
  if (rs <= 0) {
    DeviceInfo di;
    if (device_info(&di) == 0) {
      const double warps = (double)di.sm_count * 3.0 * (P->NT / 32);   // ~3 CTAs per SM resident
      const double base_units = (double)B * ((A->ncy + ur - 1) / ur) * ((A->ncx + cpw - 1) / cpw);
      rs = (int)(4.0 * warps / base_units + 0.5);                      // aim at >= ~4 units per warp
      if (rs > 4) rs = 4;
    }
  }
  if (ur > 1) rs = 1;
  A->rs = rs < 1 ? 1 : rs;
  A->units_y = ((A->ncy + ur - 1) / ur) * A->rs;
  A->units_x = (A->ncx + cpw - 1) / cpw;
  A->nunits = (long long)B * A->units_y * A->units_x;
  const bool bwd = mode != MODE_FWD;
  const size_t nw = (size_t)(P->NT / 32);
  P->smem = nw * 4 * (P->CPL / 2) * 32 * 8 + (size_t)C * P->CKP * 4 +
            (bwd ? nw * kEdgeRows * (P->CKP + 1) * 4 : 0) + (size_t)(A->ncx + A->ncy + 2) * 4;
#ifdef SIMT_EXP_LXTAB
  P->smem += (size_t)(W + H) * 4;
#endif
  return 0;
}

static int validate(const float* logits, int B, int CK, int h, int w, int C, const void* labels, int label_bytes,
                    int H, int W, const float* T) {
  if (!logits || !labels) return SIMT_EINVAL;
  if (B <= 0 || CK <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return SIMT_EINVAL;
  if (label_bytes != 1 && label_bytes != 8) return SIMT_EINVAL;
  if (CK > kMaxCKP || C > 254) return SIMT_EUNSUPPORTED;
  if (!T && C != CK) return SIMT_EINVAL;
  return 0;
}

static int run_head(int mode, const float* logits, int B, int CK, int h, int w, const float* T, int C,
                    const void* labels, int label_bytes, int H, int W, int ignore, float gscale, float* dlogits,
                    double* stats, float* loss_mean, float* dT_out, int* err_flag, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  int rc = validate(logits, B, CK, h, w, C, labels, label_bytes, H, W, T);
  if (rc) return rc;
  if (!err_flag || !workspace) return SIMT_EINVAL;
  if (mode != MODE_FWD && !dlogits) return SIMT_EINVAL;
  if (workspace_bytes < simt_head_workspace_bytes(B, CK, C, h, w, H, W)) return SIMT_EWORKSPACE;
  HeadArgs A{};
  Plan P{};
  A.logits = logits; A.T = T; A.labels = labels;
  A.B = B; A.CK = CK; A.C = C; A.h = h; A.w = w; A.H = H; A.W = W; A.ignore = ignore;
  A.gscale = gscale; A.dlogits = dlogits; A.err = err_flag;
  A.label_words_ok = (label_bytes == 1 && (reinterpret_cast<uintptr_t>(labels) & 3) == 0 &&
                      (((long long)B * H * W) & 3) == 0) ? 1 : 0;
  rc = make_plan(mode, B, CK, C, h, w, H, W, &A, &P);
  if (rc) return rc;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc) return rc;
  const size_t G = (size_t)di.sm_count * kMaxGridPerSm;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  A.counter = reinterpret_cast<unsigned long long*>(ws);
  A.part_loss = reinterpret_cast<double*>(ws + kWsHeader + ws_staging_bytes(CK, C));
  A.part_cnt = reinterpret_cast<long long*>(ws + kWsHeader + ws_staging_bytes(CK, C) + G * 8);
  A.part_dT = reinterpret_cast<float*>(ws + kWsHeader + ws_staging_bytes(CK, C) + G * 16);
  A.ntiles = di.sm_count < kFinSlices * kFinMaxPer ? di.sm_count : kFinSlices * kFinMaxPer;
  if (mode != MODE_FWD)
    SIMT_CUDA_TRY(cudaMemsetAsync(dlogits, 0, (size_t)B * CK * h * w * sizeof(float), st));
  int grid = 0;
  rc = dispatch_all(mode, label_bytes, A, P, st, &grid);
  if (rc) return rc;
  const int fgrid = (C * P.CKP + 31) / 32 + 1;
  head_finalize_kernel<<<fgrid, 1024, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, A.ntiles, CK, P.CKP, C, mode, gscale,
                                              A.counter, stats, loss_mean, dT_out, err_flag, nullptr, nullptr, nullptr,
                                              XchgArgs{}, 0);
  return (int)cudaGetLastError();
}

// One whole training step of the head on one GPU (the path HeadRunner.step takes): label count + dLogits zeroing, the
// fused kernel applying the final scale, finalize.  Three launches, no pass over dLogits after the kernel.
static int run_step(const float* logits, int B, int CK, int h, int w, const float* T, int C, const void* labels,
                    int label_bytes, int H, int W, int ignore, const float* grad_out, float* dlogits, float* dT,
                    double* stats, float* loss_mean, int* err_flag, void* workspace, size_t workspace_bytes,
                    const XchgArgs& X, const void* next_labels, int defer, cudaStream_t st) {
  int rc = validate(logits, B, CK, h, w, C, labels, label_bytes, H, W, T);
  if (rc) return rc;
  if (!err_flag || !workspace || !dlogits || !stats) return SIMT_EINVAL;
  if (workspace_bytes < simt_head_workspace_bytes(B, CK, C, h, w, H, W)) return SIMT_EWORKSPACE;
  HeadArgs A{};
  Plan P{};
  A.logits = logits; A.T = T; A.labels = labels;
  A.B = B; A.CK = CK; A.C = C; A.h = h; A.w = w; A.H = H; A.W = W; A.ignore = ignore;
  A.gscale = 1.f; A.dlogits = dlogits; A.err = err_flag; A.grad_out = grad_out;
  A.X = X;
  A.label_words_ok = (label_bytes == 1 && (reinterpret_cast<uintptr_t>(labels) & 3) == 0 &&
                      (((long long)B * H * W) & 3) == 0) ? 1 : 0;
  rc = make_plan(MODE_STEP, B, CK, C, h, w, H, W, &A, &P);
  if (rc) return rc;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc) return rc;
  const size_t G = (size_t)di.sm_count * kMaxGridPerSm;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned long long* wsh = reinterpret_cast<unsigned long long*>(ws);     // the 64-byte header (see WS_*)
  A.counter = wsh + WS_COUNTER;
  double* count_local = reinterpret_cast<double*>(wsh + WS_COUNT_LOCAL);
  A.count_local = count_local;
  A.count_global = reinterpret_cast<double*>(wsh + WS_COUNT_GLOBAL);
  A.ws_hdr = wsh;
  A.fin = FinishArgs{stats, loss_mean, dT, grad_out, err_flag, CK, C, P.CKP};
  A.part_loss = reinterpret_cast<double*>(ws + kWsHeader + ws_staging_bytes(CK, C));
  A.part_cnt = reinterpret_cast<long long*>(ws + kWsHeader + ws_staging_bytes(CK, C) + G * 8);
  A.part_dT = reinterpret_cast<float*>(ws + kWsHeader + ws_staging_bytes(CK, C) + G * 16);
  A.ntiles = di.sm_count < kFinSlices * kFinMaxPer ? di.sm_count : kFinSlices * kFinMaxPer;
  FinishArgs F{stats, loss_mean, dT, grad_out, err_flag, CK, C, P.CKP};
  const long long n_dl = (long long)B * CK * h * w, npix = (long long)B * H * W;
  const int pgrid = di.sm_count * 4;
  // (development aid: SIMT_PROF_WHICH=1 / 2 moves the launch profiler's event pair from the fused kernel to the
  // prologue / finalize kernel)
  static const int prof_which = getenv("SIMT_PROF_WHICH") ? atoi(getenv("SIMT_PROF_WHICH")) : 0;
  if (prof_which == 1) prof_begin(st);
  if (label_bytes == 1)
    head_prep_kernel<uint8_t><<<pgrid, 256, 0, st>>>(dlogits, n_dl, static_cast<const uint8_t*>(labels),
                                                     static_cast<const uint8_t*>(next_labels), npix, C, ignore, wsh, X, F);
  else
    head_prep_kernel<long long><<<pgrid, 256, 0, st>>>(dlogits, n_dl, static_cast<const long long*>(labels),
                                                       static_cast<const long long*>(next_labels), npix, C, ignore, wsh, X, F);
  if (prof_which == 1) prof_end(st);
  SIMT_CUDA_TRY(cudaGetLastError());
  int grid = 0;
  rc = dispatch_all(X.world > 1 ? MODE_STEPX : MODE_STEP, label_bytes, A, P, st, &grid);
  if (rc) return rc;
  const int fgrid = (C * P.CKP + 31) / 32 + 1;
  if (prof_which == 2) prof_begin(st);
  head_finalize_kernel<<<fgrid, 1024, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, A.ntiles, CK, P.CKP, C, MODE_STEP,
                                              1.f, A.counter, stats, loss_mean, dT, err_flag, grad_out, count_local,
                                              wsh, X, (X.world > 1 && defer) ? 1 : 0);
  if (prof_which == 2) prof_end(st);
  return (int)cudaGetLastError();
}

static int run_place(const float* logits, int B, int CK, int h, int w, int C, int H, int W, float thres, float lambda_place,
                     float* dlogits, double* stats, float* loss_mean, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
  if (!logits || !dlogits || !workspace || (!stats && !loss_mean)) return SIMT_EINVAL;
  if (B <= 0 || CK <= 0 || C <= 0 || C > CK || h <= 0 || w <= 0 || H <= 0 || W <= 0) return SIMT_EINVAL;
  if (CK > kMaxCKP || C > 254) return SIMT_EUNSUPPORTED;
  if (workspace_bytes < simt_head_workspace_bytes(B, CK, C, h, w, H, W)) return SIMT_EWORKSPACE;
  HeadArgs A{};
  Plan P{};
  A.logits = logits; A.T = nullptr; A.labels = nullptr;
  A.B = B; A.CK = CK; A.C = C; A.h = h; A.w = w; A.H = H; A.W = W; A.ignore = 255;
  A.gscale = 1.f; A.dlogits = dlogits; A.err = nullptr;
  A.place_thres = thres; A.place_lambda = lambda_place;
  int rc = make_plan(MODE_PLACE, B, CK, C, h, w, H, W, &A, &P);
  if (rc) return rc;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc) return rc;
  const size_t G = (size_t)di.sm_count * kMaxGridPerSm;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  A.counter = reinterpret_cast<unsigned long long*>(ws);
  A.part_loss = reinterpret_cast<double*>(ws + kWsHeader + ws_staging_bytes(CK, C));
  A.part_cnt = reinterpret_cast<long long*>(ws + kWsHeader + ws_staging_bytes(CK, C) + G * 8);
  A.part_dT = reinterpret_cast<float*>(ws + kWsHeader + ws_staging_bytes(CK, C) + G * 16);
  A.ntiles = di.sm_count < kFinSlices * kFinMaxPer ? di.sm_count : kFinSlices * kFinMaxPer;
  SIMT_CUDA_TRY(cudaMemsetAsync(dlogits, 0, (size_t)B * CK * h * w * sizeof(float), st));
  int grid = 0;
  rc = dispatch_all(MODE_PLACE, 1, A, P, st, &grid);
  if (rc) return rc;
  // one block: loss / count partials and the scheduler re-arm (there are no dT tiles in this mode)
  head_finalize_kernel<<<1, 1024, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, 0, CK, P.CKP, C, MODE_PLACE, 1.f,
                                           A.counter, stats, loss_mean, nullptr, nullptr, nullptr, nullptr, nullptr,
                                           XchgArgs{}, 0);
  return (int)cudaGetLastError();
}

}  // namespace simt

using namespace simt;

static int make_xchg(int rank, int world, void* const* mailboxes, int CK, int C, XchgArgs* X) {
  if (!mailboxes || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || CK <= 0 || C <= 0) return SIMT_EINVAL;
  for (int r = 0; r < world; ++r) {
    if (!mailboxes[r]) return SIMT_EINVAL;
    X->mail[r] = static_cast<unsigned char*>(mailboxes[r]);
  }
  X->rank = rank; X->world = world; X->n_stats = 2 + CK * C;
  X->slot_entries = 2 + C * kXchgMaxCKP;
  X->max_spins = xchg_max_spins();
  return 0;
}

extern "C" {

size_t simt_head_workspace_bytes(int B, int CK, int C, int h, int w, int H, int W) {
  (void)B; (void)h; (void)w; (void)H; (void)W;
  DeviceInfo di;
  if (device_info(&di)) di.sm_count = 256;
  const size_t G = (size_t)di.sm_count * kMaxGridPerSm;
  return kWsHeader + ws_staging_bytes(CK > 0 ? CK : 1, C > 0 ? C : 1) + G * 16 +
         G * (size_t)(C > 0 ? C : 1) * kMaxCKP * sizeof(double);
}

void simt_head_set_tuning(int cell_rows_per_unit, int reserved, int threads, int lpr) {
  std::lock_guard<std::mutex> lock(g_head_mutex);
  g_tuning = {cell_rows_per_unit, reserved, threads, lpr};
}

int simt_head_fwd(const float* logits, int B, int CK, int h, int w, const float* T, int C, const void* labels,
                  int label_bytes, int H, int W, int ignore, double* stats, float* loss_mean, int* err_flag,
                  void* workspace, size_t workspace_bytes, void* stream) {
  if (!stats && !loss_mean) return SIMT_EINVAL;
  return run_head(MODE_FWD, logits, B, CK, h, w, T, C, labels, label_bytes, H, W, ignore, 1.f, nullptr, stats,
                  loss_mean, nullptr, err_flag, workspace, workspace_bytes, (cudaStream_t)stream);
}

int simt_head_fwdbwd(const float* logits, int B, int CK, int h, int w, const float* T, int C, const void* labels,
                     int label_bytes, int H, int W, int ignore, float* dlogits_raw, double* stats, float* loss_mean,
                     int* err_flag, void* workspace, size_t workspace_bytes, void* stream) {
  if (!stats) return SIMT_EINVAL;
  return run_head(MODE_FWDBWD, logits, B, CK, h, w, T, C, labels, label_bytes, H, W, ignore, 1.f, dlogits_raw, stats,
                  loss_mean, nullptr, err_flag, workspace, workspace_bytes, (cudaStream_t)stream);
}

int simt_head_bwd(const float* logits, int B, int CK, int h, int w, const float* T, int C, const void* labels,
                  int label_bytes, int H, int W, int ignore, float scale, float* dlogits, float* dT, int* err_flag,
                  void* workspace, size_t workspace_bytes, void* stream) {
  return run_head(MODE_BWD, logits, B, CK, h, w, T, C, labels, label_bytes, H, W, ignore, scale, dlogits, nullptr,
                  nullptr, dT, err_flag, workspace, workspace_bytes, (cudaStream_t)stream);
}

int simt_head_step(const float* logits, int B, int CK, int h, int w, const float* T, int C, const void* labels,
                   int label_bytes, int H, int W, int ignore, const float* grad_out, float* dlogits, float* dT,
                   double* stats, float* loss_mean, int* err_flag, void* workspace, size_t workspace_bytes, void* stream) {
  return run_step(logits, B, CK, h, w, T, C, labels, label_bytes, H, W, ignore, grad_out, dlogits, dT, stats, loss_mean,
                  err_flag, workspace, workspace_bytes, XchgArgs{}, nullptr, 0, (cudaStream_t)stream);
}

int simt_head_step_sharded(const float* logits, int B, int CK, int h, int w, const float* T, int C, const void* labels,
                           int label_bytes, int H, int W, int ignore, const float* grad_out, float* dlogits, float* dT,
                           double* stats, float* loss_mean, int* err_flag, void* workspace, size_t workspace_bytes,
                           int rank, int world, void* const* mailboxes, const void* next_labels, int defer,
                           void* stream) {
  XchgArgs X{};
  int rc = make_xchg(rank, world, mailboxes, CK, C, &X);
  if (rc) return rc;
  return run_step(logits, B, CK, h, w, T, C, labels, label_bytes, H, W, ignore, grad_out, dlogits, dT, stats, loss_mean,
                  err_flag, workspace, workspace_bytes, X, next_labels, defer, (cudaStream_t)stream);
}

int simt_head_finish_sharded(int CK, int C, const float* grad_out, float* dT, double* stats, float* loss_mean,
                             int* err_flag, void* workspace, size_t workspace_bytes, int rank, int world,
                             void* const* mailboxes, void* stream) {
  if (!workspace || workspace_bytes < kWsHeader + ws_staging_bytes(CK, C) || !stats) return SIMT_EINVAL;
  XchgArgs X{};
  int rc = make_xchg(rank, world, mailboxes, CK, C, &X);
  if (rc) return rc;
  if (world <= 1) return 0;
  Plan P{};
  rc = choose_config(CK, current_tuning().lpr, &P);     // the dT tile's row length of this channel count
  if (rc) return rc;
  FinishArgs F{stats, loss_mean, dT, grad_out, err_flag, CK, C, P.CKP};
  head_finish_kernel<<<(X.n_stats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      static_cast<unsigned long long*>(workspace), X, F);
  return (int)cudaGetLastError();
}

int simt_placeholder_fwdbwd(const float* logits, int B, int CK, int h, int w, int C, int H, int W, float thres,
                            float lambda_place, float* dlogits_raw, double* stats, float* loss_mean, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return run_place(logits, B, CK, h, w, C, H, W, thres, lambda_place, dlogits_raw, stats, loss_mean, workspace,
                   workspace_bytes, (cudaStream_t)stream);
}

int simt_head_scale(float* dlogits, long long n_dlogits, const double* stats, int CK, int C, const float* grad_out,
                    float* dT, void* stream) {
  if (!stats || n_dlogits < 0 || (n_dlogits > 0 && !dlogits)) return SIMT_EINVAL;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  long long blocks = (n_dlogits / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (long long)di.sm_count * 8) blocks = (long long)di.sm_count * 8;
  head_scale_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dlogits, n_dlogits, stats, dT ? CK * C : 0,
                                                                    grad_out, dT);
  return (int)cudaGetLastError();
}

}  // extern "C"
