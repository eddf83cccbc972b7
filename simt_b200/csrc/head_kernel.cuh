// Fused SimT head for B200 (sm_100a): bilinear upsample (align_corners=True) ->
// channel softmax -> p.T -> masked NLL, and in the same pass the gradients
// dLogits (at LOW resolution: the transposed bilinear operator is applied
// in-kernel) and dT.
//
// Replaces tools/trainV2_simt.py:371-372,402-409 + the autograd backward (:428) and
// utils/loss.py:14-40 of the reference.  Maths (verified against
// autograd): with p = softmax_k(z), q = sum_k p_k T[k,y]:
//     loss = -(1/N) sum_valid log q
//     dz_k = (1/N) (p_k - p_k T[k,y] / q)          (high-res, then U^T to low-res)
//     dT[k,y] += -(1/N) p_k / q
// Only column y of T is touched per pixel, so there is no GEMM here: the kernel
// is bound by the MUFU (one ex2 per channel per pixel) and FP32 issue, not by
// HBM (3.4 B/pixel algorithmic) -- see DESIGN.md.
//
// Work decomposition (v2: warp-autonomous, no CTA barriers in the main loop)
//   low-res "cell" (cy, cx) = the square between 4 neighbouring low-res nodes; every
//   high-res pixel lies in exactly one cell (torch's i0 = min(floor(src), in-1) is
//   re-expressed as cell = min(floor(src), in-2), lambda = clamp(src - cell, 0, 1),
//   which gives identical values: for the last node lambda becomes exactly 1).
//   unit       = UR cell-rows x CPW = 32/LPR cells, claimed dynamically by ONE WARP.
//   lane group = LPR lanes own one cell; each lane owns CPL channels (CK <= CPL*LPR) and
//                stages the cell's 4 corner logits of its channels once per cell-row in a
//                warp-private shared-memory slice (pre-scaled by log2 e, conflict-free), so per
//                pixel row the vertical lerp is 2 FFMA2 per channel pair and per pixel the
//                horizontal lerp is ONE FFMA2 per channel pair:  t_k = a_k + lambda * d_k.
//   Softmax uses a per-row upper bound M of the logits instead of the per-pixel max (the
//   interpolant is a convex combination of the row's end points); a pixel whose exp-sum
//   underflows (only with > 2^40 dynamic range inside one cell) is redone with the exact max.
//   Backward: per pixel row the horizontal transposed lerp is accumulated in registers
//   (Gs = sum g, G1 = sum lambda g), the left neighbour's G1 arrives by one warp shuffle, and
//   the vertical transposed lerp is accumulated in registers too (Vt, Vb); per cell-row each
//   lane group adds its two node rows to dLogits with red.global.add.f32 (coalesced across the
//   warp); the unit's right edge column goes out the same way.
//   dT: per-thread register accumulators D2[] for the thread's current label column, handed to the CTA's fp32 tile
//   in global memory (L2 resident) with 8-byte red.global.add.v2.f32 whenever the lane's label changes and at the
//   end of the CTA (shared-memory float atomics are CAS loops on sm_100 and collapse under contention).
//   Per-CTA partials (loss and count in fp64, the dT tiles in fp32) are reduced in a fixed order by a small
//   finalize kernel.  Units are claimed dynamically, so the grouping of the partial sums (and the order of the
//   fp32 red.adds into dLogits, as in torch's own CUDA backward of upsample_bilinear2d) is not run-to-run
//   deterministic in the last bit.
//   MODE_PLACE reuses the same machinery for Placeholder_loss (tools/trainV2_simt.py:202-230): no labels, no T,
//   the per-pixel body derives both label maps from the logits (see the body's comment).
#include <cstdlib>
#include <mutex>
#include <type_traits>
#pragma once
#include "step_xchg.cuh"

namespace simt {

struct HeadArgs {
  const float* logits;
  const float* T;  // may be null (identity)
  const void* labels;
  int B, CK, C, h, w, H, W, ignore;
  float sy, sx;    // torch's align_corners scales (float)(in-1)/(out-1)
  int ncy, ncx;    // number of cells = max(in-1, 1)
  int ur;          // cell-rows per unit
  int rs;          // a cell-row's pixel rows are split over rs units (only with ur == 1): finer tail
  int units_y, units_x;
  long long nunits;
  float gscale;
  float* dlogits;
  unsigned long long* counter;  // dynamic unit scheduler (zero on entry; finalize re-zeroes it)
  float* part_dT;       // [ntiles][C*CKP] per-SM dT tiles (zero on entry; finalize re-zeroes them)
  int ntiles;           // = SM count: the CTAs resident on one SM share a tile (fewer tiles for finalize to reduce)
  double* part_loss;    // [grid]
  long long* part_cnt;  // [grid]
  int* err;
  int label_words_ok;  // uint8 labels: buffer 4-byte aligned and a multiple of 4 bytes long
  float place_thres, place_lambda;  // MODE_PLACE: confidence threshold (< 0: none), weight of the open-set term
  // MODE_STEP: upstream gradient (device scalar or null = 1) and the valid-pixel count written by head_prep_kernel
  const float* grad_out;
  const double* count_local;
  double* count_global;   // sharded step: the summed count, kept for the kernels that finish the step
  unsigned long long* ws_hdr;   // workspace header (WS_* words)
  FinishArgs fin;               // where a finished step's loss / dT / stats go (sharded, deferred mode)
  // MODE_STEP, batch sharded over several GPUs: the GLOBAL valid-pixel count is the sum of the ranks' counts, which
  // head_prep_kernel pushes into every peer's mailbox; every CTA acquires them in its prologue
  XchgArgs X;
};

template <typename LabelT>
__device__ __forceinline__ int load_label(const LabelT* p, long long idx);
template <>
__device__ __forceinline__ int load_label<uint8_t>(const uint8_t* p, long long idx) {
  return (int)__ldg(p + idx);
}
template <>
__device__ __forceinline__ int load_label<long long>(const long long* p, long long idx) {
  long long v = __ldg(p + idx);
  // negatives are "ignored" (utils/loss.py:29); anything above int range is out of range
  return v < 0 ? -1 : (v > 0x7fffffffLL ? 0x7fffffff : (int)v);
}

template <int LPR>
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
  if (LPR >= 2) v += __shfl_xor_sync(gmask, v, 1);
  if (LPR >= 4) v += __shfl_xor_sync(gmask, v, 2);
  return v;
}
template <int LPR>
__device__ __forceinline__ float group_max(float v, unsigned gmask) {
  if (LPR >= 2) v = fmaxf(v, __shfl_xor_sync(gmask, v, 1));
  if (LPR >= 4) v = fmaxf(v, __shfl_xor_sync(gmask, v, 2));
  return v;
}

template <int LPR>
__device__ __forceinline__ int group_min_i(int v) {
  if (LPR >= 2) v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
  if (LPR >= 4) v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2/FADD2/FMUL2: two fp32 lanes per issue slot) ----
// Operands are packed/unpacked with mov.b64 {lo, hi} inside the asm block (ptxas folds these into
// register-pair allocation); reinterpret_cast of float2 references forces the values through local memory.
#ifdef SIMT_CPU_EMULATION   /* tests/cpu_simt: the same arithmetic, one IEEE operation per component */
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#else
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
#endif
__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }
// fire-and-forget 8-byte vector reduction (sm_90+): *(float2*)p += v, p 8-byte aligned
__device__ __forceinline__ void red_add_v2(float* p, float2 v) {
#ifdef SIMT_CPU_EMULATION
  atomicAdd(p, v.x); atomicAdd(p + 1, v.y);
#else
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
#endif
}

static constexpr float kPadLogit = -1.0e30f;  // padded channels: exp2 -> 0, no inf/NaN arithmetic

// ---- label fetch -----------------------------------------------------------------------------
// A run's first 8 labels travel as one 64-bit word of raw label bytes (pixels past the run's end
// are filled with 0xFF, which is never a valid class because C <= 254).  uint8 labels are fetched
// as three ALIGNED 32-bit words covering the (unaligned) run and are only funnel-shifted together
// when the row is processed, one row after the loads were issued, so their latency is hidden
// behind the previous row's arithmetic.  int64 labels (the reference's dtype) are converted at
// load time (slower, drop-in path).
struct RawRun {
  unsigned w0, w1, w2, sh;  // uint8: aligned words + byte shift ; int64: w0/w1 hold the packed bytes
};

__device__ __forceinline__ unsigned long long fill_tail(unsigned long long v, int n) {
  const unsigned long long keep = (n >= 8) ? ~0ULL : ((1ULL << (8 * (n < 0 ? 0 : n))) - 1ULL);
  return v | ~keep;
}

template <typename LabelT>
struct LabelFetch;

template <>
struct LabelFetch<uint8_t> {
  // words_ok: the label buffer is 4-byte aligned and its size a multiple of 4 (so aligned word
  // loads never leave the buffer except past `end`, which is guarded)
  static __device__ __forceinline__ RawRun issue(const uint8_t* labels, long long idx, int n, const uint8_t* end,
                                                 bool words_ok, int, int) {
    RawRun r;
    if (words_ok) {
      const uint8_t* a = labels + idx;
      const uintptr_t ai = reinterpret_cast<uintptr_t>(a);
      const unsigned* p = reinterpret_cast<const unsigned*>(ai & ~(uintptr_t)3);
      r.sh = (unsigned)(ai & 3) * 8u;
      const unsigned* e = reinterpret_cast<const unsigned*>(end);
      r.w0 = (n > 0 && p < e) ? __ldg(p) : 0xffffffffu;
      r.w1 = (n > 0 && p + 1 < e) ? __ldg(p + 1) : 0xffffffffu;
      r.w2 = (n > 0 && p + 2 < e) ? __ldg(p + 2) : 0xffffffffu;
    } else {
      unsigned long long v = ~0ULL;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < n) v = (v & ~(0xffULL << (8 * q))) | ((unsigned long long)__ldg(labels + idx + q) << (8 * q));
      r.w0 = (unsigned)v; r.w1 = (unsigned)(v >> 32); r.w2 = 0xffffffffu; r.sh = 0u;
    }
    return r;
  }
  static __device__ __forceinline__ unsigned long long finish(const RawRun& r, int n) {
    const unsigned lo = __funnelshift_r(r.w0, r.w1, r.sh);
    const unsigned hi = __funnelshift_r(r.w1, r.w2, r.sh);
    return fill_tail(((unsigned long long)hi << 32) | lo, n);
  }
  static __device__ __forceinline__ unsigned one(const uint8_t* labels, long long idx, int, int) {
    return (unsigned)__ldg(labels + idx);
  }
};

template <>
struct LabelFetch<long long> {
  // int64 -> byte code: valid class as is, ignore/negative -> ign8 (or 0xFF), anything else -> 0xFE
  static __device__ __forceinline__ unsigned one(const long long* labels, long long idx, int ignore, int C) {
    const long long y = __ldg(labels + idx);
    if (y == (long long)ignore || y < 0) return (ignore >= 0 && ignore <= 255) ? (unsigned)ignore : 0xffu;
    return (y < (long long)C) ? (unsigned)y : 0xfeu;
  }
  static __device__ __forceinline__ RawRun issue(const long long* labels, long long idx, int n, const long long*,
                                                 bool, int ignore, int C) {
    unsigned long long v = ~0ULL;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < n) v = (v & ~(0xffULL << (8 * q))) | ((unsigned long long)one(labels, idx + q, ignore, C) << (8 * q));
    RawRun r;
    r.w0 = (unsigned)v; r.w1 = (unsigned)(v >> 32); r.w2 = 0xffffffffu; r.sh = 0u;
    return r;
  }
  static __device__ __forceinline__ unsigned long long finish(const RawRun& r, int n) {
    return fill_tail(((unsigned long long)r.w1 << 32) | r.w0, n);
  }
};

static constexpr int kEdgeRows = 16;  // pixel rows per cell-row whose edge column is staged in smem

template <int CPL, int LPR, int MODE, typename LabelT, int NT, int MINB, bool IDENT>
__global__ void __launch_bounds__(NT, MINB) head_kernel(const HeadArgs A) {
  static_assert(CPL % 2 == 0, "channels per lane are processed as fp32x2 pairs");
  constexpr bool BWD = (MODE != MODE_FWD);
  constexpr bool PLACE = (MODE == MODE_PLACE);   // Placeholder_loss: labels are derived from the logits, no T
  constexpr int NP = CPL / 2;    // channel pairs per lane
  constexpr int CKP = CPL * LPR;
  constexpr int CPW = 32 / LPR;  // cells per warp unit
  constexpr int NW = NT / 32;
#ifdef SIMT_CPU_EMULATION
  unsigned char* smem_raw = cpusimt::dynamic_smem();
#else
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const int CK = A.CK, C = A.C;
  // [NW][4][NP][32] float2: the cell corners of every lane (log2 domain), re-read once per pixel row
  float2* Lsm = reinterpret_cast<float2*>(smem_raw);
  unsigned char* sp = smem_raw + (size_t)NW * 4 * NP * 32 * sizeof(float2);
  float* Ts = reinterpret_cast<float*>(sp);                           // [C][CKP] = -T^T
  sp += (size_t)C * CKP * 4;
  float* Esm = reinterpret_cast<float*>(sp);                          // [NW][kEdgeRows][CKP + 1] edge column (BWD)
  sp += BWD ? (size_t)NW * kEdgeRows * (CKP + 1) * 4 : 0;
  int* xs_tab = reinterpret_cast<int*>(sp);                           // [ncx + 1] first pixel column of every cell
  int* ys_tab = xs_tab + (A.ncx + 1);                                 // [ncy + 1] first pixel row of every cell-row
#ifdef SIMT_EXP_LXTAB
  float* lx_tab = reinterpret_cast<float*>(ys_tab + (A.ncy + 1));     // [W] horizontal lerp weight of every pixel column
  float* ly_tab = lx_tab + A.W;                                       // [H] vertical lerp weight of every pixel row
#endif
  __shared__ double red_d[NW];
  __shared__ long long red_i[NW];
  __shared__ float s_gs;   // MODE_STEP: grad_out / N_valid (global)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int sub = (LPR > 1) ? (tid & (LPR - 1)) : 0;
  const int kbase = sub * CPL;
  const int pidx = lane / LPR;  // this lane group's cell within the unit
  const LabelT* labels = reinterpret_cast<const LabelT*>(A.labels);
  const LabelT* labels_end = labels + (long long)A.B * A.H * A.W;
  const int h = A.h, w = A.w;
  // the byte code that means "ignored".  An ignore label outside 0..255 matches no uint8 label (256); int64 labels are
  // converted by LabelFetch<long long>::one, which emits 0xFF for ignored / negative labels in that case
  const int ign8 = (A.ignore >= 0 && A.ignore <= 255) ? A.ignore : (sizeof(LabelT) == 8 ? 255 : 256);
  // plain CE (T = NULL -> I): q = p_y can underflow where torch's log_softmax stays finite; its own instantiation, so
  // that the T-corrected kernels carry none of the extra range tracking
  constexpr bool ident = IDENT;
  float2* Lw = Lsm + (size_t)(tid >> 5) * 4 * NP * 32 + lane;  // + (arr * NP + q) * 32
  float* Ew = Esm + (size_t)(tid >> 5) * kEdgeRows * (CKP + 1);

  // ---- one-time per CTA: -T transposed ([y][k], zero padded), pixel/cell tables ----
  for (int i = tid; i < (PLACE ? 0 : C * CKP); i += NT) {
    int y = i / CKP, k = i - y * CKP;
    float v = 0.f;
    if (k < CK) v = A.T ? __ldg(A.T + (size_t)k * C + y) : (k == y ? 1.f : 0.f);
    Ts[i] = -v;
  }
  if (MODE == MODE_STEP && tid == 0)   // the valid-pixel count was produced by head_prep_kernel before this launch
    s_gs = (float)((A.grad_out ? (double)__ldg(A.grad_out) : 1.0) / *A.count_local);
#include "stepx_acquire.inc"   // MODE_STEPX: warp 0 acquires the ranks' valid counts
  for (int i = tid; i <= A.ncx; i += NT) xs_tab[i] = first_px_of_cell(i, A.sx, A.ncx, A.W);
  for (int i = tid; i <= A.ncy; i += NT) ys_tab[i] = first_px_of_cell(i, A.sy, A.ncy, A.H);
#ifdef SIMT_EXP_LXTAB
  for (int i = tid; i < A.W; i += NT) lx_tab[i] = lambda_of(i, A.sx, cell_of(i, A.sx, A.ncx));
  for (int i = tid; i < A.H; i += NT) ly_tab[i] = lambda_of(i, A.sy, cell_of(i, A.sy, A.ncy));
#endif
#include "stepx_cta0.inc"      // MODE_STEPX: CTA 0 settles the deferred exchange of earlier steps
  unsigned smid;
#ifdef SIMT_CPU_EMULATION
  smid = blockIdx.x;
#else
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
  float* ct = A.part_dT + (size_t)(smid % (unsigned)A.ntiles) * C * CKP;  // this SM's dT tile in global memory (L2 resident)
  __syncthreads();

  float2 D2[NP];     // dT accumulators (p_k / q) for the thread's current label column
  float2 nTc[NP];    // -T[:, cur] for this lane's channels
#pragma unroll
  for (int q = 0; q < NP; ++q) { D2[q] = make_float2(0.f, 0.f); nTc[q] = make_float2(0.f, 0.f); }
  int cur = -1;
  double loss_d = 0.0;  // sum of log2 q over this thread's valid pixels
  long long cnt = 0;
  int badf = 0;  // contract violation seen (a label that is neither a class nor the ignore label)

  // lane 0 claims; the value is only broadcast when the claimed unit is started, so the atomic's
  // round trip to L2 overlaps the unit in progress
#ifdef SIMT_EXP_CLAIM
  // inline PTX: the compiler turns a one-lane atomicAdd into its warp-aggregated form, whose broadcast shuffle waits
  // for the atomic's round trip on the spot; this one stays in flight until its value is used
  auto claim_raw = [&]() -> unsigned long long {
    unsigned long long r = 0ULL;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0;\n\t@p atom.global.add.u64 %0, [%2], 1;\n\t}"
                 : "+l"(r) : "r"((unsigned)lane), "l"(A.counter));
    return r;
  };
#else
  auto claim_raw = [&]() -> unsigned long long { return (lane == 0) ? atomicAdd(A.counter, 1ULL) : 0ULL; };
#endif
  auto claim_get = [&](unsigned long long raw) -> long long { return (long long)__shfl_sync(0xffffffffu, raw, 0); };
  auto switch_column = [&](int y) {
    cur = y;
    const float2* src = reinterpret_cast<const float2*>(Ts + y * CKP + kbase);
#pragma unroll
    for (int q = 0; q < NP; ++q) nTc[q] = src[q];
  };
  // Hand the lane's dT accumulators (column `cur`) to the CTA tile: native red.global.add.f32, fire and
  // forget (shared-memory float atomics are CAS loops on sm_100 and need warp-collective workarounds).
  auto flush_lane = [&]() {
    // (padded channels carry exact zeros and the tile has CKP columns, so pairs go out unguarded: one
    // 8-byte red.global.add.v2.f32 per channel pair)
    float* dst = ct + cur * CKP + kbase;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      red_add_v2(dst + 2 * q, D2[q]);
      D2[q] = make_float2(0.f, 0.f);
    }
  };

  auto grad_scale = [&]() -> float {
    return (MODE == MODE_BWD) ? A.gscale : ((MODE == MODE_STEP || MODE == MODE_STEPX) ? s_gs : 1.f);
  };

  long long unit = claim_get(claim_raw());
#ifdef SIMT_EXP_PREFETCH
  long long unit_n = claim_get(claim_raw());   // claims run two units ahead: the next unit's id is known early
#endif
  unsigned long long next_raw = claim_raw();
  while (unit < A.nunits) {
    const int per_img = A.units_y * A.units_x;
#ifdef SIMT_EXP_PREFETCH
    if (unit_n < A.nunits) {
      // the logit rows the next unit starts with (2 node rows x CK channels x (CPW + 1) floats, one or two 128-byte
      // lines each) are pulled into L1 behind this unit's arithmetic
      const int nb = (int)(unit_n / per_img);
      const int nrem = (int)(unit_n - (long long)nb * per_img);
      const int nuyr = nrem / A.units_x, nux = nrem - nuyr * A.units_x;
      const int ncy0 = (nuyr / A.rs) * A.ur;
      const int x0 = min(nux * CPW, w - 1), x1 = min(nux * CPW + CPW, w - 1);
      for (int i = lane; i < 2 * CK; i += 32) {
        const int k = i >> 1, r = i & 1;
        const float* pr = A.logits + ((size_t)nb * CK + k) * (h * w) + (size_t)min(ncy0 + r, h - 1) * w;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + x0));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + x1));
      }
    }
#endif
    const int b = (int)(unit / per_img);
    const int urem = (int)(unit - (long long)b * per_img);
    const int uyr = urem / A.units_x, ux = urem - uyr * A.units_x;
    const int uy = uyr / A.rs, part = uyr - uy * A.rs;   // units_y counts (cell-row group, row part) pairs
    const int cx = ux * CPW + pidx;
    const bool cell_ok = cx < A.ncx;
    const int xa = cell_ok ? xs_tab[cx] : 0;
    const int nrun = cell_ok ? xs_tab[cx + 1] - xa : 0;
    const int ncell_u = min(CPW, A.ncx - ux * CPW);            // cells of this unit (warp-uniform)
    const bool last_cell = pidx == ncell_u - 1;
    const int nmax = __reduce_max_sync(0xffffffffu, nrun);
    const int gx0 = min(cx, w - 1), gx1 = min(cx + 1, w - 1);
    const int edge_gx = min(ux * CPW + ncell_u, w - 1);          // node column right of the unit
    float loss_acc = 0.f;
    const int cy_begin = uy * A.ur, cy_end = min(A.ncy, (uy + 1) * A.ur);
    // this unit's pixel rows: all rows of its cell-rows, or the part-th slice of the single cell-row
    const int Yall0 = ys_tab[cy_begin], Yall1 = ys_tab[cy_end];
    const int Yfirst = Yall0 + (int)(((long long)(Yall1 - Yall0) * part) / A.rs);
    const int Ylast = Yall0 + (int)(((long long)(Yall1 - Yall0) * (part + 1)) / A.rs);  // one past the last row
    RawRun raw_next = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u};
    if (!PLACE && Yfirst < Ylast)
      raw_next = LabelFetch<LabelT>::issue(labels, ((long long)b * A.H + Yfirst) * A.W + xa, nrun, labels_end,
                                           A.label_words_ok != 0, A.ignore, C);

    for (int cy = cy_begin; cy < cy_end; ++cy) {
      const int Y0 = max(ys_tab[cy], Yfirst);
      const int Y1 = min(ys_tab[cy + 1], Ylast);
      if (Y1 <= Y0) continue;  // warp-uniform
      const int gy0 = min(cy, h - 1), gy1 = min(cy + 1, h - 1);
      const bool edge_smem = BWD && (Y1 - Y0 <= kEdgeRows);
      // ---- stage the cell's 4 corners of this lane's channels (log2 domain) in the warp's smem slice ----
      {
        // address arithmetic hoisted: 4 corner pointers once per cell-row, then + j * plane (one IMAD.WIDE each)
        const unsigned plane = (unsigned)(h * w);
        const float* q00 = A.logits + ((size_t)b * CK + kbase) * plane + (gy0 * w + gx0);
        const float* q01 = q00 + (gx1 - gx0);
        const float* q10 = q00 + (gy1 - gy0) * w;
        const float* q11 = q10 + (gx1 - gx0);
        __syncwarp();
        // all 4*CPL corner loads are issued before the first one is used (one exposed L2 latency, not ten)
        float c00[CPL], c01[CPL], c10[CPL], c11[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const float pad = cell_ok ? kPadLogit : 0.f;
          c00[j] = c01[j] = c10[j] = c11[j] = pad;
          if (cell_ok && kbase + j < CK) {
            const size_t off = (size_t)((unsigned)j * plane);
            c00[j] = __ldg(q00 + off); c01[j] = __ldg(q01 + off);
            c10[j] = __ldg(q10 + off); c11[j] = __ldg(q11 + off);
          }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const bool r0 = cell_ok && kbase + 2 * q < CK, r1 = cell_ok && kbase + 2 * q + 1 < CK;  // real channels
          const float s0 = r0 ? kLog2e : 1.f, s1 = r1 ? kLog2e : 1.f;   // padding stays at kPadLogit / 0
          const float a0 = c00[2 * q] * s0, a1 = c00[2 * q + 1] * s1;
          const float b0 = c01[2 * q] * s0, b1 = c01[2 * q + 1] * s1;
          Lw[(0 * NP + q) * 32] = make_float2(a0, a1);
          Lw[(1 * NP + q) * 32] = make_float2(c10[2 * q] * s0 - a0, c10[2 * q + 1] * s1 - a1);
          Lw[(2 * NP + q) * 32] = make_float2(b0, b1);
          Lw[(3 * NP + q) * 32] = make_float2(c11[2 * q] * s0 - b0, c11[2 * q + 1] * s1 - b1);
        }
        __syncwarp();
      }
      float2 Vt[NP], Vb[NP];
      if (BWD) {
#pragma unroll
        for (int q = 0; q < NP; ++q) { Vt[q] = make_float2(0.f, 0.f); Vb[q] = make_float2(0.f, 0.f); }
      }

      for (int Y = Y0; Y < Y1; ++Y) {
#ifdef SIMT_EXP_LXTAB
        const float ly = ly_tab[Y];
#else
        const float ly = lambda_of(Y, A.sy, cy);
#endif
        const long long rowbase = ((long long)b * A.H + Y) * A.W + xa;
        const unsigned long long codes = PLACE ? 0ULL : LabelFetch<LabelT>::finish(raw_next, nrun);
        if (!PLACE && Y + 1 < Ylast)  // next row's labels are in flight during this row's arithmetic
          raw_next = LabelFetch<LabelT>::issue(labels, rowbase + A.W, nrun, labels_end, A.label_words_ok != 0,
                                               A.ignore, C);
        float2 Gs[NP], G1[NP];
        if (BWD) {
#pragma unroll
          for (int q = 0; q < NP; ++q) { Gs[q] = make_float2(0.f, 0.f); G1[q] = make_float2(0.f, 0.f); }
        }
        {
          // vertical lerp once per row; a = v0 - M, d = v1 - v0  (M = upper bound of the row's logits)
          float2 a[NP], d[NP];
          float M = -INFINITY, mlo = -INFINITY;  // max_k max(v0,v1) >= every pixel's max >= max_k min(v0,v1)
          float mall = INFINITY;                  // min over the REAL channels of min(v0,v1) (plain CE only)
          const float2 ly2 = bcast2(ly);
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const float2 v0 = ffma2(ly2, Lw[(1 * NP + q) * 32], Lw[(0 * NP + q) * 32]);
            const float2 v1 = ffma2(ly2, Lw[(3 * NP + q) * 32], Lw[(2 * NP + q) * 32]);
            a[q] = v0;
            d[q] = ffma2(v0, bcast2(-1.f), v1);
            M = fmaxf(M, fmaxf(fmaxf(v0.x, v0.y), fmaxf(v1.x, v1.y)));
            mlo = fmaxf(mlo, fmaxf(fminf(v0.x, v1.x), fminf(v0.y, v1.y)));
            if (!PLACE && ident) {
              if (kbase + 2 * q < CK) mall = fminf(mall, fminf(v0.x, v1.x));
              if (kbase + 2 * q + 1 < CK) mall = fminf(mall, fminf(v0.y, v1.y));
            }
          }
          M = group_max<LPR>(M, 0xffffffffu);
          mlo = group_max<LPR>(mlo, 0xffffffffu);
          // exp-sum >= 2^(mlo - M): no pixel of this row can underflow the softmax denominator
          bool range_safe = (M - mlo) < 38.f;
          // plain CE: the label's own exponential must not underflow either (any real channel can be the label)
          // (the shuffle inside group_max is a full-mask collective: it must be executed by every lane, so it may not
          // sit behind `range_safe &&` -- lanes whose row already failed the first test would skip it)
          if (!PLACE && ident) {
            const float nmall = group_max<LPR>(-mall, 0xffffffffu);
            range_safe = range_safe & ((M + nmall) < 60.f);
          }
          const float2 nM = bcast2(-M);
#pragma unroll
          for (int q = 0; q < NP; ++q) a[q] = fadd2(a[q], nM);

          if constexpr (PLACE) {
            // ---- Placeholder_loss (tools/trainV2_simt.py:202-230) on this row's pixels --------------------
            // Per pixel: a = arg-max channel (first on ties); valid iff a < C and max prob > thres;
            //   known   = -log softmax(z)_a
            //   unknown = CE(z', y) with z' = z except z'_a = 0 (a CONSTANT: `ones` at :208 is zeros_like), and
            //             y = the first best open-set channel if its logit is > 0, else class 0 (:220-222)
            // Both softmaxes are taken relative to their own exact maximum (no range assumptions).
            const float tz = -M;  // the logit 0 in this row's shifted log2 domain
            auto pixel = [&](const float lam, const bool wv) {
              float t[CPL], e[CPL], f[CPL];
              const float2 L = bcast2(lam);
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                const float2 tt = ffma2(L, d[q], a[q]);
                t[2 * q] = tt.x; t[2 * q + 1] = tt.y;
              }
              float m0 = -INFINITY, mo = -INFINITY;
#pragma unroll
              for (int j = 0; j < CPL; ++j) {
                m0 = fmaxf(m0, t[j]);
                mo = fmaxf(mo, (kbase + j >= C) ? t[j] : -INFINITY);
              }
              m0 = group_max<LPR>(m0, 0xffffffffu);
              mo = group_max<LPR>(mo, 0xffffffffu);
              int ia = 1 << 20, io = 1 << 20;
#pragma unroll
              for (int j = CPL - 1; j >= 0; --j) {
                if (t[j] == m0) ia = kbase + j;
                if (kbase + j >= C && t[j] == mo) io = kbase + j;
              }
              ia = group_min_i<LPR>(ia);
              io = group_min_i<LPR>(io);
              float m2 = -INFINITY;  // best channel other than the arg-max: the maximum of z' is max(m2, 0)
#pragma unroll
              for (int j = 0; j < CPL; ++j) m2 = fmaxf(m2, (kbase + j == ia) ? -INFINITY : t[j]);
              m2 = group_max<LPR>(m2, 0xffffffffu);
              const float ms = fmaxf(m2, tz);
              float su = 0.f, sp = 0.f;
#pragma unroll
              for (int j = 0; j < CPL; ++j) {
                e[j] = ex2_approx(t[j] - m0);
                f[j] = (kbase + j == ia) ? 0.f : ex2_approx(t[j] - ms);
                su += e[j];
                sp += f[j];
              }
              su = group_sum<LPR>(su, 0xffffffffu);                          // >= 1; max prob = 1 / su
              sp = group_sum<LPR>(sp, 0xffffffffu) + ex2_approx(tz - ms);    // >= 1
              const bool valid = wv && ia < C && (1.f > A.place_thres * su);
              const bool open_pos = mo > tz;                                  // an open-set logit > 0
              const int y = open_pos ? io : 0;
              const float t_first = __shfl_sync(0xffffffffu, t[0], lane & ~(LPR - 1));  // channel 0 of this pixel
              const float ty = (y == ia) ? tz : (open_pos ? mo : t_first);   // z'_y in the shifted domain
              if (valid) {
                loss_acc -= lg2_approx(su) + A.place_lambda * (lg2_approx(sp) + (ms - ty));
                cnt += 1;
              }
              const float r = valid ? rcp_approx(su * sp) : 0.f;
              const float rs = r * sp, rp = A.place_lambda * (r * su);
              const float oa = valid ? 1.f : 0.f, oy = (valid && y != ia) ? A.place_lambda : 0.f;
#pragma unroll
              for (int j = 0; j < CPL; ++j) {
                float g = fmaf(f[j], rp, e[j] * rs);
                g -= (kbase + j == ia) ? oa : 0.f;
                g -= (kbase + j == y) ? oy : 0.f;
                if (j & 1) { Gs[j >> 1].y += g; G1[j >> 1].y = fmaf(lam, g, G1[j >> 1].y); }
                else       { Gs[j >> 1].x += g; G1[j >> 1].x = fmaf(lam, g, G1[j >> 1].x); }
              }
            };
            for (int p = 0; p < nmax; p += 2) {  // warp-uniform: lanes past their run execute predicated-off pixels
              pixel(lambda_of(xa + p, A.sx, cx), p < nrun);
              pixel(lambda_of(xa + p + 1, A.sx, cx), p + 1 < nrun);
            }
          } else {
          // One step = two pixels of every lane's run (two independent MUFU/FMA chains per lane,
          // channel pairs packed into fp32x2 instructions).  WARP-UNIFORM: lane groups exchange partial
          // sums with full-mask shuffles, so invalid lanes/pixels are predicated off (w0/w1), never
          // branched around.
          // pixel 0 belongs to the lane's current label column `cur`; pixel 1 too unless `c1x >= 0`, in which
          // case the pair straddles a label boundary and pixel 1 uses column c1x (its T column is read from
          // shared memory where needed -- predicated loads, no extra live registers, no split step)
          auto body = [&](auto check_underflow, auto may_straddle, const float lam0, const float lam1, const bool w0,
                          const bool w1, const int c1x) {
            // the straddle operands (predicated T-column loads + register copies) exist only in the instantiation
            // that is entered when some lane of the warp has a label boundary inside its pair
            const bool strad = decltype(may_straddle)::value && c1x >= 0;
            const float2* T1 = reinterpret_cast<const float2*>(Ts + (strad ? c1x : 0) * CKP + kbase);
            const float2 L0 = bcast2(lam0), L1 = bcast2(lam1);
            float2 e0[NP], e1[NP];
            float2 sum0 = make_float2(0.f, 0.f), sum1 = sum0, ns0 = sum0, ns1 = sum0;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const float2 t0 = ffma2(L0, d[q], a[q]);
              const float2 t1 = ffma2(L1, d[q], a[q]);
              e0[q] = make_float2(ex2_approx(t0.x), ex2_approx(t0.y));
              e1[q] = make_float2(ex2_approx(t1.x), ex2_approx(t1.y));
              sum0 = fadd2(sum0, e0[q]);
              sum1 = fadd2(sum1, e1[q]);
              ns0 = ffma2(e0[q], nTc[q], ns0);
              ns1 = ffma2(e1[q], strad ? T1[q] : nTc[q], ns1);
            }
            float su0 = group_sum<LPR>(sum0.x + sum0.y, 0xffffffffu);
            float su1 = group_sum<LPR>(sum1.x + sum1.y, 0xffffffffu);
            float s0 = -(ns0.x + ns0.y), s1 = -(ns1.x + ns1.y);
            float lq0 = 0.f, lq1 = 0.f;   // plain CE, rare path: log2 q taken directly (see below)
            bool direct = false;
            if (decltype(check_underflow)::value) {
              s0 = group_sum<LPR>(s0, 0xffffffffu);
              s1 = group_sum<LPR>(s1, 0xffffffffu);
            }
            if (decltype(check_underflow)::value &&
                __any_sync(0xffffffffu, (w0 && (su0 < 1e-12f || (ident && s0 < 1e-25f))) ||
                                            (w1 && (su1 < 1e-12f || (ident && s1 < 1e-25f))))) {
              // the row-level bound M was far above some pixel's true max: redo with the exact max
              float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                const float2 t0 = ffma2(L0, d[q], a[q]), t1 = ffma2(L1, d[q], a[q]);
                tm0 = fmaxf(tm0, fmaxf(t0.x, t0.y));
                tm1 = fmaxf(tm1, fmaxf(t1.x, t1.y));
              }
              tm0 = group_max<LPR>(tm0, 0xffffffffu);
              tm1 = group_max<LPR>(tm1, 0xffffffffu);
              su0 = su1 = s0 = s1 = 0.f;
              float ty0 = 0.f, ty1 = 0.f;
#pragma unroll
              for (int q = 0; q < NP; ++q) {  // (static indices only: a dynamic q would push the arrays to local memory)
                const float2 t0 = ffma2(L0, d[q], a[q]), t1 = ffma2(L1, d[q], a[q]);
                e0[q] = make_float2(ex2_approx(t0.x - tm0), ex2_approx(t0.y - tm0));
                e1[q] = make_float2(ex2_approx(t1.x - tm1), ex2_approx(t1.y - tm1));
                const float2 tq1 = strad ? T1[q] : nTc[q];
                if (ident) {
                  // T = I: the column is minus a one-hot vector.  q = e_y / sum can underflow although log q and the
                  // gradient p_k - [k == y] are finite (torch's log_softmax): keep the label's logit for the loss and
                  // clamp its exponential so that e_y / s = 1 exactly (p_y itself is below fp32 resolution there)
                  if (nTc[q].x != 0.f) { ty0 = t0.x - tm0; e0[q].x = fmaxf(e0[q].x, 1e-30f); }
                  if (nTc[q].y != 0.f) { ty0 = t0.y - tm0; e0[q].y = fmaxf(e0[q].y, 1e-30f); }
                  if (tq1.x != 0.f) { ty1 = t1.x - tm1; e1[q].x = fmaxf(e1[q].x, 1e-30f); }
                  if (tq1.y != 0.f) { ty1 = t1.y - tm1; e1[q].y = fmaxf(e1[q].y, 1e-30f); }
                }
                su0 += e0[q].x + e0[q].y;
                su1 += e1[q].x + e1[q].y;
                s0 -= e0[q].x * nTc[q].x + e0[q].y * nTc[q].y;
                s1 -= e1[q].x * tq1.x + e1[q].y * tq1.y;
              }
              su0 = group_sum<LPR>(su0, 0xffffffffu);
              su1 = group_sum<LPR>(su1, 0xffffffffu);
              s0 = group_sum<LPR>(s0, 0xffffffffu);
              s1 = group_sum<LPR>(s1, 0xffffffffu);
              if (ident) {
                direct = true;
                lq0 = group_sum<LPR>(ty0, 0xffffffffu) - lg2_approx(su0);
                lq1 = group_sum<LPR>(ty1, 0xffffffffu) - lg2_approx(su1);
              }
            }
            if (!decltype(check_underflow)::value) {
              s0 = group_sum<LPR>(s0, 0xffffffffu);
              s1 = group_sum<LPR>(s1, 0xffffffffu);
            }
            // 1/sum and 1/s from ONE reciprocal of the product (MUFU is the binding pipe)
            const float r0 = w0 ? rcp_approx(su0 * s0) : 0.f;
            const float r1 = w1 ? rcp_approx(su1 * s1) : 0.f;
            const float rs0 = r0 * s0, rs1 = r1 * s1;    // 1 / sum
            if (MODE != MODE_BWD) {
              if (decltype(check_underflow)::value && direct) {
                loss_acc += (w0 ? lq0 : 0.f) + (w1 ? lq1 : 0.f);
              } else {
                // log2 q0 + log2 q1 = log2(q0 q1); q in (0, 1] and an invalid pixel contributes q = 1
                const float q0 = w0 ? s0 * rs0 : 1.f, q1 = w1 ? s1 * rs1 : 1.f;
                const float qq = q0 * q1;
                if (qq > 1e-30f) loss_acc += lg2_approx(qq);
                else loss_acc += lg2_approx(q0) + lg2_approx(q1);
              }
            }
            cnt += (int)w0 + (int)w1;
            if (BWD) {
              const float is0 = r0 * su0, is1 = r1 * su1;  // 1 / s
              const float2 I0 = bcast2(is0), I1 = bcast2(is1);
              const float2 R0 = bcast2(rs0), R1 = bcast2(rs1);
              const float2 LI0 = bcast2(lam0 * is0), LI1 = bcast2(lam1 * is1);
              const float2 LR0 = bcast2(lam0 * rs0), LR1 = bcast2(lam1 * rs1);
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                // c = (p_k - p_k T_ky / q) / e_k = rs - T_ky * is ;  c1 = lambda * c
                const float2 tq1 = strad ? T1[q] : nTc[q];
                const float2 ca = ffma2(nTc[q], I0, R0), cb = ffma2(tq1, I1, R1);
                const float2 c1a = ffma2(nTc[q], LI0, LR0), c1b = ffma2(tq1, LI1, LR1);
                Gs[q] = ffma2(e0[q], ca, Gs[q]);
                Gs[q] = ffma2(e1[q], cb, Gs[q]);
                G1[q] = ffma2(e0[q], c1a, G1[q]);
                G1[q] = ffma2(e1[q], c1b, G1[q]);
                D2[q] = ffma2(e0[q], I0, D2[q]);   // p_k / q, column `cur`
              }
              if (strad) {  // pixel 1 opens a new label column
                if (cur >= 0) flush_lane();
                switch_column(c1x);
              }
#pragma unroll
              for (int q = 0; q < NP; ++q) D2[q] = ffma2(e1[q], I1, D2[q]);
            }
          };

          // ---- the pixel loop: two pixels of every lane's run per step; the column switch is per lane and
          // shuffle-free.  The loop is WARP-UNIFORM (vote on "anyone left"), lanes that are done run
          // predicated-off steps.
          // Rows alternate direction (boustrophedon): a run that contains a label boundary A|B is walked
          // A..B on one row and B..A on the next, so the lane switches its label column (and flushes its dT
          // accumulators) once per row instead of twice.
          const bool reverse = ((Y - Yall0) & 1) != 0;
          auto run_row = [&](auto check_underflow) {
            int done = 0;  // pixels of this lane's run already consumed
            while (__any_sync(0xffffffffu, done < nrun)) {
              // the next two pixels in walking order: p0, then p1 = p0 +- 1
              const int p0 = reverse ? nrun - 1 - done : done;
              const int p1 = reverse ? p0 - 1 : p0 + 1;
              const bool in0 = done < nrun, in1 = done + 1 < nrun;
              unsigned c0, c1;
              if (nrun <= 8) {
                c0 = in0 ? (unsigned)(codes >> (8 * p0)) & 0xffu : 0xffu;
                c1 = in1 ? (unsigned)(codes >> (8 * p1)) & 0xffu : 0xffu;
              } else {  // runs longer than the 8 prefetched labels (large up-sampling factors)
                c0 = in0 ? LabelFetch<LabelT>::one(labels, rowbase + p0, A.ignore, C) : 0xffu;
                c1 = in1 ? LabelFetch<LabelT>::one(labels, rowbase + p1, A.ignore, C) : 0xffu;
              }
              // branch-free validity / contract / pair-split logic (bitwise on purpose: no short-circuit branches)
              const bool k0 = c0 < (unsigned)C, k1 = c1 < (unsigned)C;        // is a class id
              const bool g0 = (int)c0 != ign8, g1 = (int)c1 != ign8;          // is not the ignore label
              badf |= (int)((!k0 & g0 & in0) | (!k1 & g1 & in1));             // neither class nor ignore (nor padding)
              const bool v0 = k0 & g0, v1 = k1 & g1;
              const int lab = v0 ? (int)c0 : (int)c1;
              if ((v0 | v1) & (lab != cur)) {   // a column switch before the pair (rare on coherent maps)
                if (BWD && cur >= 0) flush_lane();
                switch_column(lab);
              }
              const int c1x = (v0 & v1 & (c0 != c1)) ? (int)c1 : -1;   // label boundary inside the pair
#ifdef SIMT_EXP_LXTAB
              const float lam0 = lx_tab[min(max(xa + p0, 0), A.W - 1)], lam1 = lx_tab[min(max(xa + p1, 0), A.W - 1)];
#else
              const float lam0 = lambda_of(xa + p0, A.sx, cx), lam1 = lambda_of(xa + p1, A.sx, cx);
#endif
              if (__any_sync(0xffffffffu, c1x >= 0)) body(check_underflow, std::true_type{}, lam0, lam1, v0, v1, c1x);
              else body(check_underflow, std::false_type{}, lam0, lam1, v0, v1, c1x);
              done += 2;
            }
          };
          if (__all_sync(0xffffffffu, range_safe)) run_row(std::false_type{});
          else run_row(std::true_type{});
          }  // !PLACE
        }
        if (BWD) {
          // node column cx of this row = G0(cx) + G1(cx-1); the left neighbour is LPR lanes below
          const float2 wy1 = bcast2(ly), wy0 = bcast2(1.f - ly);
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            float px = __shfl_up_sync(0xffffffffu, G1[q].x, LPR);
            float py = __shfl_up_sync(0xffffffffu, G1[q].y, LPR);
            if (pidx == 0) { px = 0.f; py = 0.f; }
            const float2 n = make_float2((Gs[q].x - G1[q].x) + px, (Gs[q].y - G1[q].y) + py);
            Vt[q] = ffma2(wy0, n, Vt[q]);
            Vb[q] = ffma2(wy1, n, Vb[q]);
          }
          // right edge of the unit: node column edge_gx belongs to the next unit (or is the image's
          // last column).  Its per-row values wait in the warp's smem slice until the cell-row is done.
          const float gs = edge_smem ? 0.f : grad_scale();   // (warp-uniform call site)
          if (edge_smem) {
            if (last_cell) {
              float* er = Ew + (Y - Y0) * (CKP + 1) + kbase;
#pragma unroll
              for (int q = 0; q < NP; ++q) { er[2 * q] = G1[q].x; er[2 * q + 1] = G1[q].y; }
              if (sub == 0) Ew[(Y - Y0) * (CKP + 1) + CKP] = ly;
            }
          } else if (last_cell) {
            float* dst = A.dlogits + ((size_t)b * CK + kbase) * h * w;
            const float w0y = (1.f - ly) * gs, w1y = ly * gs;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              if (kbase + j < CK) {
                float* pk = dst + (size_t)j * h * w;
                const float g1 = (j & 1) ? G1[j >> 1].y : G1[j >> 1].x;
                atomicAdd(pk + gy0 * w + edge_gx, w0y * g1);
                atomicAdd(pk + gy1 * w + edge_gx, w1y * g1);
              }
            }
          }
        }
      }  // rows of the cell-row

      if (BWD) {
        float* dst = A.dlogits + (size_t)b * CK * h * w;
        const float gs = grad_scale();
        if (cell_ok) {
          const unsigned plane = (unsigned)(h * w);
          float* d0 = dst + (size_t)kbase * plane + (gy0 * w + gx0);
          float* d1 = d0 + (gy1 - gy0) * w;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            if (kbase + j < CK) {
              const size_t off = (size_t)((unsigned)j * plane);
              const float vt = (j & 1) ? Vt[j >> 1].y : Vt[j >> 1].x;
              const float vb = (j & 1) ? Vb[j >> 1].y : Vb[j >> 1].x;
              atomicAdd(d0 + off, vt * gs);
              atomicAdd(d1 + off, vb * gs);
            }
          }
        }
        if (edge_smem) {
          // vertical transposed lerp of the staged edge column: one lane per channel
          __syncwarp();
          for (int k = lane; k < CK; k += 32) {
            float et = 0.f, eb = 0.f;
            for (int r = 0; r < Y1 - Y0; ++r) {
              const float g1 = Ew[r * (CKP + 1) + k], lyr = Ew[r * (CKP + 1) + CKP];
              et = fmaf(1.f - lyr, g1, et);
              eb = fmaf(lyr, g1, eb);
            }
            float* pk = dst + (size_t)k * h * w;
            atomicAdd(pk + gy0 * w + edge_gx, et * gs);
            atomicAdd(pk + gy1 * w + edge_gx, eb * gs);
          }
          __syncwarp();
        }
      }
    }  // cell-rows of the unit

    loss_d += (double)loss_acc;
#ifdef SIMT_EXP_PREFETCH
    unit = unit_n;
    unit_n = claim_get(next_raw);
#else
    unit = claim_get(next_raw);
#endif
    next_raw = claim_raw();
  }

  // ---- CTA epilogue: partials ---------------------------------------------------------------
  if (BWD && !PLACE && cur >= 0) flush_lane();
  if (badf) atomicOr(A.err, SIMT_ERRBIT_LABEL_RANGE);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss_d += __shfl_xor_sync(0xffffffffu, loss_d, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  // every lane of a lane group accumulated the same loss / count: undo the LPR-fold replication
  if (lane == 0) { red_d[tid >> 5] = loss_d / (double)LPR; red_i[tid >> 5] = cnt / LPR; }
  __syncthreads();
  if (tid == 0) {
    double tl = 0; long long tc = 0;
    for (int wv = 0; wv < NW; ++wv) { tl += red_d[wv]; tc += red_i[wv]; }
    A.part_loss[blockIdx.x] = tl;
    A.part_cnt[blockIdx.x] = tc;
  }
}


// ------------------------------------------------------------------------------------------
// launch plumbing shared by the translation units that instantiate the kernel
// ------------------------------------------------------------------------------------------
static constexpr int kMaxGridPerSm = 8;   // G = SM count * 8 bounds the grid (loss / count partials are per CTA)
static constexpr int kMaxCKP = 64;
struct Plan {
  int CPL, LPR, NT, MINB, CKP;
  size_t smem;
};

#ifndef SIMT_CPU_EMULATION
extern std::mutex g_head_mutex;   // guards the tuning hook and the per-instantiation launch caches (head.cu)

template <int CPL, int LPR, int MODE, typename LabelT, int NT, int MINB, bool IDENT>
static int launch_cfg(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
  auto kern = head_kernel<CPL, LPR, MODE, LabelT, NT, MINB, IDENT>;
  // per-instantiation, per-device cache of the attribute / occupancy queries (function attributes are per context)
  struct Cache { size_t smem = 0; int occ = -1; };
  static Cache cache[64];
  int dev = 0;
  SIMT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SIMT_EUNSUPPORTED;
  int occ;
  {
    std::lock_guard<std::mutex> lock(g_head_mutex);
    Cache& c = cache[dev];
    if (c.occ < 0 || P.smem != c.smem) {
      SIMT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
      SIMT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.occ, kern, NT, P.smem));
      c.smem = P.smem;
    }
    occ = c.occ;
  }
  if (occ < 1) return SIMT_ENOSMEM;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  long long g = (long long)occ * di.sm_count;
  if (g > (long long)di.sm_count * kMaxGridPerSm) g = (long long)di.sm_count * kMaxGridPerSm;
  const long long need = (A.nunits + NT / 32 - 1) / (NT / 32);
  if (g > need) g = need;
  if (g < 1) g = 1;
  *grid_out = (int)g;
  static const bool prof_here = !getenv("SIMT_PROF_WHICH") || atoi(getenv("SIMT_PROF_WHICH")) == 0;
  if (prof_here) prof_begin(st);
  kern<<<(int)g, NT, P.smem, st>>>(A);
  if (prof_here) prof_end(st);
  return (int)cudaGetLastError();
}

#endif  // !SIMT_CPU_EMULATION

// channel-count -> (CPL, LPR, threads, min CTAs/SM) instantiations
#ifndef SIMT_MINB_BWD
#define SIMT_MINB_BWD 3
#endif
#ifndef SIMT_MINB_FWD
#define SIMT_MINB_FWD 4
#endif
#ifdef SIMT_HEAD_BENCH_ONLY   /* development builds: only the bench instantiations (seconds instead of minutes) */
#define SIMT_HEAD_CONFIGS(X) X(10, 2, 128, SIMT_MINB_FWD, SIMT_MINB_BWD)
#else
#define SIMT_HEAD_CONFIGS(X) \
  X(10, 2, 128, SIMT_MINB_FWD, SIMT_MINB_BWD)        \
  X(12, 2, 128, 3, 2)                                \
  X(6, 4, 128, 4, 4)         \
  X(10, 4, 128, 4, 3)        \
  X(16, 4, 128, 3, 2)
#endif

#ifndef SIMT_CPU_EMULATION
template <int MODE, typename LabelT, bool IDENT>
static int dispatch(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
#define X(cpl, lpr, nt, minb_fwd, minb_bwd) \
  if (P.CPL == cpl && P.LPR == lpr)          \
    return launch_cfg<cpl, lpr, MODE, LabelT, nt, (MODE == MODE_FWD ? minb_fwd : minb_bwd), IDENT>(A, P, st, grid_out);
  SIMT_HEAD_CONFIGS(X)
#undef X
  return SIMT_EUNSUPPORTED;
}

template <bool IDENT>
static int dispatch_modes(int mode, int label_bytes, const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
#ifdef SIMT_HEAD_BENCH_ONLY
  if (label_bytes != 1) return SIMT_EUNSUPPORTED;
  if (mode == MODE_STEP) return dispatch<MODE_STEP, uint8_t, IDENT>(A, P, st, grid_out);
  if (mode == MODE_STEPX) return dispatch<MODE_STEPX, uint8_t, IDENT>(A, P, st, grid_out);
  if (mode == MODE_FWD) return dispatch<MODE_FWD, uint8_t, IDENT>(A, P, st, grid_out);
  if (mode == MODE_FWDBWD) return dispatch<MODE_FWDBWD, uint8_t, IDENT>(A, P, st, grid_out);
  return SIMT_EUNSUPPORTED;
#else
  if (mode == MODE_PLACE) return dispatch<MODE_PLACE, uint8_t, false>(A, P, st, grid_out);
  if (mode == MODE_STEP)
    return label_bytes == 1 ? dispatch<MODE_STEP, uint8_t, IDENT>(A, P, st, grid_out) : dispatch<MODE_STEP, long long, IDENT>(A, P, st, grid_out);
  if (mode == MODE_STEPX)
    return label_bytes == 1 ? dispatch<MODE_STEPX, uint8_t, IDENT>(A, P, st, grid_out) : dispatch<MODE_STEPX, long long, IDENT>(A, P, st, grid_out);
  if (label_bytes == 1) {
    if (mode == MODE_FWD) return dispatch<MODE_FWD, uint8_t, IDENT>(A, P, st, grid_out);
    if (mode == MODE_FWDBWD) return dispatch<MODE_FWDBWD, uint8_t, IDENT>(A, P, st, grid_out);
    return dispatch<MODE_BWD, uint8_t, IDENT>(A, P, st, grid_out);
  }
  if (mode == MODE_FWD) return dispatch<MODE_FWD, long long, IDENT>(A, P, st, grid_out);
  if (mode == MODE_FWDBWD) return dispatch<MODE_FWDBWD, long long, IDENT>(A, P, st, grid_out);
  return dispatch<MODE_BWD, long long, IDENT>(A, P, st, grid_out);
#endif
}
#endif  // !SIMT_CPU_EMULATION

}  // namespace simt
#include "head_plan.cuh"
