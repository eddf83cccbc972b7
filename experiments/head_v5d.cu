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
#include <mutex>
#include <type_traits>
#include "common.cuh"

namespace simt {

// MODE_STEP: forward + backward with the 1/N_valid scale known ON THE DEVICE before the kernel starts (a label-only
// count pass), so dLogits leave the kernel final and there is no scale pass.  Single GPU only: sharded, the count
// would be a second rendezvous per step on top of the stats exchange (measured slower than scaling after ONE exchange).
enum { MODE_FWD = 0, MODE_FWDBWD = 1, MODE_BWD = 2, MODE_PLACE = 3, MODE_STEP = 4 };

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr double kLn2 = 0.6931471805599453094;

// exact unsigned division by a launch constant: q = (t + ((n - t) >> s1)) >> s2 with t = umulhi(mul, n)
struct FastDiv { unsigned mul, s1, s2; };
static FastDiv make_fastdiv(unsigned d) {
  FastDiv r;
  unsigned l = 0;
  while ((1ULL << l) < (unsigned long long)d) ++l;          // ceil(log2 d)
  r.mul = (unsigned)((((1ULL << l) - d) << 32) / d + 1ULL);
  r.s1 = l < 1 ? l : 1;
  r.s2 = l > 0 ? l - 1 : 0;
  return r;
}

struct HeadArgs {
  const float* logits;
  const float* T;  // may be null (identity)
  const void* labels;
  int B, CK, C, h, w, H, W, ignore;
  float sy, sx;    // torch's align_corners scales (float)(in-1)/(out-1)
  int ncy, ncx;    // number of cells = max(in-1, 1)
  int ur;          // cell-rows per group
  // unit list: groups (image, cell-row group, 16-cell column block) in order; the first `nbig` groups are one unit
  // each, the groups after them are split into 2^rs_log2 row slices (finer granularity for the tail of the schedule)
  unsigned units_x, groups_per_img, nbig, rs_log2, nunits;
  FastDiv div_img, div_ux;   // exact division by groups_per_img / units_x
  int prefetch;    // prefetch the next unit's logit rows into L1 while the current unit computes
  float gscale;
  float* dlogits;
  unsigned long long* counter;  // dynamic unit scheduler (zero on entry; finalize re-zeroes it)
  float* part_dT;       // [ntiles][C*CKP] per-SM dT tiles (zero on entry; finalize re-zeroes them)
  int ntiles;           // = SM count: the CTAs resident on one SM share a tile (fewer tiles for finalize to reduce)
  double* part_loss;    // [grid]
  long long* part_cnt;  // [grid]
  int* err;
  int label_words_ok;  // uint8 labels: buffer 4-byte aligned and a multiple of 4 bytes long
  int boustrophedon;   // odd pixel rows walk their runs right-to-left (fewer label-column switches at A|B boundaries)
  float place_thres, place_lambda;  // MODE_PLACE: confidence threshold (< 0: none), weight of the open-set term
  // MODE_STEP: upstream gradient (device scalar or null = 1) and the valid-pixel count written by head_prep_kernel
  const float* grad_out;
  const double* count_local;
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
__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }
// fire-and-forget 8-byte vector reduction (sm_90+): *(float2*)p += v, p 8-byte aligned
__device__ __forceinline__ void red_add_v2(float* p, float2 v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

static constexpr float kPadLogit = -1.0e30f;  // padded channels: exp2 -> 0, no inf/NaN arithmetic

// ---- label fetch -----------------------------------------------------------------------------
// A run's first 8 labels travel as one 64-bit word of raw label bytes; pixels past the run's end are filled with the
// PAD byte = the ignore label when it fits a byte, else 0xFF -- a value that is "ignored, silently" (never a class
// because C <= 254, never flagged).  uint8 labels are fetched as three ALIGNED 32-bit words covering the (unaligned)
// run and are only funnel-shifted together when the row is processed, one row after the loads were issued, so their
// latency is hidden behind the previous row's arithmetic.  int64 labels (the reference's dtype) are converted at load
// time (slower, drop-in path).
struct RawRun {
  unsigned w0, w1, w2, sh;  // uint8: aligned words + bit shift ; int64: w0/w1 hold the packed bytes
};

// 32-bit streaming load that does not allocate in L1 (labels are touched once; L1 is kept for the logits)
__device__ __forceinline__ unsigned ldg_stream_u32(const unsigned* p) {
  unsigned r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

template <typename LabelT>
struct LabelFetch;

template <>
struct LabelFetch<uint8_t> {
  // window_ok: the three aligned words of every row of the unit lie inside the label buffer
  static __device__ __forceinline__ RawRun issue(const uint8_t* a, int n, bool window_ok, int, int) {
    RawRun r;
    if (window_ok) {
      const uintptr_t ai = reinterpret_cast<uintptr_t>(a);
      const unsigned* p = reinterpret_cast<const unsigned*>(ai & ~(uintptr_t)3);
      r.sh = ((unsigned)ai & 3u) * 8u;
      r.w0 = ldg_stream_u32(p);
      r.w1 = ldg_stream_u32(p + 1);
      r.w2 = ldg_stream_u32(p + 2);
    } else {
      unsigned long long v = ~0ULL;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < n) v = (v & ~(0xffULL << (8 * q))) | ((unsigned long long)__ldg(a + q) << (8 * q));
      r.w0 = (unsigned)v; r.w1 = (unsigned)(v >> 32); r.w2 = 0xffffffffu; r.sh = 0u;
    }
    return r;
  }
  static __device__ __forceinline__ void raw(const RawRun& r, unsigned& lo, unsigned& hi) {
    lo = __funnelshift_r(r.w0, r.w1, r.sh);
    hi = __funnelshift_r(r.w1, r.w2, r.sh);
  }
  static __device__ __forceinline__ unsigned one(const uint8_t* p, int, int) { return (unsigned)__ldg(p); }
};

template <>
struct LabelFetch<long long> {
  // int64 -> byte code: valid class as is, ignore/negative -> the pad byte, anything else -> 0xFE
  static __device__ __forceinline__ unsigned one(const long long* p, int ignore, int C) {
    const long long y = __ldg(p);
    if (y == (long long)ignore || y < 0) return (ignore >= 0 && ignore <= 255) ? (unsigned)ignore : 0xffu;
    return (y < (long long)C) ? (unsigned)y : 0xfeu;
  }
  static __device__ __forceinline__ RawRun issue(const long long* a, int n, bool, int ignore, int C) {
    unsigned long long v = ~0ULL;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < n) v = (v & ~(0xffULL << (8 * q))) | ((unsigned long long)one(a + q, ignore, C) << (8 * q));
    RawRun r;
    r.w0 = (unsigned)v; r.w1 = (unsigned)(v >> 32); r.w2 = 0xffffffffu; r.sh = 0u;
    return r;
  }
  static __device__ __forceinline__ void raw(const RawRun& r, unsigned& lo, unsigned& hi) { lo = r.w0; hi = r.w1; }
};

// rare path of the pairwise log (q0 q1 underflows)
__device__ __forceinline__ float log2_pair_slow(float q0, float q1) { return lg2_approx(q0) + lg2_approx(q1); }

// exact unsigned division by a divisor fixed per launch (host: make_fastdiv): 3 integer instructions
__device__ __forceinline__ unsigned fastdiv(unsigned n, const FastDiv d) {
  const unsigned t = __umulhi(d.mul, n);
  return (t + ((n - t) >> d.s1)) >> d.s2;
}

static constexpr int kEdgeRows = 16;  // pixel rows per cell-row whose edge column is staged in smem
static constexpr int kRun = 8;        // pixels of a run handled by the pipelined row body
static constexpr int kLamFwd = kRun + 2;    // entries a forward walk can read (two pixels of read-ahead)
static constexpr int kLamGuard = kRun + 2;  // zero entries a reversed walk of an empty run can reach

// kernel flavours: forward only / forward + backward (raw, host-scaled or device-scaled gradients) / Placeholder_loss
enum { K_FWD = 0, K_BWD = 1, K_PLACE = 2 };

template <int CPL, int LPR, int KMODE, typename LabelT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) head_kernel(const HeadArgs A) {
  static_assert(CPL % 2 == 0, "channels per lane are processed as fp32x2 pairs");
  constexpr bool BWD = (KMODE != K_FWD);
  constexpr bool PLACE = (KMODE == K_PLACE);   // Placeholder_loss: labels are derived from the logits, no T
  constexpr int NP = CPL / 2;    // channel pairs per lane
  constexpr int CKP = CPL * LPR;
  constexpr int CPW = 32 / LPR;  // cells per warp unit
  constexpr int NW = NT / 32;
  constexpr int ES = CKP + 2;    // edge-column row stride (floats, even: float2 stores)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CK = A.CK, C = A.C;
  // [NW][2][NP][32] float4: the cell's bilinear coefficients of every lane's channel pairs (log2 domain, shifted by
  // the cell-row's bound M), re-read once per pixel row:  a(ly) = A0 + ly A1,  d(ly) = D0 + ly D1,  t(lx) = a + lx d
  float4* Lsm = reinterpret_cast<float4*>(smem_raw);
  unsigned char* sp = smem_raw + (size_t)NW * 2 * NP * 32 * sizeof(float4);
  float* Ts = reinterpret_cast<float*>(sp);                           // [C][CKP] = -T^T
  sp += (size_t)(PLACE ? 0 : C) * CKP * 4;
  float* Esm = reinterpret_cast<float*>(sp);                          // [NW][kEdgeRows][ES] edge column (BWD)
  sp += BWD ? (size_t)NW * kEdgeRows * ES * 4 : 0;
  // [NW][kLamGuard + kLamFwd][32]: the horizontal lerp weights of every lane's run in walk order, read with a running
  // pointer (forward rows +1 entry per pixel, reversed rows -1); the guard entries below index 0 stay zero so that
  // lanes walking past the start of a short run read a harmless weight
  float* lam_sm = reinterpret_cast<float*>(sp);
  sp += PLACE ? 0 : (size_t)NW * (kLamGuard + kLamFwd) * 32 * 4;
  float* lx_tab = reinterpret_cast<float*>(sp);                       // [W] horizontal lerp weight of every pixel column
  float* ly_tab = lx_tab + A.W;                                       // [H] vertical lerp weight of every pixel row
  int* xs_tab = reinterpret_cast<int*>(ly_tab + A.H);                 // [ncx + 1] first pixel column of every cell
  int* ys_tab = xs_tab + (A.ncx + 1);                                 // [ncy + 1] first pixel row of every cell-row
  __shared__ double red_d[NW];
  __shared__ long long red_i[NW];
  __shared__ float s_gs;   // gradient scale applied by the kernel: 1 (raw), host scale, or grad_out / N_valid

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int sub = (LPR > 1) ? (tid & (LPR - 1)) : 0;
  const int kbase = sub * CPL;
  const int pidx = lane / LPR;  // this lane group's cell within the unit
  const LabelT* labels = reinterpret_cast<const LabelT*>(A.labels);
  const int h = A.h, w = A.w;
  const unsigned plane = (unsigned)(h * w);
  const bool ign_fits = (A.ignore >= 0 && A.ignore <= 255);
  const unsigned pad8 = ign_fits ? (unsigned)A.ignore : 0xffu;   // "ignored, silently"
  const bool ident = (A.T == nullptr);                            // plain CE: T = I
  // byte-parallel "label >= C" (see the row loop): C <= 128: ((x & 0x7f..) + (128 - C)) | x ; else ((x & 0x7f..) + (256 - C)) & x
  const unsigned pad4 = pad8 * 0x01010101u;
  const unsigned kge = (unsigned)(C <= 128 ? 128 - C : 256 - C) * 0x01010101u;
  const unsigned ge_or = C <= 128 ? 0xffffffffu : 0u;
  float4* Lw = Lsm + (size_t)(tid >> 5) * 2 * NP * 32 + lane;   // + (arr * NP + q) * 32
  float* Ew = Esm + (size_t)(tid >> 5) * kEdgeRows * ES;
  float* lamF = lam_sm + ((size_t)(tid >> 5) * (kLamGuard + kLamFwd) + kLamGuard) * 32 + lane;   // entry i at lamF[i * 32]

  // ---- one-time per CTA: -T transposed ([y][k], zero padded), pixel/cell tables ----
  for (int i = tid; i < (PLACE ? 0 : C * CKP); i += NT) {
    int y = i / CKP, k = i - y * CKP;
    float v = 0.f;
    if (k < CK) v = A.T ? __ldg(A.T + (size_t)k * C + y) : (k == y ? 1.f : 0.f);
    Ts[i] = -v;
  }
  if (tid == 0) {
    float gs = A.gscale;   // raw single pass: 1 ; host-scaled backward: the caller's scale
    if (A.count_local)     // step mode: the valid-pixel count was produced by head_prep_kernel before this launch
      gs = (float)((A.grad_out ? (double)__ldg(A.grad_out) : 1.0) / *A.count_local);
    s_gs = gs;
  }
  for (int i = tid; i <= A.ncx; i += NT) xs_tab[i] = first_px_of_cell(i, A.sx, A.ncx, A.W);
  for (int i = tid; i <= A.ncy; i += NT) ys_tab[i] = first_px_of_cell(i, A.sy, A.ncy, A.H);
  for (int i = tid; i < A.W; i += NT) lx_tab[i] = lambda_of(i, A.sx, cell_of(i, A.sx, A.ncx));
  for (int i = tid; i < A.H; i += NT) ly_tab[i] = lambda_of(i, A.sy, cell_of(i, A.sy, A.ncy));
  for (int i = tid; i < (PLACE ? 0 : NW * (kLamGuard + kLamFwd) * 32); i += NT) lam_sm[i] = 0.f;
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  float* ct = A.part_dT + (size_t)(smid % (unsigned)A.ntiles) * C * CKP;  // this SM's dT tile in global memory (L2 resident)
  __syncthreads();
  const float gs = s_gs;

  float2 D2[NP];     // dT accumulators (sum of e_k / s) for the thread's current label column
  float2 nTc[NP];    // -T[:, cur] for this lane's channels
#pragma unroll
  for (int q = 0; q < NP; ++q) { D2[q] = make_float2(0.f, 0.f); nTc[q] = make_float2(0.f, 0.f); }
  int cur = -1;
  double loss_d = 0.0;  // sum of log2 q over this thread's valid pixels
  long long cnt = 0;
  unsigned badf = 0;  // contract violation seen (a label that is neither a class nor the ignore label)

  // Dynamic unit scheduler: lane 0 claims, the id is broadcast through REDUX (its result is a uniform register, so the
  // loops below are provably warp-uniform).  Claims run TWO units ahead: the id of the next unit is known while the
  // current one is processed, so its logit rows can be prefetched into L1 behind the current unit's arithmetic.
  // (inline PTX: the compiler turns a one-lane atomicAdd into its warp-aggregated form, whose broadcast shuffle waits
  // for the atomic's round trip on the spot; this one stays in flight until its value is used)
  auto claim_raw = [&]() -> unsigned {
    unsigned r = 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0;\n\t@p atom.global.add.u32 %0, [%2], 1;\n\t}"
                 : "+r"(r) : "r"((unsigned)lane), "l"(A.counter) : "memory");
    return r;
  };
  auto uniform = [&](unsigned v) -> unsigned { return __reduce_max_sync(0xffffffffu, v); };
  struct UnitId { int b, gyi, ux, part, nparts_log2; };
  auto decode = [&](unsigned u) -> UnitId {
    UnitId r;
    unsigned g = u;
    r.part = 0; r.nparts_log2 = 0;
    if (u >= A.nbig) {   // the tail of the unit list: groups split into 2^rs_log2 row slices (finer load balance)
      const unsigned v = u - A.nbig;
      g = A.nbig + (v >> A.rs_log2);
      r.part = (int)(v & ((1u << A.rs_log2) - 1u));
      r.nparts_log2 = A.rs_log2;
    }
    const unsigned bg = fastdiv(g, A.div_img);        // image
    const unsigned rem = g - bg * A.groups_per_img;
    const unsigned gy = fastdiv(rem, A.div_ux);       // cell-row group
    r.b = (int)bg; r.gyi = (int)gy; r.ux = (int)(rem - gy * A.units_x);
    return r;
  };
  // the logit rows a unit starts with: 2 node rows x CK channels x (CPW + 1) floats, one or two 128-byte lines each
  auto prefetch_unit = [&](const UnitId& U) {
    const int cy0 = U.gyi * A.ur;
    for (int i = lane; i < 2 * CK; i += 32) {
      const int k = i >> 1, r = i & 1;
      const int gy = min(cy0 + r, h - 1);
      const float* p = A.logits + ((size_t)U.b * CK + k) * plane + (gy * w + min(U.ux * CPW, w - 1));
      const float* pe = p + min(CPW, w - 1 - min(U.ux * CPW, w - 1));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pe));
    }
  };
  // Hand the lane's dT accumulators (column `cur`) to the SM's tile: native red.global.add.v2.f32, fire and forget
  // (shared-memory float atomics are CAS loops on sm_100).  Padded channels carry exact zeros and the tile has CKP
  // columns, so pairs go out unguarded.
  auto flush_lane = [&]() {
    float* dst = ct + cur * CKP + kbase;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      red_add_v2(dst + 2 * q, D2[q]);
      D2[q] = make_float2(0.f, 0.f);
    }
  };

  unsigned unit = uniform(claim_raw());
  unsigned unit_n = uniform(claim_raw());      // the next unit
  unsigned raw_nn = claim_raw();               // the one after, in flight
  while (unit < A.nunits) {
    const UnitId U = decode(unit);
    if (A.prefetch && unit_n < A.nunits) prefetch_unit(decode(unit_n));
    const int b = U.b, ux = U.ux;
    const int cx = ux * CPW + pidx;
    const bool cell_ok = cx < A.ncx;
    const int xa = cell_ok ? xs_tab[cx] : 0;
    const int nrun = cell_ok ? xs_tab[cx + 1] - xa : 0;
    const int ncell_u = min(CPW, A.ncx - ux * CPW);            // cells of this unit (warp-uniform)
    const bool last_cell = pidx == ncell_u - 1;
    const int nmax = (int)uniform((unsigned)nrun);
    const int gx0 = min(cx, w - 1), gx1 = min(cx + 1, w - 1);
    const int edge_gx = min(ux * CPW + ncell_u, w - 1);          // node column right of the unit
    float loss_acc = 0.f;
    int cnt_u = 0;
    const int cy_begin = U.gyi * A.ur, cy_end = min(A.ncy, cy_begin + A.ur);
    // this unit's pixel rows: all rows of its cell-rows, or the part-th slice of them
    const int Yall0 = (int)uniform((unsigned)ys_tab[cy_begin]), Yall1 = (int)uniform((unsigned)ys_tab[cy_end]);
    const int Yfirst = Yall0 + (((Yall1 - Yall0) * U.part) >> U.nparts_log2);
    const int Ylast = Yall0 + (((Yall1 - Yall0) * (U.part + 1)) >> U.nparts_log2);  // one past the last row
    // horizontal lerp weights of the lane's run: they depend on the pixel column only
    float lam[kRun];
    if (PLACE) {
#pragma unroll
      for (int p = 0; p < kRun; ++p) lam[p] = lx_tab[min(xa + p, A.W - 1)];
    } else {
      __syncwarp();
#pragma unroll
      for (int p = 0; p < kLamFwd; ++p) lamF[p * 32] = lx_tab[min(xa + p, A.W - 1)];
      __syncwarp();
    }
    // labels: a per-lane row pointer advanced by W per row
    const LabelT* lrow = labels + (((long long)b * A.H + Yfirst) * A.W + xa);
    bool window_ok = false;
    if (sizeof(LabelT) == 1 && A.label_words_ok) {
      const uintptr_t last = reinterpret_cast<uintptr_t>(lrow + (long long)(Ylast - 1 - Yfirst) * A.W) & ~(uintptr_t)3;
      window_ok = last + 12 <= reinterpret_cast<uintptr_t>(labels + (long long)A.B * A.H * A.W);
    }
    // tail fill of the 8-label word: keep mask and pad bytes (fixed per unit)
    const unsigned long long keep64 = (nrun >= 8) ? ~0ULL : ((1ULL << (8 * nrun)) - 1ULL);
    const unsigned keep_lo = (unsigned)keep64, keep_hi = (unsigned)(keep64 >> 32);
    const unsigned fill_lo = pad4 & ~keep_lo, fill_hi = pad4 & ~keep_hi;
    RawRun raw_next = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u};
    if (!PLACE && Yfirst < Ylast) raw_next = LabelFetch<LabelT>::issue(lrow, nrun, window_ok, A.ignore, C);

    for (int cy = cy_begin; cy < cy_end; ++cy) {
      const int Y0 = max((int)uniform((unsigned)ys_tab[cy]), Yfirst);
      const int Y1 = min((int)uniform((unsigned)ys_tab[cy + 1]), Ylast);
      if (Y1 <= Y0) continue;  // warp-uniform
      const int gy0 = min(cy, h - 1), gy1 = min(cy + 1, h - 1);
      const bool edge_smem = BWD && (Y1 - Y0 <= kEdgeRows);
      float M;            // upper bound of every interpolated logit of this cell (log2 domain): the softmax shift
      bool range_safe;    // no pixel of this cell can underflow the softmax denominator (or, plain CE, its own term)
      // ---- stage the cell's bilinear coefficients of this lane's channels in the warp's smem slice ----
      {
        const float* q00 = A.logits + ((size_t)b * CK + kbase) * plane + (gy0 * w + gx0);
        const float* q01 = q00 + (gx1 - gx0);
        const float* q10 = q00 + (gy1 - gy0) * w;
        const float* q11 = q10 + (gx1 - gx0);
        __syncwarp();
        // all 4*CPL corner loads are issued before the first one is used (one exposed latency, not ten)
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
        float mx = -INFINITY, mlo = -INFINITY, mall = INFINITY;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const bool real = cell_ok && kbase + j < CK;
          const float sc = real ? kLog2e : 1.f;    // padding stays at kPadLogit / 0
          c00[j] *= sc; c01[j] *= sc; c10[j] *= sc; c11[j] *= sc;
          const float lo4 = fminf(fminf(c00[j], c01[j]), fminf(c10[j], c11[j]));
          mx = fmaxf(mx, fmaxf(fmaxf(c00[j], c01[j]), fmaxf(c10[j], c11[j])));
          mlo = fmaxf(mlo, lo4);                    // (padding: -1e30, no effect)
          mall = fminf(mall, real ? lo4 : INFINITY);
        }
        // every interpolated value of channel k lies between the min and the max of its 4 corners, so
        //   M = max_k max4 >= every pixel's max >= max_k min4 = mlo   and   exp-sum >= 2^(mlo - M)
        M = group_max<LPR>(mx, 0xffffffffu);
        mlo = group_max<LPR>(mlo, 0xffffffffu);
        mall = -group_max<LPR>(-mall, 0xffffffffu);
        range_safe = (M - mlo) < 38.f && (PLACE || !ident || (M - mall) < 60.f);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const float a0 = c00[2 * q], a1 = c00[2 * q + 1];
          const float e0 = c10[2 * q] - a0, e1 = c10[2 * q + 1] - a1;      // vertical difference at the left nodes
          const float b0 = c01[2 * q] - a0, b1 = c01[2 * q + 1] - a1;      // horizontal difference at the top nodes
          const float f0 = (c11[2 * q] - c10[2 * q]) - b0, f1 = (c11[2 * q + 1] - c10[2 * q + 1]) - b1;
          Lw[(0 * NP + q) * 32] = make_float4(a0 - M, a1 - M, e0, e1);     // padding: -1e30 - M stays -1e30
          Lw[(1 * NP + q) * 32] = make_float4(b0, b1, f0, f1);
        }
        __syncwarp();
      }
      // the statically unrolled row body needs runs of <= 8 pixels and a cell whose logits cannot underflow
      const bool fast = !PLACE && nmax <= kRun && uniform(range_safe ? 0u : 1u) == 0u;
      float2 Vt[NP], Vb[NP];
      if (BWD) {
#pragma unroll
        for (int q = 0; q < NP; ++q) { Vt[q] = make_float2(0.f, 0.f); Vb[q] = make_float2(0.f, 0.f); }
      }

      for (int Y = Y0; Y < Y1; ++Y) {
        const float ly = ly_tab[Y];
        // Rows alternate direction (boustrophedon): a run that contains a label boundary A|B is walked A..B on one
        // row and B..A on the next, so the lane changes its label column once per row instead of twice.
        const bool reverse = fast && A.boustrophedon && (((Y - Yall0) & 1) != 0);
        unsigned clo = 0u, chi = 0u;      // the run's 8 label bytes in walk order (tail padded)
        if (!PLACE) {
          unsigned rlo, rhi;
          LabelFetch<LabelT>::raw(raw_next, rlo, rhi);
          if (sizeof(LabelT) == 1 && !ign_fits) {
            // uint8 labels with an ignore label outside [0, 255]: a 255 byte is a contract violation, not padding
            const unsigned long long rawc = ((unsigned long long)rhi << 32) | rlo;
#pragma unroll 1
            for (int p = 0; p < 8; ++p) badf |= (unsigned)(p < nrun && ((rawc >> (8 * p)) & 0xffULL) == 0xffULL);
          }
          if (reverse) {   // walk order = the run's pixels right to left: byte i <- byte nrun-1-i
            const unsigned long long x = (((unsigned long long)rhi << 32) | rlo) << (8 * (8 - max(nrun, 1)));
            rlo = __byte_perm((unsigned)(x >> 32), 0u, 0x0123u);
            rhi = __byte_perm((unsigned)x, 0u, 0x0123u);
          }
          clo = (rlo & keep_lo) | fill_lo;
          chi = (rhi & keep_hi) | fill_hi;
          lrow += A.W;
          if (Y + 1 < Ylast)  // next row's labels are in flight during this row's arithmetic
            raw_next = LabelFetch<LabelT>::issue(lrow, nrun, window_ok, A.ignore, C);
        }
        // Row accumulators of the transposed horizontal lerp of g_k = e_k (1/sum - T_ky / s):  Gs = sum_p g,
        // G1 = sum_p lambda_p g.
        float2 Gs[NP], G1[NP];
        // vertical lerp once per row
        float2 a[NP], d[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const float4 va = Lw[(0 * NP + q) * 32], vd = Lw[(1 * NP + q) * 32];
          a[q] = ffma2(bcast2(ly), make_float2(va.z, va.w), make_float2(va.x, va.y));
          d[q] = ffma2(bcast2(ly), make_float2(vd.z, vd.w), make_float2(vd.x, vd.y));
        }

        if constexpr (PLACE) {
#pragma unroll
          for (int q = 0; q < NP; ++q) { Gs[q] = make_float2(0.f, 0.f); G1[q] = make_float2(0.f, 0.f); }
          // ---- Placeholder_loss (tools/trainV2_simt.py:202-230) on this row's pixels --------------------
          // Per pixel: a = arg-max channel (first on ties); valid iff a < C and max prob > thres;
          //   known   = -log softmax(z)_a
          //   unknown = CE(z', y) with z' = z except z'_a = 0 (a CONSTANT: `ones` at :208 is zeros_like), and
          //             y = the first best open-set channel if its logit is > 0, else class 0 (:220-222)
          // Both softmaxes are taken relative to their own exact maximum (no range assumptions).
          const float tz = -M;  // the logit 0 in this cell's shifted log2 domain
          auto pixel = [&](const float lamp, const bool wv) {
            float t[CPL], e[CPL], f[CPL];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const float2 tt = ffma2(bcast2(lamp), d[q], a[q]);
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
            float su = 0.f, sp2 = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              e[j] = ex2_approx(t[j] - m0);
              f[j] = (kbase + j == ia) ? 0.f : ex2_approx(t[j] - ms);
              su += e[j];
              sp2 += f[j];
            }
            su = group_sum<LPR>(su, 0xffffffffu);                          // >= 1; max prob = 1 / su
            sp2 = group_sum<LPR>(sp2, 0xffffffffu) + ex2_approx(tz - ms);  // >= 1
            const bool valid = wv && ia < C && (1.f > A.place_thres * su);
            const bool open_pos = mo > tz;                                  // an open-set logit > 0
            const int y = open_pos ? io : 0;
            const float t_first = __shfl_sync(0xffffffffu, t[0], lane & ~(LPR - 1));  // channel 0 of this pixel
            const float ty = (y == ia) ? tz : (open_pos ? mo : t_first);   // z'_y in the shifted domain
            if (valid) {
              loss_acc -= lg2_approx(su) + A.place_lambda * (lg2_approx(sp2) + (ms - ty));
              cnt_u += 1;
            }
            const float r = valid ? rcp_approx(su * sp2) : 0.f;
            const float rs = r * sp2, rp = A.place_lambda * (r * su);
            const float oa = valid ? 1.f : 0.f, oy = (valid && y != ia) ? A.place_lambda : 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              float g = fmaf(f[j], rp, e[j] * rs);
              g -= (kbase + j == ia) ? oa : 0.f;
              g -= (kbase + j == y) ? oy : 0.f;
              if (j & 1) { Gs[j >> 1].y += g; G1[j >> 1].y = fmaf(lamp, g, G1[j >> 1].y); }
              else       { Gs[j >> 1].x += g; G1[j >> 1].x = fmaf(lamp, g, G1[j >> 1].x); }
            }
          };
          if (nmax <= kRun) {
#pragma unroll
            for (int p = 0; p < kRun; p += 2) {
              if (p < nmax) {  // warp-uniform: lanes past their run execute predicated-off pixels
                pixel(lam[p], p < nrun);
                pixel(lam[p + 1], p + 1 < nrun);
              }
            }
          } else {
            for (int p = 0; p < nmax; ++p) pixel(lx_tab[min(xa + p, A.W - 1)], p < nrun);
          }
        } else {
          // ---- T-corrected CE on this row's pixels ------------------------------------------------------
          // Row-level byte-parallel label classification (bit 7 of every byte; no carries cross a byte):
          //   ge = label >= C,  ne = label != pad byte;  valid = !ge & ne,  contract violation = ge & ne.
          // The run's tail is filled with the pad byte, so it is neither.
          unsigned vlo, vhi;
          {
            const unsigned xs[2] = {clo, chi};
            unsigned vm[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const unsigned x = xs[i];
              const unsigned tge = (x & 0x7f7f7f7fu) + kge;
              const unsigned ge = (tge | (x & ge_or)) & (x | ge_or);     // C <= 128: tge | x ; C > 128: tge & x
              const unsigned y = x ^ pad4;
              const unsigned ne = (((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u;
              vm[i] = ~ge & ne;
              badf |= ge & ne;
            }
            vlo = vm[0]; vhi = vm[1];
            if (nrun <= 8) cnt_u += __popc(vlo) + __popc(vhi);   // (long runs are counted per pixel)
          }
          // A label change: the lane moves to T column y; the old column's dT accumulators go to the SM's tile.
          auto do_switch = [&](const int y) {
            const float2* src = reinterpret_cast<const float2*>(Ts + y * CKP + kbase);
            if (BWD && cur >= 0) flush_lane();
#pragma unroll
            for (int q = 0; q < NP; ++q) nTc[q] = src[q];
            cur = y;
          };
          float qe = 1.f;   // q of the pending even pixel: log2 q0 + log2 q1 = log2(q0 q1), one MUFU per pair
          // A pixel is processed in two stages so that consecutive pixels overlap (software pipeline, explicit double
          // buffer eA / eB): stage A is label independent (lerp, 10 ex2, exp-sum, 1/sum) and is issued for pixel i+1
          // in the same basic block as stage B of pixel i (T-mix, 1/s, the accumulations).
          struct PxBuf { float2 e[NP]; float rsv; };
          auto stage_a = [&](const float lamp, const bool v, PxBuf& X) {
            float2 (&e)[NP] = X.e;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const float2 t = ffma2(bcast2(lamp), d[q], a[q]);
              e[q] = make_float2(ex2_approx(t.x), ex2_approx(t.y));
            }
            float2 sum = e[0];
#pragma unroll
            for (int q = 1; q < NP; ++q) sum = fadd2(sum, e[q]);
            const float su = group_sum<LPR>(sum.x + sum.y, 0xffffffffu);
            X.rsv = v ? rcp_approx(su) : 0.f;
          };
          auto switch_check = [&](const unsigned c, const bool v) {
            const bool need = v && (int)c != cur;
            if (__any_sync(0xffffffffu, need)) {       // warp-uniform test: rare on coherent label maps
              if (need) do_switch((int)c);
            }
          };
          auto stage_b = [&](const bool v, const float lamp, const PxBuf& X, auto first_tag, auto odd_tag) {
            constexpr bool first = decltype(first_tag)::value;
            const float2 (&e)[NP] = X.e;
            const float rsv = X.rsv;
            float2 ns0 = fmul2(e[0], nTc[0]), ns1 = fmul2(e[NP > 1 ? 1 : 0], nTc[NP > 1 ? 1 : 0]);
#pragma unroll
            for (int q = 2; q < NP; ++q) {
              if (q & 1) ns1 = ffma2(e[q], nTc[q], ns1);
              else ns0 = ffma2(e[q], nTc[q], ns0);
            }
            const float2 ns = (NP > 1) ? fadd2(ns0, ns1) : ns0;
            const float s = group_sum<LPR>(-ns.x - ns.y, 0xffffffffu);
            const float isv = v ? rcp_approx(s) : 0.f;
            {
              const float q = v ? s * rsv : 1.f;               // q in (0, 1]; an invalid pixel contributes log 1
              if (decltype(odd_tag)::value) {
                const float qq = qe * q;
                if (qq > 1e-30f) loss_acc += lg2_approx(qq);
                else loss_acc += log2_pair_slow(qe, q);
              } else {
                qe = q;
              }
            }
            if (BWD) {
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                // c = (p_k - p_k T_ky / q) / e_k = 1/sum - T_ky / s ;  c1 = lambda c
                const float2 ca = ffma2(nTc[q], bcast2(isv), bcast2(rsv));
                const float2 c1 = fmul2(ca, bcast2(lamp));
                if (first) { Gs[q] = fmul2(e[q], ca); G1[q] = fmul2(e[q], c1); }
                else { Gs[q] = ffma2(e[q], ca, Gs[q]); G1[q] = ffma2(e[q], c1, G1[q]); }
                D2[q] = ffma2(e[q], bcast2(isv), D2[q]);   // e_k / s, column `cur`
              }
            }
          };
          // exact variant: softmax relative to the pixel's own maximum (cells whose logit range could underflow the
          // cell-level shift; plain CE with a huge margin; runs longer than 8 pixels)
          auto pixel_exact = [&](const unsigned c, const float lamp, const bool long_run) {
            const bool k = c < (unsigned)C, g = c != pad8;
            badf |= (unsigned)(!k & g);
            const bool v = k & g;
            if (v && (int)c != cur) do_switch((int)c);
            float2 t[NP], e[NP];
            float tm = -INFINITY;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              t[q] = ffma2(bcast2(lamp), d[q], a[q]);
              tm = fmaxf(tm, fmaxf(t[q].x, t[q].y));
            }
            tm = group_max<LPR>(tm, 0xffffffffu);
            float su = 0.f, s = 0.f, ty = 0.f;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              e[q] = make_float2(ex2_approx(t[q].x - tm), ex2_approx(t[q].y - tm));
              su += e[q].x + e[q].y;
              s -= e[q].x * nTc[q].x + e[q].y * nTc[q].y;
              // plain CE: the label's own logit (nTc is minus the one-hot column; padded channels have nTc = 0)
              ty -= (nTc[q].x != 0.f ? t[q].x * nTc[q].x : 0.f) + (nTc[q].y != 0.f ? t[q].y * nTc[q].y : 0.f);
            }
            su = group_sum<LPR>(su, 0xffffffffu);   // >= 1
            s = group_sum<LPR>(s, 0xffffffffu);
            const float rs = v ? rcp_approx(su) : 0.f;
            float is = 0.f;
            if (ident) {
              // q = e_y / sum may underflow although log q and the gradient p_k - [k == y] are finite (torch's
              // log_softmax): take log2 q = (t_y - max) - log2 sum and the one-hot term explicitly
              ty = group_sum<LPR>(ty, 0xffffffffu);
              if (v) loss_acc += (ty - tm) - lg2_approx(su);
            } else {
              is = v ? rcp_approx(s) : 0.f;
              if (v) loss_acc += lg2_approx(s * rs);
            }
            if (long_run) cnt_u += (int)v;   // short runs were counted from the label word
            if (BWD) {
              const float vf = v ? 1.f : 0.f;
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                float2 g;
                if (ident) g = ffma2(nTc[q], bcast2(vf), fmul2(e[q], bcast2(rs)));     // p_k - [k == y]
                else g = fmul2(e[q], ffma2(nTc[q], bcast2(is), bcast2(rs)));
                Gs[q] = fadd2(Gs[q], g);
                G1[q] = ffma2(g, bcast2(lamp), G1[q]);
                D2[q] = ffma2(e[q], bcast2(is), D2[q]);
              }
            }
          };

#pragma unroll
          for (int q = 0; q < NP; ++q) { Gs[q] = make_float2(0.f, 0.f); G1[q] = make_float2(0.f, 0.f); }
          if (fast) {
            // two pixels per iteration (compact loop: the hot code stays in the instruction cache); pixels past the
            // run's end carry the pad label and are predicated off
            const float* lp = lamF + (reverse ? (nrun - 1) * 32 : 0);
            const int lstep = reverse ? -32 : 32;
            PxBuf eA, eB;
            float lam0 = lp[0];
            stage_a(lam0, (vlo & 0x80u) != 0u, eA);
#pragma unroll 1
            for (int j = 0; j < nmax; j += 2) {
              const float lam1 = lp[lstep];
              const bool v0 = (vlo & 0x80u) != 0u, v1 = (vlo & 0x8000u) != 0u;
              switch_check(clo & 0xffu, v0);
              stage_a(lam1, v1, eB);
              stage_b(v0, lam0, eA, std::false_type{}, std::false_type{});
              lp += 2 * lstep;
              lam0 = lp[0];
              const unsigned c1 = __byte_perm(clo, 0u, 0x4441u);
              clo = __funnelshift_r(clo, chi, 16); chi >>= 16;
              vlo = __funnelshift_r(vlo, vhi, 16); vhi >>= 16;
              switch_check(c1, v1);
              stage_a(lam0, (vlo & 0x80u) != 0u, eA);
              stage_b(v1, lam1, eB, std::false_type{}, std::true_type{});
            }
          } else {
#pragma unroll 1
            for (int p = 0; p < nmax; ++p) {   // warp-uniform trip count; lanes past their run see the pad label
              unsigned c = pad8;
              if (p < nrun) {
                const unsigned long long codes = ((unsigned long long)chi << 32) | clo;
                c = (nrun <= 8) ? (unsigned)(codes >> (8 * p)) & 0xffu
                                : LabelFetch<LabelT>::one(lrow - A.W + p, A.ignore, C);
                if (sizeof(LabelT) == 1 && !ign_fits && nrun > 8 && c == 0xffu) badf = 1u;
              }
              pixel_exact(c, lx_tab[min(xa + p, A.W - 1)], nrun > 8);
            }
          }
        }  // !PLACE

        if (BWD) {
          // node column cx of this row = G0(cx) + G1(cx-1); the left neighbour is LPR lanes below
          const float lmask = (pidx == 0) ? 0.f : 1.f;
          const float wy0 = 1.f - ly;
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const float px = __shfl_up_sync(0xffffffffu, G1[q].x, LPR);
            const float py = __shfl_up_sync(0xffffffffu, G1[q].y, LPR);
            float2 n = ffma2(G1[q], bcast2(-1.f), Gs[q]);
            n = ffma2(make_float2(px, py), bcast2(lmask), n);
            Vt[q] = ffma2(bcast2(wy0), n, Vt[q]);
            Vb[q] = ffma2(bcast2(ly), n, Vb[q]);
          }
          // right edge of the unit: node column edge_gx belongs to the next unit (or is the image's
          // last column).  Its per-row values wait in the warp's smem slice until the cell-row is done.
          if (edge_smem) {
            if (last_cell) {
              float2* er = reinterpret_cast<float2*>(Ew + (Y - Y0) * ES + kbase);
#pragma unroll
              for (int q = 0; q < NP; ++q) er[q] = G1[q];
              if (sub == 0) Ew[(Y - Y0) * ES + CKP] = ly;
            }
          } else if (last_cell) {
            float* dst = A.dlogits + ((size_t)b * CK + kbase) * plane;
            const float w0y = (1.f - ly) * gs, w1y = ly * gs;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              if (kbase + j < CK) {
                float* pk = dst + (size_t)j * plane;
                const float g1 = (j & 1) ? G1[j >> 1].y : G1[j >> 1].x;
                atomicAdd(pk + gy0 * w + edge_gx, w0y * g1);
                atomicAdd(pk + gy1 * w + edge_gx, w1y * g1);
              }
            }
          }
        }
      }  // rows of the cell-row

      if (BWD) {
        float* dst = A.dlogits + (size_t)b * CK * plane;
        if (cell_ok) {
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
              const float g1 = Ew[r * ES + k], lyr = Ew[r * ES + CKP];
              et = fmaf(1.f - lyr, g1, et);
              eb = fmaf(lyr, g1, eb);
            }
            float* pk = dst + (size_t)k * plane;
            atomicAdd(pk + gy0 * w + edge_gx, et * gs);
            atomicAdd(pk + gy1 * w + edge_gx, eb * gs);
          }
          __syncwarp();
        }
      }
    }  // cell-rows of the unit

    loss_d += (double)loss_acc;
    cnt += cnt_u;
    unit = unit_n;
    unit_n = uniform(raw_nn);
    raw_nn = claim_raw();
  }

  // ---- CTA epilogue: partials ---------------------------------------------------------------
  if (BWD && !PLACE && cur >= 0) flush_lane();
  if (badf && A.err) atomicOr(A.err, SIMT_ERRBIT_LABEL_RANGE);
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


// Step prologue (MODE_STEP): zero dLogits and count this rank's valid pixels in ONE pass over the labels, so that the
// main kernel can apply grad_out / N_valid itself.  The last block to finish publishes the count in `count_local`.
// Validity is the main kernel's rule exactly: a class id below C that is not the ignore label.
template <typename LabelT>
__global__ void __launch_bounds__(256) head_prep_kernel(float* __restrict__ dlogits, long long n_dl,
                                                         const LabelT* __restrict__ labels, long long npix, int C,
                                                         int ignore, unsigned long long* __restrict__ accum,
                                                         unsigned long long* __restrict__ ticket,
                                                         double* __restrict__ count_local) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // ---- zero dLogits ----
  const long long n4 = ((reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) ? (n_dl >> 2) : 0;
  float4* d4 = reinterpret_cast<float4*>(dlogits);
  for (long long i = i0; i < n4; i += stride) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = n4 * 4 + i0; i < n_dl; i += stride) dlogits[i] = 0.f;
  // ---- count valid labels ----
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
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  __shared__ unsigned long long s_w[8];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long b = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) b += s_w[k];
    atomicAdd(accum, b);
    __threadfence();
    const unsigned long long t = atomicAdd(ticket, 1ULL);
    if (t == (unsigned long long)gridDim.x - 1ULL) {      // last block: every partial is in
      const unsigned long long total = atomicAdd(accum, 0ULL);
      *accum = 0ULL;
      *ticket = 0ULL;
      *count_local = (double)total;
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
    double* __restrict__ stats, float* __restrict__ loss_mean, float* __restrict__ dT_out, const int* __restrict__ err,
    const float* __restrict__ grad_out, const double* __restrict__ count_dev) {
  const int ndt = C * CKP;
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
      if (k < CK) {
        if (stats) stats[2 + k * C + y] = -t;
        // MODE_STEP on one GPU: grad_out / N_valid is already known on the device (count pass)
        const double sc = count_dev ? (grad_out ? (double)__ldg(grad_out) : 1.0) / *count_dev : (double)gscale;
        if (dT_out) dT_out[k * C + y] = (float)(-t * sc);
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
      if (count_dev) c = (long long)*count_dev;   // MODE_STEP: counted by head_prep_kernel, not by the main kernel
      *counter = 0ULL;  // the main kernel of this call has finished: re-arm the unit scheduler
      const double ls = -kLn2 * l;
      if (stats) { stats[0] = ls; stats[1] = (double)c; }
      if (loss_mean) {
        float m = (float)(ls / (double)c);  // 0/0 -> NaN like the reference's mean over nothing
        if (err && (*err & SIMT_ERRBIT_LABEL_RANGE)) m = nanf("");
        *loss_mean = m;
      }
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
static constexpr int kMaxGridPerSm = 8;   // G = SM count * 8 bounds the grid (loss / count partials are per CTA)
static constexpr int kMaxCKP = 64;

// Benchmark tuning (simt_head_set_tuning): process-global, read under the same mutex that guards the launch caches.
struct Tuning { int ur, small_pct, flags, lpr; };
static Tuning g_tuning = {0, 0, 0, 0};
static std::mutex g_head_mutex;   // guards g_tuning and the per-instantiation launch caches

struct Plan {
  int CPL, LPR, NT, MINB, CKP;
  size_t smem;
};

template <int CPL, int LPR, int KMODE, typename LabelT, int NT, int MINB>
static int launch_cfg(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
  auto kern = head_kernel<CPL, LPR, KMODE, LabelT, NT, MINB>;
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
  const long long need = ((long long)A.nunits + NT / 32 - 1) / (NT / 32);
  if (g > need) g = need;
  if (g < 1) g = 1;
  *grid_out = (int)g;
  prof_begin(st);
  kern<<<(int)g, NT, P.smem, st>>>(A);
  prof_end(st);
  return (int)cudaGetLastError();
}

// channel-count -> (CPL, LPR, threads, min CTAs/SM) instantiations
#ifndef SIMT_MINB_BWD
#define SIMT_MINB_BWD 3
#endif
#ifndef SIMT_MINB_FWD
#define SIMT_MINB_FWD 4
#endif
#ifdef SIMT_HEAD_BENCH_ONLY   /* development builds: only the bench instantiation (seconds instead of minutes) */
#define SIMT_HEAD_CONFIGS(X) X(10, 2, 128, SIMT_MINB_FWD, SIMT_MINB_BWD)
#else
#define SIMT_HEAD_CONFIGS(X) \
  X(10, 2, 128, SIMT_MINB_FWD, SIMT_MINB_BWD)        \
  X(12, 2, 128, 3, 2)                                \
  X(6, 4, 128, 4, 4)         \
  X(10, 4, 128, 4, 3)        \
  X(16, 4, 128, 3, 2)
#endif

template <int KMODE, typename LabelT>
static int dispatch(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
#define X(cpl, lpr, nt, minb_fwd, minb_bwd) \
  if (P.CPL == cpl && P.LPR == lpr)          \
    return launch_cfg<cpl, lpr, KMODE, LabelT, nt, (KMODE == K_FWD ? minb_fwd : minb_bwd)>(A, P, st, grid_out);
  SIMT_HEAD_CONFIGS(X)
#undef X
  return SIMT_EUNSUPPORTED;
}

// entry-point mode -> kernel flavour: the three gradient-producing modes share ONE instantiation (they differ in
// where the gradient scale comes from: HeadArgs::gscale / count_local)
static int dispatch_all(int mode, int label_bytes, const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
  if (mode == MODE_PLACE) return dispatch<K_PLACE, uint8_t>(A, P, st, grid_out);
#ifdef SIMT_HEAD_BENCH_ONLY
  if (label_bytes != 1) return SIMT_EUNSUPPORTED;
#else
  if (label_bytes != 1)
    return mode == MODE_FWD ? dispatch<K_FWD, long long>(A, P, st, grid_out) : dispatch<K_BWD, long long>(A, P, st, grid_out);
#endif
  return mode == MODE_FWD ? dispatch<K_FWD, uint8_t>(A, P, st, grid_out) : dispatch<K_BWD, uint8_t>(A, P, st, grid_out);
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
  Tuning tune;
  {
    std::lock_guard<std::mutex> lock(g_head_mutex);
    tune = g_tuning;
  }
  int rc = choose_config(CK, tune.lpr, P);
  if (rc) return rc;
  A->sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  A->sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  A->ncy = h > 1 ? h - 1 : 1;
  A->ncx = w > 1 ? w - 1 : 1;
  // cell-rows per group: ~8 pixel rows per group keeps the per-unit overhead amortised
  int ur = tune.ur;
  if (ur <= 0) {
    const double rows_per_cell = (double)H / (double)A->ncy;
    ur = (int)(8.0 / rows_per_cell + 0.5);
    if (ur < 1) ur = 1;
    if (ur > 32) ur = 32;
  }
  const int cpw = 32 / P->LPR;
  A->ur = ur;
  A->boustrophedon = (tune.flags & 1) ? 0 : 1;
  A->prefetch = (tune.flags & 2) ? 0 : 1;
  const long long ngy = (A->ncy + ur - 1) / ur;
  const long long units_x = (A->ncx + cpw - 1) / cpw;
  const long long groups = (long long)B * ngy * units_x;
  // The tail of the dynamically scheduled unit list is made of row slices of a group (halves or quarters) so that the
  // warps finish together; the head of the list stays whole groups (one staging of the cell's corners per 8 rows).
  // Only single-cell-row groups with several pixel rows are worth slicing.
  unsigned rs_log2 = 0;
  long long nsmall = 0;
  const double rows_per_group = (double)H / (double)ngy;
  if (ur == 1 && rows_per_group >= 4.0) {
    DeviceInfo di;
    if (device_info(&di)) return SIMT_EUNSUPPORTED;
    const long long warps = (long long)di.sm_count * 3 * (P->NT / 32);   // ~3 CTAs per SM resident
    rs_log2 = rows_per_group >= 8.0 ? 2 : 1;
    // default: as many sliced groups as there are warps (every warp ends on small units); small_pct overrides
    nsmall = tune.small_pct > 0 ? groups * (tune.small_pct > 100 ? 100 : tune.small_pct) / 100 : warps;
    if (tune.small_pct < 0) nsmall = 0;
    if (nsmall > groups) nsmall = groups;
  }
  const long long nunits = (groups - nsmall) + (nsmall << rs_log2);
  if (nunits > 0x7fff0000LL || groups > 0x7fff0000LL) return SIMT_EUNSUPPORTED;   // unit ids are 32-bit
  A->units_x = (unsigned)units_x;
  A->groups_per_img = (unsigned)(ngy * units_x);
  A->nbig = (unsigned)(groups - nsmall);
  A->rs_log2 = rs_log2;
  A->nunits = (unsigned)nunits;
  A->div_img = make_fastdiv(A->groups_per_img);
  A->div_ux = make_fastdiv(A->units_x);
  const bool bwd = mode != MODE_FWD;
  const size_t nw = (size_t)(P->NT / 32);
  P->smem = nw * 2 * (P->CPL / 2) * 32 * 16 + (mode == MODE_PLACE ? 0 : (size_t)C * P->CKP * 4) +
            (bwd ? nw * kEdgeRows * (P->CKP + 2) * 4 : 0) +
            (mode == MODE_PLACE ? 0 : nw * (kLamGuard + kLamFwd) * 32 * 4) + (size_t)(W + H) * 4 + (size_t)(A->ncx + A->ncy + 2) * 4;
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
  A.part_loss = reinterpret_cast<double*>(ws + 64);
  A.part_cnt = reinterpret_cast<long long*>(ws + 64 + G * 8);
  A.part_dT = reinterpret_cast<float*>(ws + 64 + G * 16);
  A.ntiles = di.sm_count < kFinSlices * kFinMaxPer ? di.sm_count : kFinSlices * kFinMaxPer;
  if (mode != MODE_FWD)
    SIMT_CUDA_TRY(cudaMemsetAsync(dlogits, 0, (size_t)B * CK * h * w * sizeof(float), st));
  int grid = 0;
  rc = dispatch_all(mode, label_bytes, A, P, st, &grid);
  if (rc) return rc;
  const int fgrid = (C * P.CKP + 31) / 32 + 1;
  head_finalize_kernel<<<fgrid, 1024, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, A.ntiles, CK, P.CKP, C, mode, gscale,
                                              A.counter, stats, loss_mean, dT_out, err_flag, nullptr, nullptr);
  return (int)cudaGetLastError();
}

// One whole training step of the head on one GPU (the path HeadRunner.step takes): label count + dLogits zeroing, the
// fused kernel applying the final scale, finalize.  Three launches, no pass over dLogits after the kernel.
static int run_step(const float* logits, int B, int CK, int h, int w, const float* T, int C, const void* labels,
                    int label_bytes, int H, int W, int ignore, const float* grad_out, float* dlogits, float* dT,
                    double* stats, float* loss_mean, int* err_flag, void* workspace, size_t workspace_bytes,
                    cudaStream_t st) {
  int rc = validate(logits, B, CK, h, w, C, labels, label_bytes, H, W, T);
  if (rc) return rc;
  if (!err_flag || !workspace || !dlogits || !stats) return SIMT_EINVAL;
  if (workspace_bytes < simt_head_workspace_bytes(B, CK, C, h, w, H, W)) return SIMT_EWORKSPACE;
  HeadArgs A{};
  Plan P{};
  A.logits = logits; A.T = T; A.labels = labels;
  A.B = B; A.CK = CK; A.C = C; A.h = h; A.w = w; A.H = H; A.W = W; A.ignore = ignore;
  A.gscale = 1.f; A.dlogits = dlogits; A.err = err_flag; A.grad_out = grad_out;
  A.label_words_ok = (label_bytes == 1 && (reinterpret_cast<uintptr_t>(labels) & 3) == 0 &&
                      (((long long)B * H * W) & 3) == 0) ? 1 : 0;
  rc = make_plan(MODE_STEP, B, CK, C, h, w, H, W, &A, &P);
  if (rc) return rc;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc) return rc;
  const size_t G = (size_t)di.sm_count * kMaxGridPerSm;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  A.counter = reinterpret_cast<unsigned long long*>(ws);
  unsigned long long* accum = reinterpret_cast<unsigned long long*>(ws + 8);     // the 64-byte header has room
  unsigned long long* ticket = reinterpret_cast<unsigned long long*>(ws + 16);
  double* count_local = reinterpret_cast<double*>(ws + 24);
  A.count_local = count_local;
  A.part_loss = reinterpret_cast<double*>(ws + 64);
  A.part_cnt = reinterpret_cast<long long*>(ws + 64 + G * 8);
  A.part_dT = reinterpret_cast<float*>(ws + 64 + G * 16);
  A.ntiles = di.sm_count < kFinSlices * kFinMaxPer ? di.sm_count : kFinSlices * kFinMaxPer;
  const long long n_dl = (long long)B * CK * h * w, npix = (long long)B * H * W;
  const int pgrid = di.sm_count * 4;
  if (label_bytes == 1)
    head_prep_kernel<uint8_t><<<pgrid, 256, 0, st>>>(dlogits, n_dl, static_cast<const uint8_t*>(labels), npix, C, ignore,
                                                     accum, ticket, count_local);
  else
    head_prep_kernel<long long><<<pgrid, 256, 0, st>>>(dlogits, n_dl, static_cast<const long long*>(labels), npix, C,
                                                       ignore, accum, ticket, count_local);
  SIMT_CUDA_TRY(cudaGetLastError());
  int grid = 0;
  rc = dispatch_all(MODE_STEP, label_bytes, A, P, st, &grid);
  if (rc) return rc;
  const int fgrid = (C * P.CKP + 31) / 32 + 1;
  head_finalize_kernel<<<fgrid, 1024, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, A.ntiles, CK, P.CKP, C, MODE_STEP,
                                              1.f, A.counter, stats, loss_mean, dT, err_flag, grad_out, count_local);
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
  A.part_loss = reinterpret_cast<double*>(ws + 64);
  A.part_cnt = reinterpret_cast<long long*>(ws + 64 + G * 8);
  A.part_dT = reinterpret_cast<float*>(ws + 64 + G * 16);
  A.ntiles = di.sm_count < kFinSlices * kFinMaxPer ? di.sm_count : kFinSlices * kFinMaxPer;
  SIMT_CUDA_TRY(cudaMemsetAsync(dlogits, 0, (size_t)B * CK * h * w * sizeof(float), st));
  int grid = 0;
  rc = dispatch_all(MODE_PLACE, 1, A, P, st, &grid);
  if (rc) return rc;
  // one block: loss / count partials and the scheduler re-arm (there are no dT tiles in this mode)
  head_finalize_kernel<<<1, 1024, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, 0, CK, P.CKP, C, MODE_PLACE, 1.f,
                                           A.counter, stats, loss_mean, nullptr, nullptr, nullptr, nullptr);
  return (int)cudaGetLastError();
}

}  // namespace simt

using namespace simt;

extern "C" {

size_t simt_head_workspace_bytes(int B, int CK, int C, int h, int w, int H, int W) {
  (void)B; (void)h; (void)w; (void)H; (void)W; (void)CK;
  DeviceInfo di;
  if (device_info(&di)) di.sm_count = 256;
  const size_t G = (size_t)di.sm_count * kMaxGridPerSm;
  return 64 + G * 16 + G * (size_t)(C > 0 ? C : 1) * kMaxCKP * sizeof(double);
}

void simt_head_set_tuning(int cell_rows_per_unit, int small_pct, int flags, int lpr) {
  std::lock_guard<std::mutex> lock(g_head_mutex);
  g_tuning = {cell_rows_per_unit, small_pct, flags, lpr};
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
                  err_flag, workspace, workspace_bytes, (cudaStream_t)stream);
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
