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
//                keeps the cell's 4 corner logits of its channels IN REGISTERS (pre-scaled
//                by log2 e), so per pixel row the vertical lerp is 2 FFMA/channel and per
//                pixel the horizontal lerp is ONE FFMA/channel:  t_k = a_k + lambda * d_k.
//   Softmax uses a per-row upper bound M of the logits instead of the per-pixel max (the
//   interpolant is a convex combination of the row's end points); a pixel whose exp-sum
//   underflows (only with > 2^40 dynamic range inside one cell) is redone with the exact max.
//   Backward: per pixel row the horizontal transposed lerp is accumulated in registers
//   (Gs = sum g, G1 = sum lambda g), the left neighbour's G1 arrives by one warp shuffle, and
//   the vertical transposed lerp is accumulated in registers too (Vt, Vb); per cell-row each
//   lane group adds its two node rows to dLogits with red.global.add.f32 (coalesced across the
//   warp); the unit's right edge column goes out the same way.
//   dT: per-thread register accumulators D[] for the thread's current label column; flushed
//   warp-collectively (shuffle tree, no atomics) into the warp's private fp64 shared tile
//   when the label changes at a row boundary and at the end of every unit; a label change
//   INSIDE a pixel run (rare on real label maps) takes a shared-memory atomic.
//   Per-CTA partials (loss, count, dT; all fp64) are reduced in a fixed order by a small
//   finalize kernel.  Units are claimed dynamically, so the grouping of the fp64 partial sums
//   (and the order of the fp32 red.adds into dLogits, as in torch's own CUDA backward of
//   upsample_bilinear2d) is not run-to-run deterministic in the last bit.
#include "common.cuh"

namespace simt {

enum { MODE_FWD = 0, MODE_FWDBWD = 1, MODE_BWD = 2 };

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr double kLn2 = 0.6931471805599453094;

struct HeadArgs {
  const float* logits;
  const float* T;  // may be null (identity)
  const void* labels;
  int B, CK, C, h, w, H, W, ignore;
  float sy, sx;    // torch's align_corners scales (float)(in-1)/(out-1)
  int ncy, ncx;    // number of cells = max(in-1, 1)
  int ur;          // cell-rows per unit
  int units_y, units_x;
  long long nunits;
  float gscale;
  float* dlogits;
  unsigned long long* counter;  // dynamic unit scheduler (zero on entry; finalize re-zeroes it)
  double* part_dT;      // [grid][C*CKP]
  double* part_loss;    // [grid]
  long long* part_cnt;  // [grid]
  int* err;
};

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

// Warp-collective flush of the per-thread dT accumulators D[] (column `old` of dT) into the
// warp's PRIVATE fp64 shared tile wt[y][k].  Called with the whole warp converged.  Lanes with
// the same label are summed with a shuffle tree and written by one lane per channel slice (plain
// read-modify-write: nobody else touches this warp's tile) -- float/double atomicAdd on shared
// memory is a CAS loop on sm_100 and collapses when a whole warp flushes the same hot class.
// More than two distinct labels in one flush (incoherent label maps) fall back to the CAS path,
// where the contention is spread over many addresses anyway.
template <int CPL, int LPR>
__device__ __forceinline__ void warp_flush_dT(float (&D)[CPL], int old, bool need, double* wt, int CKP, int CK,
                                              int kbase, int lane) {
  unsigned m = __ballot_sync(0xffffffffu, need);
#pragma unroll 1
  for (int round = 0; m != 0u && round < 2; ++round) {
    const int src = __ffs(m) - 1;
    const int lab = __shfl_sync(0xffffffffu, old, src);
    const bool mine = need && old == lab;
    const unsigned grp = __ballot_sync(0xffffffffu, mine);
    float tot[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      float v = mine ? D[j] : 0.f;
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      tot[j] = v;
      if (mine) D[j] = 0.f;
    }
    if (lane < LPR) {  // lane `sub` writes its channel slice
      double* dst = wt + lab * CKP + kbase;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (kbase + j < CK) dst[j] += (double)tot[j];
    }
    if (mine) need = false;
    m &= ~grp;
    __syncwarp();
  }
  if (need) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      if (kbase + j < CK) atomicAdd(&wt[old * CKP + kbase + j], (double)D[j]);
      D[j] = 0.f;
    }
  }
  __syncwarp();
}

template <int CPL, int LPR, int MODE, typename LabelT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) head_kernel(const HeadArgs A) {
  constexpr bool BWD = (MODE != MODE_FWD);
  constexpr int CKP = CPL * LPR;
  constexpr int CPW = 32 / LPR;  // cells per warp unit
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CK = A.CK, C = A.C;
  double* tiles = reinterpret_cast<double*>(smem_raw);                               // [NT/32][C*CKP] (BWD)
  float* Ts = reinterpret_cast<float*>(smem_raw + (BWD ? (size_t)(NT / 32) * C * CKP * 8 : 0));  // [C][CKP]
  __shared__ double red_d[NT / 32];
  __shared__ long long red_i[NT / 32];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int sub = (LPR > 1) ? (tid & (LPR - 1)) : 0;
  const unsigned gmask = (LPR == 1) ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
  const int kbase = sub * CPL;
  const int pidx = lane / LPR;  // this lane group's cell within the unit
  const LabelT* labels = reinterpret_cast<const LabelT*>(A.labels);
  const int h = A.h, w = A.w;

  // ---- one-time per CTA: T transposed ([y][k], zero padded) and the per-warp dT tiles ----
  for (int i = tid; i < C * CKP; i += NT) {
    int y = i / CKP, k = i - y * CKP;
    float v = 0.f;
    if (k < CK) v = A.T ? __ldg(A.T + (size_t)k * C + y) : (k == y ? 1.f : 0.f);
    Ts[i] = v;
  }
  if (BWD)
    for (int i = tid; i < (NT / 32) * C * CKP; i += NT) tiles[i] = 0.0;
  double* wt = tiles + (size_t)(tid >> 5) * C * CKP;
  __syncthreads();

  float D[CPL];   // dT accumulators for the thread's current label column
  float Tc[CPL];  // T[:, cur] for this lane's channels
#pragma unroll
  for (int j = 0; j < CPL; ++j) { D[j] = 0.f; Tc[j] = 0.f; }
  int cur = -1;
  double loss_d = 0.0;  // sum of log2 q over this thread's valid pixels
  long long cnt = 0;
  bool bad_label = false;

  auto claim = [&]() -> long long {
    unsigned long long u = 0;
    if (lane == 0) u = atomicAdd(A.counter, 1ULL);
    return (long long)__shfl_sync(0xffffffffu, u, 0);
  };

  long long unit = claim();
  while (unit < A.nunits) {
    const long long next_unit = claim();  // in flight while this unit is processed
    const int per_img = A.units_y * A.units_x;
    const int b = (int)(unit / per_img);
    const int urem = (int)(unit - (long long)b * per_img);
    const int uy = urem / A.units_x, ux = urem - uy * A.units_x;
    const int cx = ux * CPW + pidx;
    const bool cell_ok = cx < A.ncx;
    const int xa = cell_ok ? first_px_of_cell(cx, A.sx, A.ncx, A.W) : 0;
    const int xb = cell_ok ? first_px_of_cell(cx + 1, A.sx, A.ncx, A.W) : 0;
    const bool last_cell = cell_ok && (pidx == CPW - 1 || cx == A.ncx - 1);
    const int nrun = xb - xa;
    const int nmax = __reduce_max_sync(0xffffffffu, nrun);
    const int gx0 = min(cx, w - 1), gx1 = min(cx + 1, w - 1);
    float loss_acc = 0.f;

    const int cy_end = min(A.ncy, (uy + 1) * A.ur);
    for (int cy = uy * A.ur; cy < cy_end; ++cy) {
      const int Y0 = first_px_of_cell(cy, A.sy, A.ncy, A.H);
      const int Y1 = first_px_of_cell(cy + 1, A.sy, A.ncy, A.H);
      if (Y1 <= Y0) continue;  // warp-uniform
      const int gy0 = min(cy, h - 1), gy1 = min(cy + 1, h - 1);
      // ---- the cell's 4 corners for this lane's channels, in registers (log2 domain) ----
      float l0[CPL], dl0[CPL], l1[CPL], dl1[CPL];
      {
        const float* src = A.logits + ((size_t)b * CK + kbase) * h * w;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
          if (cell_ok && kbase + j < CK) {
            const float* p = src + (size_t)j * h * w;
            c00 = __ldg(p + gy0 * w + gx0); c01 = __ldg(p + gy0 * w + gx1);
            c10 = __ldg(p + gy1 * w + gx0); c11 = __ldg(p + gy1 * w + gx1);
          }
          l0[j] = c00 * kLog2e; dl0[j] = (c10 - c00) * kLog2e;
          l1[j] = c01 * kLog2e; dl1[j] = (c11 - c01) * kLog2e;
        }
      }
      float Vt[CPL], Vb[CPL];
      if (BWD) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { Vt[j] = 0.f; Vb[j] = 0.f; }
      }

      for (int Y = Y0; Y < Y1; ++Y) {
        const float ly = lambda_of(Y, A.sy, cy);
        const long long rowbase = ((long long)b * A.H + Y) * A.W;
        if (Y + 1 < Y1 && xb > xa)  // next row's labels on their way
          asm volatile("prefetch.global.L2 [%0];" ::"l"(labels + rowbase + A.W + xa));
        if (BWD) {
          // Row boundary, warp converged: if the run's first valid label differs from the column
          // the D[] accumulators belong to, flush them collectively before switching column.
          int first_lab = -1;
          for (int X = xa; X < xb; ++X) {
            const int y = load_label<LabelT>(labels, rowbase + X);
            if (y != A.ignore && y >= 0 && y < C) { first_lab = y; break; }
          }
          const bool sw = first_lab >= 0 && first_lab != cur;
          const bool need = sw && cur >= 0;
          if (__any_sync(0xffffffffu, need)) warp_flush_dT<CPL, LPR>(D, cur, need, wt, CKP, CK, kbase, lane);
          if (sw) {
            cur = first_lab;
#pragma unroll
            for (int j = 0; j < CPL; ++j) Tc[j] = Ts[cur * CKP + kbase + j];
          }
        }
        float Gs[CPL], G1[CPL];
        if (BWD) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) { Gs[j] = 0.f; G1[j] = 0.f; }
        }
        {
          // vertical lerp once per row; a = v0 - M, d = v1 - v0
          float a[CPL], d[CPL];
          float M = -INFINITY;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float v0 = fmaf(ly, dl0[j], l0[j]);
            const float v1 = fmaf(ly, dl1[j], l1[j]);
            a[j] = v0;
            d[j] = v1 - v0;
            if (kbase + j < CK) M = fmaxf(M, fmaxf(v0, v1));
          }
          M = group_max<LPR>(M, 0xffffffffu);
#pragma unroll
          for (int j = 0; j < CPL; ++j) a[j] = (kbase + j < CK) ? a[j] - M : -INFINITY;

          // The pixel loop is WARP-UNIFORM (nmax iterations, invalid lanes predicated off): the lane
          // groups exchange partial sums with full-mask shuffles, which must not sit behind a
          // divergent `continue` (group-masked shuffles serialise the 32/LPR groups).
          for (int i = 0; i < nmax; ++i) {
            const int X = xa + i;
            const bool inb = i < nrun;
            const int y = inb ? load_label<LabelT>(labels, rowbase + X) : A.ignore;
            bool valid = inb && y != A.ignore && y >= 0;
            if (valid && y >= C) { bad_label = true; valid = false; }
            if (valid && y != cur) {
              // label changed INSIDE a run (rare on coherent maps): per-lane flush
              if (BWD && cur >= 0) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                  if (kbase + j < CK) atomicAdd(&wt[cur * CKP + kbase + j], (double)D[j]);
                  D[j] = 0.f;
                }
              }
              cur = y;
#pragma unroll
              for (int j = 0; j < CPL; ++j) Tc[j] = Ts[y * CKP + kbase + j];
            }
            const float lam = lambda_of(X, A.sx, cx);
            float e[CPL];
            float sum0 = 0.f, sum1 = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              e[j] = ex2_approx(fmaf(lam, d[j], a[j]));
              if (j & 1) { sum1 += e[j]; s1 = fmaf(e[j], Tc[j], s1); }
              else       { sum0 += e[j]; s0 = fmaf(e[j], Tc[j], s0); }
            }
            float sum = group_sum<LPR>(sum0 + sum1, 0xffffffffu);
            float s = s0 + s1;
            if (__any_sync(0xffffffffu, valid && sum < 1e-12f)) {
              // the row-level bound M was far above some pixel's true max: redo with the exact max
              float tm = -INFINITY;
#pragma unroll
              for (int j = 0; j < CPL; ++j) tm = fmaxf(tm, fmaf(lam, d[j], a[j]));
              tm = group_max<LPR>(tm, 0xffffffffu);
              sum = 0.f; s = 0.f;
#pragma unroll
              for (int j = 0; j < CPL; ++j) {
                e[j] = ex2_approx(fmaf(lam, d[j], a[j]) - tm);
                sum += e[j];
                s = fmaf(e[j], Tc[j], s);
              }
              sum = group_sum<LPR>(sum, 0xffffffffu);
            }
            s = group_sum<LPR>(s, 0xffffffffu);
            const float rsum = valid ? rcp_approx(sum) : 0.f;
            if (MODE != MODE_BWD && sub == 0 && valid) loss_acc += lg2_approx(s * rsum);
            cnt += (sub == 0 && valid);
            if (BWD) {
              const float is = valid ? rcp_approx(s) : 0.f;
              const float lis = lam * is, lrs = lam * rsum;
#pragma unroll
              for (int j = 0; j < CPL; ++j) {
                const float c = fmaf(-Tc[j], is, rsum);     // (p_k - p_k T_ky / q) / e_k
                const float c1 = fmaf(-Tc[j], lis, lrs);    // lambda * c
                Gs[j] = fmaf(e[j], c, Gs[j]);
                G1[j] = fmaf(e[j], c1, G1[j]);
                D[j] = fmaf(e[j], is, D[j]);                // p_k / q
              }
            }
          }
        }
        if (BWD) {
          // node column cx of this row = G0(cx) + G1(cx-1); the left neighbour is LPR lanes below
          __syncwarp();
          const float wy1 = ly, wy0 = 1.f - ly;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            float prev = __shfl_up_sync(0xffffffffu, G1[j], LPR);
            if (pidx == 0) prev = 0.f;
            const float n = (Gs[j] - G1[j]) + prev;
            Vt[j] = fmaf(wy0, n, Vt[j]);
            Vb[j] = fmaf(wy1, n, Vb[j]);
          }
          if (last_cell) {
            // right edge of the unit: node column cx+1 belongs to the next unit (or is the image's
            // last column); add this row's share directly
            float* dst = A.dlogits + ((size_t)b * CK + kbase) * h * w;
            const float gs = (MODE == MODE_BWD) ? A.gscale : 1.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              if (kbase + j < CK) {
                float* pk = dst + (size_t)j * h * w;
                atomicAdd(pk + gy0 * w + gx1, wy0 * G1[j] * gs);
                atomicAdd(pk + gy1 * w + gx1, wy1 * G1[j] * gs);
              }
            }
          }
        }
      }  // rows of the cell-row

      if (BWD && cell_ok) {
        float* dst = A.dlogits + ((size_t)b * CK + kbase) * h * w;
        const float gs = (MODE == MODE_BWD) ? A.gscale : 1.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          if (kbase + j < CK) {
            float* pk = dst + (size_t)j * h * w;
            atomicAdd(pk + gy0 * w + gx0, Vt[j] * gs);
            atomicAdd(pk + gy1 * w + gx0, Vb[j] * gs);
          }
        }
      }
    }  // cell-rows of the unit

    // per-unit hand-off: fp32 partials of a unit are summed in a fixed order; across units in fp64
    if (BWD) {
      warp_flush_dT<CPL, LPR>(D, cur, cur >= 0, wt, CKP, CK, kbase, lane);
    }
    loss_d += (double)loss_acc;
    unit = next_unit;
  }

  // ---- CTA epilogue: partials ---------------------------------------------------------------
  if (bad_label) atomicOr(A.err, SIMT_ERRBIT_LABEL_RANGE);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss_d += __shfl_xor_sync(0xffffffffu, loss_d, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { red_d[tid >> 5] = loss_d; red_i[tid >> 5] = cnt; }
  __syncthreads();
  if (tid == 0) {
    double tl = 0; long long tc = 0;
    for (int wv = 0; wv < NT / 32; ++wv) { tl += red_d[wv]; tc += red_i[wv]; }
    A.part_loss[blockIdx.x] = tl;
    A.part_cnt[blockIdx.x] = tc;
  }
  if (BWD) {
    double* pd = A.part_dT + (size_t)blockIdx.x * C * CKP;
    for (int i = tid; i < C * CKP; i += NT) {
      double v = 0.0;
#pragma unroll
      for (int wv = 0; wv < NT / 32; ++wv) v += tiles[(size_t)wv * C * CKP + i];
      pd[i] = v;
    }
  }
}

// One warp per output: outputs 0 .. C*CKP-1 are dT entries ([y][k] layout of the partials),
// then loss and count.  Fixed summation order over the CTA partials.
__global__ void __launch_bounds__(256) head_finalize_kernel(
    const double* __restrict__ part_dT, const double* __restrict__ part_loss, const long long* __restrict__ part_cnt,
    int nparts, int CK, int CKP, int C, int mode, float gscale, unsigned long long* __restrict__ counter,
    double* __restrict__ stats, float* __restrict__ loss_mean, float* __restrict__ dT_out, const int* __restrict__ err) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int ndt = C * CKP;
  if (warp < ndt) {
    const int y = warp / CKP, k = warp - y * CKP;
    if (k >= CK) return;
    double s = 0;
    if (mode != MODE_FWD)
      for (int g = lane; g < nparts; g += 32) s += part_dT[(size_t)g * ndt + warp];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      if (stats) stats[2 + k * C + y] = -s;
      if (dT_out) dT_out[k * C + y] = (float)(-s * (double)gscale);
    }
  } else if (warp == ndt) {
    double l = 0; long long c = 0;
    for (int g = lane; g < nparts; g += 32) { l += part_loss[g]; c += part_cnt[g]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
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
// workspace: [counter u64 (+pad to 64 B)][part_loss f64 x G][part_cnt i64 x G][part_dT f64 x G*C*CKPmax]
static constexpr int kMaxGridPerSm = 8;
static constexpr int kMaxCKP = 64;

struct Tuning { int ur, unused, threads, lpr; };
static Tuning g_tuning = {0, 0, 0, 0};

struct Plan {
  int CPL, LPR, NT, MINB, CKP;
  size_t smem;
};

template <int CPL, int LPR, int MODE, typename LabelT, int NT, int MINB>
static int launch_cfg(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
  auto kern = head_kernel<CPL, LPR, MODE, LabelT, NT, MINB>;
  // per-instantiation cache of the attribute / occupancy queries (keyed by device and smem size)
  static int c_dev = -1, c_occ = 0;
  static size_t c_smem = 0;
  int dev = 0;
  SIMT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev != c_dev || P.smem != c_smem) {
    SIMT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    SIMT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c_occ, kern, NT, P.smem));
    c_dev = dev;
    c_smem = P.smem;
  }
  const int occ = c_occ;
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
  prof_begin(st);
  kern<<<(int)g, NT, P.smem, st>>>(A);
  prof_end(st);
  return (int)cudaGetLastError();
}

// channel-count -> (CPL, LPR, threads, min CTAs/SM) instantiations
#define SIMT_HEAD_CONFIGS(X) \
  X(10, 2, 128, 3)           \
  X(6, 4, 128, 4)            \
  X(9, 4, 128, 3)            \
  X(16, 4, 128, 2)

template <int MODE, typename LabelT>
static int dispatch(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
#define X(cpl, lpr, nt, minb) \
  if (P.CPL == cpl && P.LPR == lpr) return launch_cfg<cpl, lpr, MODE, LabelT, nt, minb>(A, P, st, grid_out);
  SIMT_HEAD_CONFIGS(X)
#undef X
  return SIMT_EUNSUPPORTED;
}

static int dispatch_all(int mode, int label_bytes, const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
  if (label_bytes == 1) {
    if (mode == MODE_FWD) return dispatch<MODE_FWD, uint8_t>(A, P, st, grid_out);
    if (mode == MODE_FWDBWD) return dispatch<MODE_FWDBWD, uint8_t>(A, P, st, grid_out);
    return dispatch<MODE_BWD, uint8_t>(A, P, st, grid_out);
  }
  if (mode == MODE_FWD) return dispatch<MODE_FWD, long long>(A, P, st, grid_out);
  if (mode == MODE_FWDBWD) return dispatch<MODE_FWDBWD, long long>(A, P, st, grid_out);
  return dispatch<MODE_BWD, long long>(A, P, st, grid_out);
}

static int choose_config(int CK, int lpr_req, Plan* P) {
  struct Cfg { int cpl, lpr, nt, minb; };
  static const Cfg cfgs[] = {
#define X(cpl, lpr, nt, minb) {cpl, lpr, nt, minb},
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
  int rc = choose_config(CK, g_tuning.lpr, P);
  if (rc) return rc;
  A->sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  A->sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  A->ncy = h > 1 ? h - 1 : 1;
  A->ncx = w > 1 ? w - 1 : 1;
  // cell-rows per unit: ~8 pixel rows per unit keeps the per-unit overhead amortised
  int ur = g_tuning.ur;
  if (ur <= 0) {
    const double rows_per_cell = (double)H / (double)A->ncy;
    ur = (int)(8.0 / rows_per_cell + 0.5);
    if (ur < 1) ur = 1;
    if (ur > 32) ur = 32;
  }
  const int cpw = 32 / P->LPR;
  A->ur = ur;
  A->units_y = (A->ncy + ur - 1) / ur;
  A->units_x = (A->ncx + cpw - 1) / cpw;
  A->nunits = (long long)B * A->units_y * A->units_x;
  const bool bwd = mode != MODE_FWD;
  P->smem = (bwd ? (size_t)(P->NT / 32) * C * P->CKP * 8 : 0) + (size_t)C * P->CKP * 4;
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
  A.part_dT = reinterpret_cast<double*>(ws + 64 + G * 16);
  if (mode != MODE_FWD)
    SIMT_CUDA_TRY(cudaMemsetAsync(dlogits, 0, (size_t)B * CK * h * w * sizeof(float), st));
  int grid = 0;
  rc = dispatch_all(mode, label_bytes, A, P, st, &grid);
  if (rc) return rc;
  const int nwarps = C * P.CKP + 1;
  const int fgrid = (nwarps * 32 + 255) / 256;
  head_finalize_kernel<<<fgrid, 256, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, CK, P.CKP, C, mode, gscale,
                                              A.counter, stats, loss_mean, dT_out, err_flag);
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

void simt_head_set_tuning(int cell_rows_per_unit, int reserved, int threads, int lpr) {
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
