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
// is bound by the MUFU (one ex2 per channel per pixel) and FP32 pipes, not by
// HBM (5.6 B/pixel algorithmic) -- see DESIGN.md.
//
// Work decomposition
//   low-res "cell" (cy, cx) = the square between 4 neighbouring low-res nodes;
//   every high-res pixel lies in exactly one cell (torch's i0 = min(floor(src), in-1)
//   is re-expressed as cell = min(floor(src), in-2), lambda = clamp(src - cell, 0, 1),
//   which gives identical values: for the last node lambda becomes exactly 1).
//   CTA tile   = tcy x tcx cells; its (tcy+1) x (tcx+1) x CK low-res nodes are staged
//                in shared memory, pre-multiplied by log2(e).
//   work item  = one pixel row inside one cell (a "run" of ~8 pixels at 65x129 ->
//                512x1024).  The vertical lerp is done ONCE per run, the horizontal
//                lerp is one FFMA per pixel per channel:  t_k = a_k + lambda * d_k.
//   LPR lanes share a run, each owning CPL channels (CK <= CPL*LPR).
//   Backward: per run the horizontal transposed lerp is accumulated in registers,
//   neighbouring runs are merged with one warp shuffle and written to a
//   shared-memory [k][row][node] array; phase 2 applies the vertical transposed lerp
//   and adds the tile's nodes into dLogits (red.global.add.f32 on tile borders).
//   dT: per-thread register accumulators for the thread's current label column,
//   flushed to a per-CTA shared tile when the label changes.
//   Per-CTA partials (loss, count, dT) are reduced in a fixed order by a small
//   finalize kernel, so loss / dT are run-to-run deterministic.
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
  int tcy, tcx, tcx_log2;
  int tiles_y, tiles_x;
  long long ntiles;
  int max_rows;
  float gscale;
  float* dlogits;
  float* part_dT;       // [grid][C*CKP]
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

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct SmemLayout {
  size_t Ls, Ns, Es, Ts, dTs, rowlam, rowcell, xs, rs, red, total;
};
__host__ __device__ inline SmemLayout smem_layout(int CK, int CKP, int C, int tcy, int tcx, int max_rows, bool bwd) {
  SmemLayout L;
  size_t o = 0;
  L.Ls = o;  o += align_up((size_t)CK * (tcy + 1) * (tcx + 1) * 4, 16);
  L.Ns = o;  o += bwd ? align_up((size_t)CK * max_rows * tcx * 4, 16) : 0;
  L.Es = o;  o += bwd ? align_up((size_t)CK * max_rows * 4, 16) : 0;
  L.Ts = o;  o += align_up((size_t)C * CKP * 4, 16);
  L.dTs = o; o += bwd ? align_up((size_t)C * CKP * 4, 16) : 0;
  L.rowlam = o;  o += align_up((size_t)max_rows * 4, 16);
  L.rowcell = o; o += align_up((size_t)max_rows * 4, 16);
  L.xs = o;  o += align_up((size_t)(tcx + 1) * 4, 16);
  L.rs = o;  o += align_up((size_t)(tcy + 2) * 4, 16);
  L.red = o; o += 32 * 8 + 32 * 8;
  L.total = o;
  return L;
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

template <int CPL, int LPR, int MODE, typename LabelT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) head_kernel(const HeadArgs A) {
  constexpr bool BWD = (MODE != MODE_FWD);
  constexpr int CKP = CPL * LPR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CK = A.CK, C = A.C;
  const int tcy = A.tcy, tcx = A.tcx;
  const SmemLayout SL = smem_layout(CK, CKP, C, tcy, tcx, A.max_rows, BWD);
  float* Ls = reinterpret_cast<float*>(smem_raw + SL.Ls);
  float* Ns = reinterpret_cast<float*>(smem_raw + SL.Ns);
  float* Es = reinterpret_cast<float*>(smem_raw + SL.Es);
  float* Ts = reinterpret_cast<float*>(smem_raw + SL.Ts);
  float* dTs = reinterpret_cast<float*>(smem_raw + SL.dTs);
  float* rowlam = reinterpret_cast<float*>(smem_raw + SL.rowlam);
  int* rowcell = reinterpret_cast<int*>(smem_raw + SL.rowcell);
  int* xs = reinterpret_cast<int*>(smem_raw + SL.xs);
  int* rs = reinterpret_cast<int*>(smem_raw + SL.rs);
  double* red_d = reinterpret_cast<double*>(smem_raw + SL.red);
  long long* red_i = reinterpret_cast<long long*>(smem_raw + SL.red + 32 * 8);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int sub = (LPR > 1) ? (tid & (LPR - 1)) : 0;
  const unsigned gmask = (LPR == 1) ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
  const int kbase = sub * CPL;
  const LabelT* labels = reinterpret_cast<const LabelT*>(A.labels);
  const int pitchL = tcx + 1;
  const int planeL = (tcy + 1) * pitchL;

  // ---- one-time per CTA: T transposed ([y][k], zero padded) and the dT tile ---------
  for (int i = tid; i < C * CKP; i += NT) {
    int y = i / CKP, k = i - y * CKP;
    float v = 0.f;
    if (k < CK) v = A.T ? __ldg(A.T + (size_t)k * C + y) : (k == y ? 1.f : 0.f);
    Ts[i] = v;
    if (BWD) dTs[i] = 0.f;
  }

  float D[CPL];  // dT accumulators for the thread's current label column
  float Tc[CPL]; // T[:, cur] for this lane's channels
#pragma unroll
  for (int j = 0; j < CPL; ++j) { D[j] = 0.f; Tc[j] = 0.f; }
  int cur = -1;
  float loss_acc = 0.f;  // sum of log2 q over this thread's valid pixels
  int cnt = 0;
  bool bad_label = false;

  for (long long tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
    const int tiles_per_img = A.tiles_y * A.tiles_x;
    const int b = (int)(tile / tiles_per_img);
    const int trem = (int)(tile - (long long)b * tiles_per_img);
    const int tyi = trem / A.tiles_x, txi = trem - tyi * A.tiles_x;
    const int cy0 = tyi * tcy, cx0 = txi * tcx;
    const int ncy_t = min(tcy, A.ncy - cy0), ncx_t = min(tcx, A.ncx - cx0);
    const int Y0 = first_px_of_cell(cy0, A.sy, A.ncy, A.H);
    const int Y1 = first_px_of_cell(cy0 + ncy_t, A.sy, A.ncy, A.H);
    const int rows = Y1 - Y0;

    __syncthreads();  // previous tile fully consumed (also orders the Ts/dTs init)
    // ---- tile setup ---------------------------------------------------------------------
    for (int i = tid; i <= tcx; i += NT)
      xs[i] = first_px_of_cell(min(cx0 + i, cx0 + ncx_t), A.sx, A.ncx, A.W);
    for (int i = tid; i <= tcy + 1; i += NT)
      rs[i] = first_px_of_cell(min(cy0 + i, cy0 + ncy_t), A.sy, A.ncy, A.H) - Y0;
    for (int r = tid; r < rows; r += NT) {
      int cy = cell_of(Y0 + r, A.sy, A.ncy);
      rowcell[r] = cy - cy0;
      rowlam[r] = lambda_of(Y0 + r, A.sy, cy);
    }
    {
      const float* src = A.logits + (size_t)b * CK * A.h * A.w;
      const int nL = CK * planeL;
      for (int i = tid; i < nL; i += NT) {
        int k = i / planeL, rem = i - k * planeL;
        int ty = rem / pitchL, tx = rem - ty * pitchL;
        int gy = min(cy0 + ty, A.h - 1), gx = min(cx0 + tx, A.w - 1);
        Ls[i] = __ldg(src + ((size_t)k * A.h + gy) * A.w + gx) * kLog2e;
      }
    }
    __syncthreads();

    // ---- phase 1: one run (row x cell) per LPR lanes -------------------------------------
    const int nitems = rows << A.tcx_log2;
    const int niter = (nitems * LPR + NT - 1) / NT;
    for (int it = 0; it < niter; ++it) {
      const int item = (it * NT + tid) / LPR;
      const bool active = item < nitems;
      const int r = item >> A.tcx_log2, cl = item & (tcx - 1);
      float Gs[CPL], G1[CPL];
      if (BWD) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { Gs[j] = 0.f; G1[j] = 0.f; }
      }
      const int xa = active ? xs[cl] : 0, xb = active ? xs[cl + 1] : 0;
      if (xb > xa) {
        const int cyl = rowcell[r];
        const float ly = rowlam[r];
        const long long rowbase = ((long long)b * A.H + (Y0 + r)) * A.W;
        // vertical lerp once per run; a = v0 - M, d = v1 - v0 (log2 domain)
        float a[CPL], d[CPL];
        float M = -INFINITY;
        {
          const float* L0 = Ls + cyl * pitchL + cl;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const int k = kbase + j;
            if (k < CK) {
              const float* Lk = L0 + k * planeL;
              float l00 = Lk[0], l01 = Lk[1], l10 = Lk[pitchL], l11 = Lk[pitchL + 1];
              float v0 = fmaf(ly, l10 - l00, l00);
              float v1 = fmaf(ly, l11 - l01, l01);
              a[j] = v0;
              d[j] = v1 - v0;
              M = fmaxf(M, fmaxf(v0, v1));
            } else {
              a[j] = -INFINITY;
              d[j] = 0.f;
            }
          }
        }
        M = group_max<LPR>(M, gmask);
#pragma unroll
        for (int j = 0; j < CPL; ++j) a[j] -= M;

        for (int X = xa; X < xb; ++X) {
          const int y = load_label<LabelT>(labels, rowbase + X);
          if (y == A.ignore || y < 0) continue;
          if (y >= C) { bad_label = true; continue; }
          if (y != cur) {
            if (BWD && cur >= 0) {
#pragma unroll
              for (int j = 0; j < CPL; ++j) {
                if (kbase + j < CK) atomicAdd(&dTs[cur * CKP + kbase + j], D[j]);
                D[j] = 0.f;
              }
            }
            cur = y;
#pragma unroll
            for (int j = 0; j < CPL; ++j) Tc[j] = Ts[y * CKP + kbase + j];
          }
          const float lam = lambda_of(X, A.sx, cx0 + cl);
          float e[CPL];
          float sum = 0.f, s = 0.f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            e[j] = ex2_approx(fmaf(lam, d[j], a[j]));
            sum += e[j];
            s = fmaf(e[j], Tc[j], s);
          }
          sum = group_sum<LPR>(sum, gmask);
          if (sum < 1e-12f) {
            // the run-level bound M was far above this pixel's true max: redo with the exact max
            float tm = -INFINITY;
#pragma unroll
            for (int j = 0; j < CPL; ++j) tm = fmaxf(tm, fmaf(lam, d[j], a[j]));
            tm = group_max<LPR>(tm, gmask);
            sum = 0.f; s = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              e[j] = ex2_approx(fmaf(lam, d[j], a[j]) - tm);
              sum += e[j];
              s = fmaf(e[j], Tc[j], s);
            }
            sum = group_sum<LPR>(sum, gmask);
          }
          s = group_sum<LPR>(s, gmask);
          const float rsum = rcp_approx(sum);
          if (MODE != MODE_BWD && sub == 0) loss_acc += lg2_approx(s * rsum);
          cnt += (sub == 0);
          if (BWD) {
            const float is = rcp_approx(s);
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              const float c = fmaf(-Tc[j], is, rsum);  // (p_k - p_k T_ky / q) / e_k
              Gs[j] = fmaf(e[j], c, Gs[j]);
              G1[j] = fmaf(lam * e[j], c, G1[j]);
              D[j] = fmaf(e[j], is, D[j]);            // p_k / q
            }
          }
        }
      }
      if (BWD) {
        // node value for column cl of this row = G0(cl) + G1(cl-1); G1 of the previous run
        // lives LPR lanes below (same warp because tcx*LPR divides 32).
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          float prev = __shfl_up_sync(0xffffffffu, G1[j], LPR);
          if (cl == 0) prev = 0.f;
          const int k = kbase + j;
          if (active && k < CK) {
            Ns[(k * A.max_rows + r) * tcx + cl] = (Gs[j] - G1[j]) + prev;
            if (cl == tcx - 1) Es[k * A.max_rows + r] = G1[j];
          }
        }
      }
    }

    if (BWD) {
      __syncthreads();
      // ---- phase 2: vertical transposed lerp, one thread per (k, node row, node col) -----
      const int nny = ncy_t + 1, nnx = ncx_t + 1;
      const int nnodes = CK * nny * nnx;
      float* dst = A.dlogits + (size_t)b * CK * A.h * A.w;
      for (int i = tid; i < nnodes; i += NT) {
        const int k = i / (nny * nnx);
        const int rem = i - k * (nny * nnx);
        const int ty = rem / nnx, tx = rem - ty * nnx;
        const float* col = (tx < tcx) ? (Ns + (size_t)k * A.max_rows * tcx + tx) : (Es + (size_t)k * A.max_rows);
        const int cstride = (tx < tcx) ? tcx : 1;
        float acc = 0.f;
        if (ty >= 1)
          for (int r = rs[ty - 1]; r < rs[ty]; ++r) acc = fmaf(rowlam[r], col[r * cstride], acc);
        if (ty < ncy_t)
          for (int r = rs[ty]; r < rs[ty + 1]; ++r) acc = fmaf(1.f - rowlam[r], col[r * cstride], acc);
        if (MODE == MODE_BWD) acc *= A.gscale;
        const int gy = min(cy0 + ty, A.h - 1), gx = min(cx0 + tx, A.w - 1);
        atomicAdd(dst + ((size_t)k * A.h + gy) * A.w + gx, acc);
      }
    }
  }

  // ---- CTA epilogue: partials ---------------------------------------------------------------
  if (BWD && cur >= 0) {
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (kbase + j < CK) atomicAdd(&dTs[cur * CKP + kbase + j], D[j]);
  }
  if (bad_label) atomicOr(A.err, SIMT_ERRBIT_LABEL_RANGE);
  double dl = (double)loss_acc;
  long long dc = cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dl += __shfl_xor_sync(0xffffffffu, dl, o);
    dc += __shfl_xor_sync(0xffffffffu, dc, o);
  }
  if (lane == 0) { red_d[tid >> 5] = dl; red_i[tid >> 5] = dc; }
  __syncthreads();
  if (tid == 0) {
    double tl = 0; long long tc = 0;
    for (int wv = 0; wv < NT / 32; ++wv) { tl += red_d[wv]; tc += red_i[wv]; }
    A.part_loss[blockIdx.x] = tl;
    A.part_cnt[blockIdx.x] = tc;
  }
  if (BWD) {
    float* pd = A.part_dT + (size_t)blockIdx.x * C * CKP;
    for (int i = tid; i < C * CKP; i += NT) pd[i] = dTs[i];
  }
}

// One warp per output: outputs 0 .. C*CKP-1 are dT entries ([y][k] layout of the partials),
// then loss and count.  Fixed summation order => deterministic.
__global__ void __launch_bounds__(256) head_finalize_kernel(
    const float* __restrict__ part_dT, const double* __restrict__ part_loss, const long long* __restrict__ part_cnt,
    int nparts, int CK, int CKP, int C, int mode, float gscale,
    double* __restrict__ stats, float* __restrict__ loss_mean, float* __restrict__ dT_out, const int* __restrict__ err) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int ndt = C * CKP;
  if (warp < ndt) {
    const int y = warp / CKP, k = warp - y * CKP;
    if (k >= CK) return;
    double s = 0;
    if (mode != MODE_FWD)
      for (int g = lane; g < nparts; g += 32) s += (double)part_dT[(size_t)g * ndt + warp];
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
// workspace: [part_loss f64 x G][part_cnt i64 x G][part_dT f32 x G*C*CKPmax], G = max grid
static constexpr int kMaxGridPerSm = 8;
static constexpr int kMaxCKP = 64;

struct Tuning { int tcy, tcx, threads, lpr; };
static Tuning g_tuning = {0, 0, 0, 0};

struct Plan {
  int CPL, LPR, NT, MINB;
  int tcy, tcx, tcx_log2, tiles_y, tiles_x, max_rows;
  long long ntiles;
  int CKP;
  size_t smem;
  int grid;
};

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int max_rows_for(int tcy, int ncy, float sy, int H) {
  int mr = 1;
  for (int cy0 = 0; cy0 < ncy; cy0 += tcy) {
    int n = (tcy < ncy - cy0) ? tcy : (ncy - cy0);
    int r = first_px_of_cell(cy0 + n, sy, ncy, H) - first_px_of_cell(cy0, sy, ncy, H);
    if (r > mr) mr = r;
  }
  return mr;
}

template <int CPL, int LPR, int MODE, typename LabelT, int NT, int MINB>
static int launch_cfg(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out, bool query_only) {
  auto kern = head_kernel<CPL, LPR, MODE, LabelT, NT, MINB>;
  SIMT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
  int occ = 0;
  SIMT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, P.smem));
  if (occ < 1) return SIMT_ENOSMEM;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  long long g = (long long)occ * di.sm_count;
  if (g > (long long)di.sm_count * kMaxGridPerSm) g = (long long)di.sm_count * kMaxGridPerSm;
  if (g > A.ntiles) g = A.ntiles;
  *grid_out = (int)g;
  if (query_only) return 0;
  prof_begin(st);
  kern<<<(int)g, NT, P.smem, st>>>(A);
  prof_end(st);
  return (int)cudaGetLastError();
}

// channel-count -> (CPL, LPR, threads, min CTAs/SM) instantiations
#define SIMT_HEAD_CONFIGS(X) \
  X(19, 1, 128, 3)           \
  X(10, 2, 256, 2)           \
  X(12, 2, 256, 2)           \
  X(17, 2, 128, 3)           \
  X(16, 4, 128, 3)

template <int MODE, typename LabelT>
static int dispatch(const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out, bool query_only) {
#define X(cpl, lpr, nt, minb) \
  if (P.CPL == cpl && P.LPR == lpr) return launch_cfg<cpl, lpr, MODE, LabelT, nt, minb>(A, P, st, grid_out, query_only);
  SIMT_HEAD_CONFIGS(X)
#undef X
  return SIMT_EUNSUPPORTED;
}

static int dispatch_all(int mode, int label_bytes, const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out,
                        bool query_only) {
  if (label_bytes == 1) {
    if (mode == MODE_FWD) return dispatch<MODE_FWD, uint8_t>(A, P, st, grid_out, query_only);
    if (mode == MODE_FWDBWD) return dispatch<MODE_FWDBWD, uint8_t>(A, P, st, grid_out, query_only);
    return dispatch<MODE_BWD, uint8_t>(A, P, st, grid_out, query_only);
  }
  if (mode == MODE_FWD) return dispatch<MODE_FWD, long long>(A, P, st, grid_out, query_only);
  if (mode == MODE_FWDBWD) return dispatch<MODE_FWDBWD, long long>(A, P, st, grid_out, query_only);
  return dispatch<MODE_BWD, long long>(A, P, st, grid_out, query_only);
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
    // prefer the fewest lanes per run, then the least channel padding
    if (!best || c.lpr < best->lpr || (c.lpr == best->lpr && c.cpl * c.lpr < best->cpl * best->lpr)) best = &c;
  }
  if (!best) return SIMT_EUNSUPPORTED;
  P->CPL = best->cpl; P->LPR = best->lpr; P->NT = best->nt; P->MINB = best->minb;
  P->CKP = best->cpl * best->lpr;
  return 0;
}

static int make_plan(int mode, int B, int CK, int C, int h, int w, int H, int W, HeadArgs* A, Plan* P) {
  int rc = choose_config(CK, g_tuning.lpr, P);
  if (rc) return rc;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc) return rc;
  A->sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  A->sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  A->ncy = h > 1 ? h - 1 : 1;
  A->ncx = w > 1 ? w - 1 : 1;
  const int max_tcx = 32 / P->LPR;
  int tcx = g_tuning.tcx > 0 ? g_tuning.tcx : 8;
  int tcy = g_tuning.tcy > 0 ? g_tuning.tcy : 8;
  if (tcx > max_tcx) tcx = max_tcx;
  tcx = 1 << ilog2(tcx);
  if (tcx > max_tcx) tcx = max_tcx;
  const bool bwd = mode != MODE_FWD;
  auto ntiles_of = [&](int ty, int tx) {
    return (long long)B * ((A->ncy + ty - 1) / ty) * ((A->ncx + tx - 1) / tx);
  };
  if (g_tuning.tcx <= 0 && g_tuning.tcy <= 0) {
    // enough tiles to balance a persistent grid of ~3 CTAs/SM: shrink small problems' tiles
    const long long want = 4LL * 3 * di.sm_count;
    while (ntiles_of(tcy, tcx) < want && (tcy > 2 || tcx > 4)) {
      if (tcy >= tcx && tcy > 2) tcy >>= 1; else if (tcx > 4) tcx >>= 1; else tcy >>= 1;
    }
  }
  // shrink until the tile fits in shared memory
  for (;;) {
    int mr = max_rows_for(tcy, A->ncy, A->sy, H);
    SmemLayout SL = smem_layout(CK, P->CKP, C, tcy, tcx, mr, bwd);
    const size_t budget = ((size_t)228 * 1024 - (size_t)P->MINB * 1024) / P->MINB;  // MINB CTAs per SM
    if (SL.total <= budget || (tcy == 1 && tcx == 1)) {
      if (SL.total > (size_t)di.smem_optin) return SIMT_ENOSMEM;
      P->max_rows = mr;
      P->smem = SL.total;
      break;
    }
    if (tcy > 1) tcy >>= 1; else tcx >>= 1;
  }
  P->tcy = tcy; P->tcx = tcx; P->tcx_log2 = ilog2(tcx);
  P->tiles_y = (A->ncy + tcy - 1) / tcy;
  P->tiles_x = (A->ncx + tcx - 1) / tcx;
  P->ntiles = (long long)B * P->tiles_y * P->tiles_x;
  A->tcy = tcy; A->tcx = tcx; A->tcx_log2 = P->tcx_log2;
  A->tiles_y = P->tiles_y; A->tiles_x = P->tiles_x; A->ntiles = P->ntiles;
  A->max_rows = P->max_rows;
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
  A.part_loss = reinterpret_cast<double*>(ws);
  A.part_cnt = reinterpret_cast<long long*>(ws + G * 8);
  A.part_dT = reinterpret_cast<float*>(ws + G * 16);
  if (mode != MODE_FWD)
    SIMT_CUDA_TRY(cudaMemsetAsync(dlogits, 0, (size_t)B * CK * h * w * sizeof(float), st));
  int grid = 0;
  rc = dispatch_all(mode, label_bytes, A, P, st, &grid, false);
  if (rc) return rc;
  const int nwarps = C * P.CKP + 1;
  const int fgrid = (nwarps * 32 + 255) / 256;
  head_finalize_kernel<<<fgrid, 256, 0, st>>>(A.part_dT, A.part_loss, A.part_cnt, grid, CK, P.CKP, C, mode, gscale,
                                              stats, loss_mean, dT_out, err_flag);
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
  return G * 16 + G * (size_t)(C > 0 ? C : 1) * kMaxCKP * sizeof(float);
}

void simt_head_set_tuning(int tcy, int tcx, int threads, int lpr) { g_tuning = {tcy, tcx, threads, lpr}; }

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
