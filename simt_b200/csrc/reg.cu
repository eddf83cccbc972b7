// T regularisers and anchor statistics of the SimT head (sm_100a).
//
// simt_t_regularizers: convex + volume terms of one head and their gradients in ONE single-CTA
// launch.  Replaces, per head, tools/trainV2_simt.py:412-415 (W.mm(T), MSELoss(sum)), :417-421
// (T^T T, torch.linalg.det -> cuSOLVER LU, log sqrt abs, and the host sync at torch.isinf) and the
// autograd backward of those lines:  ~11 tiny launches + a device->host sync become one launch.
//   convex = -||W T||_F^2            d/dT = -2 W^T (W T)      d/dW = -2 (W T) T^T
//   volume = 0.5 log|det(T^T T)|     d/dT = T (T^T T)^-1      (0 and zero gradient if not finite)
// G = T^T T is symmetric positive (semi-)definite, so |det| = det and a Cholesky factorisation
// in fp64 gives log det = 2 sum log L_ii and G^-1 by two triangular solves.
//
// simt_anchor_stats: Anchor_index / Exist_label of trainV2_simt.py:375-377 computed from the
// LOW-res logits (the [N, CK] upsampled tensor is never materialised): per channel the arg-max
// pixel of the bilinearly upsampled logit, and the set of classes that are the per-pixel arg-max
// somewhere.  Ties: the smallest pixel index / smallest class wins (torch's CPU argmax).
#include "common.cuh"

namespace simt {

static constexpr int kMaxC = 64;  // C and CK bound for the single-CTA regulariser kernel

__global__ void __launch_bounds__(256) t_reg_kernel(const float* __restrict__ T, const float* __restrict__ W, int CK,
                                                     int C, float* __restrict__ out2, float* __restrict__ dT_convex,
                                                     float* __restrict__ dT_volume, float* __restrict__ dW_convex) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // doubles first (alignment), then floats
  double* G = reinterpret_cast<double*>(smem_raw);          // [C][C]  T^T T, then its Cholesky factor L
  double* X = G + C * C;                                     // [C][CK] G^-1 T^T
  float* Ts = reinterpret_cast<float*>(X + C * CK);          // [CK][C]
  float* Ws = Ts + CK * C;                                   // [CK][CK]
  float* P = Ws + CK * CK;                                   // [CK][C]  W T
  __shared__ double red[8];
  __shared__ int spd_ok;
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int i = tid; i < CK * C; i += nt) Ts[i] = T[i];
  if (W)
    for (int i = tid; i < CK * CK; i += nt) Ws[i] = W[i];
  if (tid == 0) spd_ok = 1;
  __syncthreads();

  // ---- convex: P = W T, loss = -sum P^2 --------------------------------------------------------
  double part = 0;
  if (W) {
    for (int i = tid; i < CK * C; i += nt) {
      const int r = i / C, c = i - r * C;
      float acc = 0.f;
      for (int k = 0; k < CK; ++k) acc = fmaf(Ws[r * CK + k], Ts[k * C + c], acc);
      P[i] = acc;
      part += (double)acc * (double)acc;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  // ---- G = T^T T (fp64 accumulate) ---------------------------------------------------------------
  for (int i = tid; i < C * C; i += nt) {
    const int a = i / C, b = i - a * C;
    double acc = 0;
    for (int k = 0; k < CK; ++k) acc += (double)Ts[k * C + a] * (double)Ts[k * C + b];
    G[i] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int k = 0; k < nt / 32; ++k) s += red[k];
    out2[0] = W ? (float)(-s) : 0.f;
  }
  if (W) {
    // dT_convex = -2 W^T P ; dW_convex = -2 P T^T
    for (int i = tid; i < CK * C; i += nt) {
      const int r = i / C, c = i - r * C;
      float acc = 0.f;
      for (int k = 0; k < CK; ++k) acc = fmaf(Ws[k * CK + r], P[k * C + c], acc);
      dT_convex[i] = -2.f * acc;
    }
    for (int i = tid; i < CK * CK; i += nt) {
      const int r = i / CK, c = i - r * CK;
      float acc = 0.f;
      for (int k = 0; k < C; ++k) acc = fmaf(P[r * C + k], Ts[c * C + k], acc);
      dW_convex[i] = -2.f * acc;
    }
  }

  // ---- Cholesky G = L L^T in place (lower), column by column -----------------------------------------
  for (int j = 0; j < C; ++j) {
    __syncthreads();
    if (tid == 0) {
      double d = G[j * C + j];
      for (int k = 0; k < j; ++k) d -= G[j * C + k] * G[j * C + k];
      if (!(d > 0.0) || !isfinite(d)) { spd_ok = 0; d = 1.0; }
      G[j * C + j] = sqrt(d);
    }
    __syncthreads();
    const double ljj = G[j * C + j];
    for (int i = j + 1 + tid; i < C; i += nt) {
      double v = G[i * C + j];
      for (int k = 0; k < j; ++k) v -= G[i * C + k] * G[j * C + k];
      G[i * C + j] = v / ljj;
    }
  }
  __syncthreads();
  double logdet = 0;
  if (tid == 0) {
    for (int j = 0; j < C; ++j) logdet += log(G[j * C + j]);
    const double vol = logdet;  // 0.5 * log det(G) = 0.5 * 2 * sum log L_jj
    const bool fin = spd_ok && isfinite(vol);
    if (!fin) spd_ok = 0;
    out2[1] = fin ? (float)vol : 0.f;  // trainV2_simt.py:420-421
  }
  __syncthreads();
  // ---- X = G^-1 T^T: one thread per right-hand side (a row of T) -----------------------------------------
  for (int r = tid; r < CK; r += nt) {
    // forward: L y = t ; backward: L^T x = y ; X[:, r]
    for (int i = 0; i < C; ++i) {
      double v = (double)Ts[r * C + i];
      for (int k = 0; k < i; ++k) v -= G[i * C + k] * X[k * CK + r];
      X[i * CK + r] = v / G[i * C + i];
    }
    for (int i = C - 1; i >= 0; --i) {
      double v = X[i * CK + r];
      for (int k = i + 1; k < C; ++k) v -= G[k * C + i] * X[k * CK + r];
      X[i * CK + r] = v / G[i * C + i];
    }
  }
  __syncthreads();
  for (int i = tid; i < CK * C; i += nt) {
    const int r = i / C, c = i - r * C;
    dT_volume[i] = spd_ok ? (float)X[c * CK + r] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------
// anchor statistics
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_max(float v, long long idx) {
  // monotone float -> uint in the high 32 bits; low 32 bits = ~idx so that the SMALLEST index wins ties
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);
}

struct Bilin {
  int o00, o01, o10, o11;
  float ly0, ly1, lx0, lx1;
};
// torch's align_corners=True source index / lambda arithmetic (UpSample.h), unfused
__device__ __forceinline__ Bilin bilin_setup(int Y, int X, int h, int w, float sy, float sx) {
  const float fy = __fmul_rn(sy, (float)Y), fx = __fmul_rn(sx, (float)X);
  const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
  const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
  Bilin b;
  b.ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f);
  b.lx1 = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
  b.ly0 = 1.f - b.ly1;
  b.lx0 = 1.f - b.lx1;
  b.o00 = y0 * w + x0; b.o01 = y0 * w + x1; b.o10 = y1 * w + x0; b.o11 = y1 * w + x1;
  return b;
}
__device__ __forceinline__ float bilin_eval(const float* __restrict__ pl, const Bilin& b) {
  // w0h*(w0w*x00 + w1w*x01) + w1h*(w0w*x10 + w1w*x11), every product and sum rounded
  const float top = __fadd_rn(__fmul_rn(b.lx0, __ldg(pl + b.o00)), __fmul_rn(b.lx1, __ldg(pl + b.o01)));
  const float bot = __fadd_rn(__fmul_rn(b.lx0, __ldg(pl + b.o10)), __fmul_rn(b.lx1, __ldg(pl + b.o11)));
  return __fadd_rn(__fmul_rn(b.ly0, top), __fmul_rn(b.ly1, bot));
}

// (A) per channel: arg-max pixel of the upsampled logit.  grid = (blocks, CK)
__global__ void __launch_bounds__(256) anchor_channel_max_kernel(const float* __restrict__ logits, int B, int CK, int h,
                                                                  int w, int H, int W, float sy, float sx,
                                                                  unsigned long long* __restrict__ packed) {
  const int k = blockIdx.y;
  const long long npix = (long long)B * H * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long best = 0ull;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
    const int X = (int)(p % W);
    const long long t = p / W;
    const int Y = (int)(t % H);
    const int b = (int)(t / H);
    const Bilin bl = bilin_setup(Y, X, h, w, sy, sx);
    const float z = bilin_eval(logits + ((size_t)b * CK + k) * h * w, bl);
    const unsigned long long pk = pack_max(z, p);
    best = pk > best ? pk : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  __shared__ unsigned long long sb[8];
  if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) best = sb[i] > best ? sb[i] : best;
    if (best) atomicMax(&packed[k], best);
  }
}

// (B) per pixel: arg-max class (lowest class wins ties) -> presence bit mask
__global__ void __launch_bounds__(256) anchor_exist_kernel(const float* __restrict__ logits, int B, int CK, int h, int w,
                                                            int H, int W, float sy, float sx,
                                                            unsigned long long* __restrict__ exist) {
  const long long npix = (long long)B * H * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long mine = 0ull;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
    const int X = (int)(p % W);
    const long long t = p / W;
    const int Y = (int)(t % H);
    const int b = (int)(t / H);
    const Bilin bl = bilin_setup(Y, X, h, w, sy, sx);
    const float* base = logits + (size_t)b * CK * h * w;
    float best = -INFINITY;
    int bestk = 0;
    for (int k = 0; k < CK; ++k) {
      const float z = bilin_eval(base + (size_t)k * h * w, bl);
      if (z > best || k == 0) { best = z; bestk = k; }
    }
    mine |= 1ull << bestk;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine |= __shfl_xor_sync(0xffffffffu, mine, o);
  __shared__ unsigned long long se;
  if (threadIdx.x == 0) se = 0ull;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && mine) atomicOr(&se, mine);
  __syncthreads();
  if (threadIdx.x == 0 && se) atomicOr(exist, se);
}

__global__ void anchor_unpack_kernel(const unsigned long long* __restrict__ packed, int CK,
                                     long long* __restrict__ anchor_idx, float* __restrict__ anchor_val) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= CK) return;
  const unsigned long long pk = packed[k];
  anchor_idx[k] = (long long)(0xffffffffu - (unsigned)(pk & 0xffffffffull));
  unsigned u = (unsigned)(pk >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  if (anchor_val) anchor_val[k] = __uint_as_float(u);
}

// rows[r][c] = bilinear sample of src[b, c] at flat pixel idx[r]  (labelC_flat[Anchor_index], trainV2_simt.py:378)
__global__ void bilinear_gather_kernel(const float* __restrict__ src, int B, int C, int h, int w, int H, int W, float sy,
                                       float sx, const long long* __restrict__ idx, int n, float* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  const int r = i / C, c = i - r * C;
  const long long p = idx[r];
  const int X = (int)(p % W);
  const long long t = p / W;
  const int Y = (int)(t % H);
  const int b = (int)(t / H);
  const Bilin bl = bilin_setup(Y, X, h, w, sy, sx);
  rows[i] = bilin_eval(src + ((size_t)b * C + c) * h * w, bl);
}

}  // namespace simt

using namespace simt;

extern "C" {

int simt_t_regularizers(const float* T, const float* W, int CK, int C, float* out2, float* dT_convex,
                        float* dT_volume, float* dW_convex, void* stream) {
  if (!T || !out2 || !dT_volume || CK <= 0 || C <= 0) return SIMT_EINVAL;
  if (W && (!dT_convex || !dW_convex)) return SIMT_EINVAL;
  if (CK > kMaxC || C > kMaxC || C > CK) return SIMT_EUNSUPPORTED;
  const size_t smem = (size_t)(C * C + C * CK) * 8 + (size_t)(2 * CK * C + CK * CK) * 4;
  static bool attr_done[64] = {};
  {
    const int rc = ensure_dynamic_smem(t_reg_kernel, attr_done, 160 * 1024);
    if (rc) return rc;
  }
  t_reg_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(T, W, CK, C, out2, dT_convex, dT_volume, dW_convex);
  return (int)cudaGetLastError();
}

int simt_anchor_stats(const float* logits, int B, int CK, int h, int w, int H, int W, long long* anchor_idx,
                      float* anchor_val, unsigned long long* exist_mask, unsigned long long* scratch, void* stream) {
  if (!logits || !anchor_idx || !exist_mask || !scratch) return SIMT_EINVAL;
  if (B <= 0 || CK <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return SIMT_EINVAL;
  if (CK > 64) return SIMT_EUNSUPPORTED;
  if ((long long)B * H * W > 0xffffffffLL) return SIMT_EUNSUPPORTED;  // pixel index is packed in 32 bits
  cudaStream_t st = (cudaStream_t)stream;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  SIMT_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * CK, st));
  SIMT_CUDA_TRY(cudaMemsetAsync(exist_mask, 0, sizeof(unsigned long long), st));
  const float sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const long long npix = (long long)B * H * W;
  long long gx = (npix + 256 * 8 - 1) / (256 * 8);
  if (gx > (long long)di.sm_count * 2) gx = (long long)di.sm_count * 2;
  if (gx < 1) gx = 1;
  anchor_channel_max_kernel<<<dim3((unsigned)gx, (unsigned)CK), 256, 0, st>>>(logits, B, CK, h, w, H, W, sy, sx, scratch);
  SIMT_CUDA_TRY(cudaGetLastError());
  long long ge = (npix + 255) / 256;
  if (ge > (long long)di.sm_count * 8) ge = (long long)di.sm_count * 8;
  anchor_exist_kernel<<<(int)ge, 256, 0, st>>>(logits, B, CK, h, w, H, W, sy, sx, exist_mask);
  SIMT_CUDA_TRY(cudaGetLastError());
  anchor_unpack_kernel<<<1, 64, 0, st>>>(scratch, CK, anchor_idx, anchor_val);
  return (int)cudaGetLastError();
}

int simt_bilinear_gather(const float* src, int B, int C, int h, int w, int H, int W, const long long* pixel_idx, int n,
                         float* rows, void* stream) {
  if (!src || !pixel_idx || !rows || B <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || n < 0) return SIMT_EINVAL;
  if (n == 0) return 0;
  const float sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  bilinear_gather_kernel<<<(n * C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src, B, C, h, w, H, W, sy, sx, pixel_idx, n,
                                                                                rows);
  return (int)cudaGetLastError();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// (f) row 2: pseudo-label generation, tools/trainV2_simt.py:354-365 + class-posterior relabel :387-393
// ---------------------------------------------------------------------------------------------------
namespace simt {

// channel softmax of the frozen model's LOW-res logits (the reference applies softmax before the upsample)
__global__ void __launch_bounds__(256) softmax_lo_kernel(const float* __restrict__ x, int B, int C, int hw,
                                                          float* __restrict__ p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * hw) return;
  const int b = (int)(i / hw), r = (int)(i - (long long)b * hw);
  const float* xb = x + (size_t)b * C * hw + r;
  float m = -INFINITY;
  for (int k = 0; k < C; ++k) m = fmaxf(m, xb[(size_t)k * hw]);
  float s = 0.f;
  for (int k = 0; k < C; ++k) s += expf(xb[(size_t)k * hw] - m);
  float* pb = p + (size_t)b * C * hw + r;
  for (int k = 0; k < C; ++k) pb[(size_t)k * hw] = expf(xb[(size_t)k * hw] - m) / s;
}

// One thread per (image, pixel row, low-res column x0): the run of output pixels whose left neighbour node is x0
// (8 pixels at the training shape).  The four corner values of a channel are loaded ONCE per run and every pixel of
// the run is evaluated with torch's unfused formula (same bits as a per-pixel gather, 1/8 of the loads):
// upsampled class posterior -> max / arg-max -> thresholds -> (student arg-max) -> uint8.
static constexpr int kRunPx = 8;

__global__ void __launch_bounds__(256) pseudo_label_kernel(const float* __restrict__ probs_lo, const float* __restrict__ pred2_lo,
                                                            int B, int C, int CK, int h, int w, int H, int W, float sy,
                                                            float sx, float thr_hi, float thr_lo,
                                                            uint8_t* __restrict__ out) {
  const long long nthreads = (long long)B * H * w;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nthreads; t += stride) {
    const int x0 = (int)(t % w);
    const long long r = t / w;
    const int Y = (int)(r % H);
    const int b = (int)(r / H);
    // pixels with min(floor(sx * X), w - 1) == x0
    const int xa = first_px_of_cell(x0, sx, w, W);
    const int xe = (x0 + 1 < w) ? first_px_of_cell(x0 + 1, sx, w, W) : W;
    if (xa >= xe) continue;
    const float fy = __fmul_rn(sy, (float)Y);
    const int y0 = min((int)fy, h - 1), y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
    const float ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), ly0 = 1.f - ly1;
    const int o00 = y0 * w + x0, o01 = y0 * w + x1, o10 = y1 * w + x0, o11 = y1 * w + x1;
    const float* base = probs_lo + (size_t)b * C * h * w;
    const float* pb = pred2_lo ? pred2_lo + (size_t)b * CK * h * w : nullptr;
    uint8_t* orow = out + ((size_t)b * H + Y) * W;
    for (int xs = xa; xs < xe; xs += kRunPx) {
      const int n = min(kRunPx, xe - xs);
      float lx0[kRunPx], lx1[kRunPx], best[kRunPx];
      int bestk[kRunPx];
#pragma unroll
      for (int j = 0; j < kRunPx; ++j) {
        const float fx = __fmul_rn(sx, (float)(xs + j));
        lx1[j] = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
        lx0[j] = 1.f - lx1[j];
        best[j] = -INFINITY;
        bestk[j] = 0;
      }
      for (int k = 0; k < C; ++k) {
        const float* pl = base + (size_t)k * h * w;
        const float v00 = __ldg(pl + o00), v01 = __ldg(pl + o01), v10 = __ldg(pl + o10), v11 = __ldg(pl + o11);
#pragma unroll
        for (int j = 0; j < kRunPx; ++j) {
          // w0h*(w0w*x00 + w1w*x01) + w1h*(w0w*x10 + w1w*x11), every product and sum rounded
          const float top = __fadd_rn(__fmul_rn(lx0[j], v00), __fmul_rn(lx1[j], v01));
          const float bot = __fadd_rn(__fmul_rn(lx0[j], v10), __fmul_rn(lx1[j], v11));
          const float z = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
          if (z > best[j] || k == 0) { best[j] = z; bestk[j] = k; }
        }
      }
      int lab[kRunPx];
      bool low = false;
#pragma unroll
      for (int j = 0; j < kRunPx; ++j) {
        lab[j] = (best[j] > thr_hi) ? bestk[j] : 255;        // :359
        if (best[j] < thr_lo) {                              // :361 -> low confidence: :387-393
          lab[j] = pb ? 255 : C;                             // without the student: the raw marker of :361
          low = low || (j < n);
        }
      }
      if (low && pb) {
        float sb[kRunPx];
        int sk[kRunPx];
#pragma unroll
        for (int j = 0; j < kRunPx; ++j) { sb[j] = -INFINITY; sk[j] = 0; }
        for (int k = 0; k < CK; ++k) {
          const float* pl = pb + (size_t)k * h * w;
          const float v00 = __ldg(pl + o00), v01 = __ldg(pl + o01), v10 = __ldg(pl + o10), v11 = __ldg(pl + o11);
#pragma unroll
          for (int j = 0; j < kRunPx; ++j) {
            const float top = __fadd_rn(__fmul_rn(lx0[j], v00), __fmul_rn(lx1[j], v01));
            const float bot = __fadd_rn(__fmul_rn(lx0[j], v10), __fmul_rn(lx1[j], v11));
            const float z = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
            if (z > sb[j] || k == 0) { sb[j] = z; sk[j] = k; }
          }
        }
#pragma unroll
        for (int j = 0; j < kRunPx; ++j)
          if (best[j] < thr_lo && sk[j] >= C) lab[j] = sk[j];   // an open-set placeholder class
      }
#pragma unroll
      for (int j = 0; j < kRunPx; ++j)
        if (j < n) orow[xs + j] = (uint8_t)lab[j];
    }
  }
}

}  // namespace simt

extern "C" int simt_pseudo_labels(const float* fixed_logits_lo, const float* pred2_lo, int B, int C, int CK, int h, int w,
                                  int H, int W, float thres_high, float thres_low, float* probs_scratch,
                                  uint8_t* labels_out, void* stream) {
  using namespace simt;
  if (!fixed_logits_lo || !probs_scratch || !labels_out) return SIMT_EINVAL;
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || C > 254) return SIMT_EINVAL;
  if (pred2_lo && (CK < C || CK > 254)) return SIMT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  const long long nlo = (long long)B * h * w;
  softmax_lo_kernel<<<(unsigned)((nlo + 255) / 256), 256, 0, st>>>(fixed_logits_lo, B, C, h * w, probs_scratch);
  SIMT_CUDA_TRY(cudaGetLastError());
  const float sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  long long grid = ((long long)B * H * w + 255) / 256;          // one thread per (image, pixel row, low-res column)
  if (grid > (long long)di.sm_count * 16) grid = (long long)di.sm_count * 16;
  prof_begin(st);
  pseudo_label_kernel<<<(int)grid, 256, 0, st>>>(probs_scratch, pred2_lo, B, C, CK, h, w, H, W, sy, sx, thres_high,
                                                 thres_low, labels_out);
  prof_end(st);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// (f) row 3: fused eval prediction map, tools/evaluate_cityscapes.py:127-138 (two-scale) / :186-196 (one scale)
// ---------------------------------------------------------------------------------------------------
namespace simt {

// One thread per (image, pixel row, column x0 of the FIRST scale): a run of output pixels (8 at the evaluation shape)
// shares the four corners of scale a; of scale b it needs at most three neighbouring columns when b's cells are not
// narrower than the run (the two-scale case of evaluate_cityscapes.py: 129x257 and 81x161 -> 1024x2048), otherwise b
// falls back to one gather per pixel.  Every pixel is evaluated with torch's unfused formula on the same corner
// values as a per-pixel gather (identical bits, ~1/6 of the loads).
__global__ void __launch_bounds__(256) eval_argmax_kernel(const float* __restrict__ la, int CKa, int ha, int wa,
                                                           const float* __restrict__ lb, int CKb, int hb, int wb, int B,
                                                           int C, int H, int W, float sya, float sxa, float syb,
                                                           float sxb, uint8_t* __restrict__ pred) {
  const long long nthreads = (long long)B * H * wa;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nthreads; t += stride) {
    const int x0 = (int)(t % wa);
    const long long r = t / wa;
    const int Y = (int)(r % H);
    const int b = (int)(r / H);
    const int xa = first_px_of_cell(x0, sxa, wa, W);
    const int xe = (x0 + 1 < wa) ? first_px_of_cell(x0 + 1, sxa, wa, W) : W;
    if (xa >= xe) continue;
    // scale a: rows and the run's two columns
    const float fya = __fmul_rn(sya, (float)Y);
    const int ya0 = min((int)fya, ha - 1), ya1 = ya0 + (ya0 < ha - 1), x1 = x0 + (x0 < wa - 1);
    const float lya1 = fminf(fmaxf(fya - (float)ya0, 0.f), 1.f), lya0 = 1.f - lya1;
    const int a00 = ya0 * wa + x0, a01 = ya0 * wa + x1, a10 = ya1 * wa + x0, a11 = ya1 * wa + x1;
    const float* pa = la + (size_t)b * CKa * ha * wa;
    // scale b: rows
    const float* pb = lb ? lb + (size_t)b * CKb * hb * wb : nullptr;
    int yb0 = 0, yb1 = 0;
    float lyb1 = 0.f, lyb0 = 1.f;
    if (pb) {
      const float fyb = __fmul_rn(syb, (float)Y);
      yb0 = min((int)fyb, hb - 1); yb1 = yb0 + (yb0 < hb - 1);
      lyb1 = fminf(fmaxf(fyb - (float)yb0, 0.f), 1.f); lyb0 = 1.f - lyb1;
    }
    uint8_t* orow = pred + ((size_t)b * H + Y) * W;
    for (int xs = xa; xs < xe; xs += kRunPx) {
      const int n = min(kRunPx, xe - xs);
      float lx0[kRunPx], lx1[kRunPx], bx0[kRunPx], bx1[kRunPx], best[kRunPx];
      int bestk[kRunPx], ob[kRunPx];
      int c0 = 0;
      bool near3 = true;   // b's columns of this chunk are within {c0, c0 + 1}
#pragma unroll
      for (int j = 0; j < kRunPx; ++j) {
        const float fx = __fmul_rn(sxa, (float)(xs + j));
        lx1[j] = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
        lx0[j] = 1.f - lx1[j];
        best[j] = -INFINITY;
        bestk[j] = 0;
        bx0[j] = 1.f; bx1[j] = 0.f; ob[j] = 0;
        if (pb) {
          const float fxb = __fmul_rn(sxb, (float)(xs + j));
          const int xb = min((int)fxb, wb - 1);
          if (j == 0) c0 = xb;
          bx1[j] = fminf(fmaxf(fxb - (float)xb, 0.f), 1.f);
          bx0[j] = 1.f - bx1[j];
          ob[j] = xb - c0;
          if (j < n && ob[j] > 1) near3 = false;
        }
      }
      const int cb0 = min(c0, wb - 1), cb1 = min(c0 + 1, wb - 1), cb2 = min(c0 + 2, wb - 1);
      for (int k = 0; k < C; ++k) {
        const float* pl = pa + (size_t)k * ha * wa;
        const float v00 = __ldg(pl + a00), v01 = __ldg(pl + a01), v10 = __ldg(pl + a10), v11 = __ldg(pl + a11);
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
        const float* ql = pb ? pb + (size_t)k * hb * wb : nullptr;
        if (pb && near3) {
          t0 = __ldg(ql + yb0 * wb + cb0); t1 = __ldg(ql + yb0 * wb + cb1); t2 = __ldg(ql + yb0 * wb + cb2);
          u0 = __ldg(ql + yb1 * wb + cb0); u1 = __ldg(ql + yb1 * wb + cb1); u2 = __ldg(ql + yb1 * wb + cb2);
        }
#pragma unroll
        for (int j = 0; j < kRunPx; ++j) {
          // w0h*(w0w*x00 + w1w*x01) + w1h*(w0w*x10 + w1w*x11), every product and sum rounded
          const float top = __fadd_rn(__fmul_rn(lx0[j], v00), __fmul_rn(lx1[j], v01));
          const float bot = __fadd_rn(__fmul_rn(lx0[j], v10), __fmul_rn(lx1[j], v11));
          float z = __fadd_rn(__fmul_rn(lya0, top), __fmul_rn(lya1, bot));
          if (pb) {
            float l0, r0, l1, r1;
            if (near3) {
              const bool o = ob[j] != 0;
              l0 = o ? t1 : t0; r0 = o ? t2 : t1;
              l1 = o ? u1 : u0; r1 = o ? u2 : u1;
            } else {   // b's cells are narrower than the run: plain gather for this pixel
              const int xb = min(c0 + ob[j], wb - 1), xb1 = xb + (xb < wb - 1);
              l0 = __ldg(ql + yb0 * wb + xb); r0 = __ldg(ql + yb0 * wb + xb1);
              l1 = __ldg(ql + yb1 * wb + xb); r1 = __ldg(ql + yb1 * wb + xb1);
            }
            const float tb = __fadd_rn(__fmul_rn(bx0[j], l0), __fmul_rn(bx1[j], r0));
            const float bb = __fadd_rn(__fmul_rn(bx0[j], l1), __fmul_rn(bx1[j], r1));
            z = __fadd_rn(z, __fadd_rn(__fmul_rn(lyb0, tb), __fmul_rn(lyb1, bb)));   // the reference adds in fp32 on the host
          }
          if (z > best[j] || k == 0) { best[j] = z; bestk[j] = k; }                    // first maximum, like np.argmax
        }
      }
#pragma unroll
      for (int j = 0; j < kRunPx; ++j)
        if (j < n) orow[xs + j] = (uint8_t)bestk[j];
    }
  }
}

}  // namespace simt

extern "C" int simt_eval_argmax(const float* logits_a, int CKa, int ha, int wa, const float* logits_b, int CKb, int hb,
                                int wb, int B, int C, int H, int W, uint8_t* pred_out, void* stream) {
  using namespace simt;
  if (!logits_a || !pred_out || B <= 0 || C <= 0 || C > 255 || CKa < C || ha <= 0 || wa <= 0 || H <= 0 || W <= 0)
    return SIMT_EINVAL;
  if (logits_b && (CKb < C || hb <= 0 || wb <= 0)) return SIMT_EINVAL;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  auto sc = [](int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; };
  long long grid = ((long long)B * H * wa + 255) / 256;       // one thread per (image, pixel row, column of scale a)
  if (grid > (long long)di.sm_count * 16) grid = (long long)di.sm_count * 16;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  eval_argmax_kernel<<<(int)grid, 256, 0, st>>>(logits_a, CKa, ha, wa, logits_b, CKb, hb, wb, B, C, H, W, sc(ha, H),
                                                sc(wa, W), logits_b ? sc(hb, H) : 0.f, logits_b ? sc(wb, W) : 0.f,
                                                pred_out);
  prof_end(st);
  return (int)cudaGetLastError();
}
