// The inner W optimisation of tools/trainV2_simt.py:326-339 as ONE single-CTA launch per head.
//
// The reference runs, every training iteration and for each of its two heads, 10 rounds of
//   W = sig_W()                      (model/deeplab_multi.py:277-286: diag := -1e4, row softmax, minus I)
//   loss = MSELoss(sum)(W.mm(T), 0)  (:336)
//   loss.backward(retain_graph=True) (:337; this also ACCUMULATES into NTM.grad, which :317-318 zeroed once)
//   optimizer_w.step()               (:338-339, torch.optim.Adam, weight_decay 0)
// which is ~20 eager launches per round and head.  Everything is a few hundred floats, so the whole loop lives in the
// shared memory of one CTA: the Adam moments and the weight are read once and written once.
#include "common.cuh"

namespace simt {

static constexpr int kMaxW = 64;  // CK bound (same as the regulariser kernel)

__global__ void __launch_bounds__(256) w_fit_kernel(float* __restrict__ weight, float* __restrict__ exp_avg,
                                                     float* __restrict__ exp_avg_sq, const float* __restrict__ T, int n,
                                                     int C, int n_steps, long long step0, double lr, double beta1,
                                                     double beta2, double eps, float* __restrict__ dT_accum,
                                                     float* __restrict__ losses) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm;            // [n][n] sig_W.weight
  float* M = A + n * n;     // [n][n] exp_avg
  float* V = M + n * n;     // [n][n] exp_avg_sq
  float* Wm = V + n * n;    // [n][n] softmax (then softmax - I; the diagonal of the softmax is exactly 0)
  float* D = Wm + n * n;    // [n][n] dLoss/dW
  float* Ts = D + n * n;    // [n][C]
  float* R = Ts + n * C;    // [n][C] W T
  float* dTs = R + n * C;   // [n][C] accumulated dLoss/dT over the steps
  __shared__ double red[8];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;

  for (int i = tid; i < n * n; i += nt) {
    A[i] = weight[i];
    M[i] = exp_avg[i];
    V[i] = exp_avg_sq[i];
  }
  for (int i = tid; i < n * C; i += nt) {
    Ts[i] = T[i];
    dTs[i] = 0.f;
  }
  __syncthreads();

  for (int it = 0; it < n_steps; ++it) {
    // ---- sig_W.forward: diag := -1e4, softmax over dim 1, minus identity (one warp per row) ------
    for (int r = warp; r < n; r += nw) {
      const int c0 = lane, c1 = lane + 32;
      if (c0 == r) A[r * n + c0] = -10000.f;
      if (c1 == r) A[r * n + c1] = -10000.f;
      __syncwarp();
      const float a0 = c0 < n ? A[r * n + c0] : -INFINITY;
      const float a1 = c1 < n ? A[r * n + c1] : -INFINITY;
      float mx = fmaxf(a0, a1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e0 = c0 < n ? expf(a0 - mx) : 0.f;
      const float e1 = c1 < n ? expf(a1 - mx) : 0.f;
      float s = e0 + e1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (c0 < n) Wm[r * n + c0] = e0 / s - (c0 == r ? 1.f : 0.f);
      if (c1 < n) Wm[r * n + c1] = e1 / s - (c1 == r ? 1.f : 0.f);
    }
    __syncthreads();
    // ---- R = W T, loss = sum R^2 -------------------------------------------------------------------
    double part = 0;
    for (int i = tid; i < n * C; i += nt) {
      const int r = i / C, c = i - r * C;
      float acc = 0.f;
      for (int k = 0; k < n; ++k) acc = fmaf(Wm[r * n + k], Ts[k * C + c], acc);
      R[i] = acc;
      part += (double)acc * (double)acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0 && losses) {
      double s = 0;
      for (int k = 0; k < nw; ++k) s += red[k];
      losses[it] = (float)s;
    }
    // ---- backward of the sum-MSE: dW = 2 R T^T, dT += 2 W^T R ------------------------------------
    for (int i = tid; i < n * n; i += nt) {
      const int r = i / n, c = i - r * n;
      float acc = 0.f;
      for (int k = 0; k < C; ++k) acc = fmaf(R[r * C + k], Ts[c * C + k], acc);
      D[i] = 2.f * acc;
    }
    for (int i = tid; i < n * C; i += nt) {
      const int r = i / C, c = i - r * C;
      float acc = 0.f;
      for (int k = 0; k < n; ++k) acc = fmaf(Wm[k * n + r], R[k * C + c], acc);
      dTs[i] += 2.f * acc;
    }
    __syncthreads();
    // ---- softmax backward + Adam (one warp per row) ----------------------------------------------
    const double t = (double)(step0 + it + 1);
    const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
    const float step_size = (float)(lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
    const float epsf = (float)eps;
    for (int r = warp; r < n; r += nw) {
      const int c0 = lane, c1 = lane + 32;
      // softmax output y = Wm + I (the diagonal entry is exactly 0: exp(-1e4 - max) underflows)
      const float y0 = (c0 < n && c0 != r) ? Wm[r * n + c0] : 0.f;
      const float y1 = (c1 < n && c1 != r) ? Wm[r * n + c1] : 0.f;
      const float d0 = c0 < n ? D[r * n + c0] : 0.f;
      const float d1 = c1 < n ? D[r * n + c1] : 0.f;
      float dot = fmaf(y0, d0, y1 * d1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = h ? c1 : c0;
        if (c >= n) continue;
        const float y = h ? y1 : y0, d = h ? d1 : d0;
        const float g = y * (d - dot);
        const int i = r * n + c;
        const float m = fmaf(omb1, g - M[i], M[i]);          // exp_avg.lerp_(grad, 1 - beta1)
        const float v = fmaf(omb2 * g, g, V[i] * b2);        // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + epsf;
        M[i] = m;
        V[i] = v;
        A[i] = A[i] - step_size * (m / denom);
      }
    }
    __syncthreads();
  }

  for (int i = tid; i < n * n; i += nt) {
    weight[i] = A[i];
    exp_avg[i] = M[i];
    exp_avg_sq[i] = V[i];
  }
  if (dT_accum)
    for (int i = tid; i < n * C; i += nt) dT_accum[i] += dTs[i];
}

}  // namespace simt

using namespace simt;

extern "C" int simt_w_fit(float* weight, float* exp_avg, float* exp_avg_sq, const float* T, int CK, int C, int n_steps,
                          long long step0, double lr, double beta1, double beta2, double eps, float* dT_accum,
                          float* losses, void* stream) {
  if (!weight || !exp_avg || !exp_avg_sq || !T) return SIMT_EINVAL;
  if (CK <= 1 || C <= 0 || n_steps < 0 || step0 < 0) return SIMT_EINVAL;
  if (CK > kMaxW || C > kMaxW) return SIMT_EUNSUPPORTED;
  if (n_steps == 0) return 0;
  const size_t smem = (size_t)(5 * CK * CK + 3 * CK * C) * sizeof(float);
  static bool attr_done[64] = {};
  {
    const int rc = ensure_dynamic_smem(w_fit_kernel, attr_done, 160 * 1024);
    if (rc) return rc;
  }
  w_fit_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(weight, exp_avg, exp_avg_sq, T, CK, C, n_steps, step0, lr, beta1,
                                                       beta2, eps, dT_accum, losses);
  return (int)cudaGetLastError();
}
