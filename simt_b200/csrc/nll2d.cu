// CrossEntropy2d(is_softmax=False) on already-mixed probabilities [B, C, H, W]
// (utils/loss.py:14-40 of the reference, forward and backward) for callers that keep the
// reference's unfused upsample/softmax/mm lines.  A gather + log + masked mean; the
// fused head (head.cu) is the fast path, this is the literal drop-in.
#include "common.cuh"

namespace simt {

template <typename LabelT>
__device__ __forceinline__ long long label_at(const LabelT* p, long long i) { return (long long)__ldg(p + i); }

template <typename LabelT>
__global__ void __launch_bounds__(256) nll2d_fwd_kernel(const float* __restrict__ prob, int C, long long HW,
                                                        long long npix, const LabelT* __restrict__ labels, int ignore,
                                                        double* __restrict__ stats, int* __restrict__ err) {
  double acc = 0;
  long long cnt = 0;
  bool bad = false;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
    const long long y = label_at(labels, i);
    if (y < 0 || y == ignore) continue;
    if (y >= C) { bad = true; continue; }
    const long long b = i / HW, r = i - b * HW;
    acc -= (double)logf(__ldg(prob + (b * C + y) * HW + r));
    ++cnt;
  }
  if (bad) atomicOr(err, SIMT_ERRBIT_LABEL_RANGE);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ double sa[8];
  __shared__ long long sc[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sa[w] = acc; sc[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0; long long tc = 0;
    for (int k = 0; k < 8; ++k) { ta += sa[k]; tc += sc[k]; }
    atomicAdd(&stats[0], ta);
    atomicAdd(&stats[1], (double)tc);
  }
}

__global__ void nll2d_mean_kernel(const double* __restrict__ stats, float* __restrict__ loss_mean,
                                  const int* __restrict__ err) {
  float m = (float)(stats[0] / stats[1]);
  if (*err & SIMT_ERRBIT_LABEL_RANGE) m = nanf("");
  *loss_mean = m;
}

template <typename LabelT>
__global__ void __launch_bounds__(256) nll2d_bwd_kernel(const float* __restrict__ prob, int C, long long HW,
                                                        long long npix, const LabelT* __restrict__ labels, int ignore,
                                                        const double* __restrict__ stats,
                                                        const float* __restrict__ grad_out, float* __restrict__ dprob) {
  const float s = (float)((grad_out ? (double)__ldg(grad_out) : 1.0) / stats[1]);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
    const long long y = label_at(labels, i);
    if (y < 0 || y == ignore || y >= C) continue;
    const long long b = i / HW, r = i - b * HW;
    const long long o = (b * C + y) * HW + r;
    dprob[o] = -s / __ldg(prob + o);
  }
}

static int grid_for(long long npix, int* grid) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  long long g = (npix + 255) / 256;
  if (g > (long long)di.sm_count * 8) g = (long long)di.sm_count * 8;
  *grid = g < 1 ? 1 : (int)g;
  return 0;
}

}  // namespace simt

using namespace simt;

extern "C" {

int simt_nll2d_fwd(const float* prob, int B, int C, int H, int W, const void* labels, int label_bytes, int ignore,
                   double* stats, float* loss_mean, int* err_flag, void* stream) {
  if (!prob || !labels || !stats || !err_flag || B <= 0 || C <= 0 || H <= 0 || W <= 0) return SIMT_EINVAL;
  if (label_bytes != 1 && label_bytes != 8) return SIMT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const long long HW = (long long)H * W, npix = HW * B;
  int grid;
  int rc = grid_for(npix, &grid);
  if (rc) return rc;
  SIMT_CUDA_TRY(cudaMemsetAsync(stats, 0, 2 * sizeof(double), st));
  if (label_bytes == 1)
    nll2d_fwd_kernel<uint8_t><<<grid, 256, 0, st>>>(prob, C, HW, npix, (const uint8_t*)labels, ignore, stats, err_flag);
  else
    nll2d_fwd_kernel<long long><<<grid, 256, 0, st>>>(prob, C, HW, npix, (const long long*)labels, ignore, stats, err_flag);
  SIMT_CUDA_TRY(cudaGetLastError());
  if (loss_mean) nll2d_mean_kernel<<<1, 1, 0, st>>>(stats, loss_mean, err_flag);
  return (int)cudaGetLastError();
}

int simt_nll2d_bwd(const float* prob, int B, int C, int H, int W, const void* labels, int label_bytes, int ignore,
                   const double* stats, const float* grad_out, float* dprob, void* stream) {
  if (!prob || !labels || !stats || !dprob || B <= 0 || C <= 0 || H <= 0 || W <= 0) return SIMT_EINVAL;
  if (label_bytes != 1 && label_bytes != 8) return SIMT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const long long HW = (long long)H * W, npix = HW * B;
  int grid;
  int rc = grid_for(npix, &grid);
  if (rc) return rc;
  SIMT_CUDA_TRY(cudaMemsetAsync(dprob, 0, (size_t)npix * C * sizeof(float), st));
  if (label_bytes == 1)
    nll2d_bwd_kernel<uint8_t><<<grid, 256, 0, st>>>(prob, C, HW, npix, (const uint8_t*)labels, ignore, stats, grad_out, dprob);
  else
    nll2d_bwd_kernel<long long><<<grid, 256, 0, st>>>(prob, C, HW, npix, (const long long*)labels, ignore, stats, grad_out, dprob);
  return (int)cudaGetLastError();
}

}  // extern "C"
