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

// (workspace header words WS_*: step_xchg.cuh; head_prep_kernel, head_finalize_kernel, head_finish_kernel: step_kernels.cuh)
}  // namespace simt
#include "step_kernels.cuh"
namespace simt {

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

static int make_plan(int mode, int B, int CK, int C, int h, int w, int H, int W, HeadArgs* A, Plan* P) {
  const Tuning tune = current_tuning();
  DeviceInfo di;
  const int sm_count = device_info(&di) == 0 ? di.sm_count : 0;
  return make_plan_for(mode, B, CK, C, h, w, H, W, PlanTuning{tune.ur, tune.unused, tune.lpr}, sm_count, A, P);
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
