// ABI version, error strings and the launch profiler shared by all kernels.
#include <mutex>
#include <vector>
#include "common.cuh"

namespace simt {

// Opt-in per-launch timing of the DOMINANT kernel of each entry point (bench.py's roofline
// numbers): a pair of CUDA events is recorded on the caller's stream right around that kernel.
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> starts, stops;
  size_t used = 0;
};
static Profiler g_prof;
static std::mutex g_prof_mutex;   // the profiler is process-global; launches may come from several host threads

bool prof_enabled() {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  return g_prof.on;
}

// events recorded into a capturing stream become graph nodes whose timestamps cannot be read back: skip them
static bool capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
}

void prof_begin(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (!g_prof.on || capturing(st)) return;
  if (g_prof.used == g_prof.starts.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    g_prof.starts.push_back(a);
    g_prof.stops.push_back(b);
  }
  cudaEventRecord(g_prof.starts[g_prof.used], st);
}

void prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (!g_prof.on || g_prof.used >= g_prof.stops.size() || capturing(st)) return;
  cudaEventRecord(g_prof.stops[g_prof.used], st);
  ++g_prof.used;
}

}  // namespace simt

using namespace simt;

extern "C" {

int simt_b200_abi_version(void) { return SIMT_B200_ABI_VERSION; }

const char* simt_b200_strerror(int code) {
  switch (code) {
    case 0: return "success";
    case SIMT_EINVAL: return "simt_b200: invalid argument";
    case SIMT_EUNSUPPORTED: return "simt_b200: unsupported shape";
    case SIMT_EWORKSPACE: return "simt_b200: workspace too small";
    case SIMT_ENOSMEM: return "simt_b200: tile does not fit in shared memory";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "simt_b200: unknown error";
}

void simt_b200_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_prof.on = on != 0;
  g_prof.used = 0;
}

int simt_b200_profile_read(double* total_ms, long long* launches) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  double tot = 0;
  for (size_t i = 0; i < g_prof.used; ++i) {
    SIMT_CUDA_TRY(cudaEventSynchronize(g_prof.stops[i]));
    float ms = 0;
    SIMT_CUDA_TRY(cudaEventElapsedTime(&ms, g_prof.starts[i], g_prof.stops[i]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (long long)g_prof.used;
  g_prof.used = 0;
  return 0;
}

int simt_debug_resize_tables(int in, int out, int* cell, float* lambda, int* first_px) {
  if (in <= 0 || out <= 0 || !cell || !lambda || !first_px) return SIMT_EINVAL;
  // exactly make_plan()'s scale / cell count (head.cu) and the tables head_kernel builds in its prologue
  const float scale = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f;
  const int ncell = in > 1 ? in - 1 : 1;
  for (int X = 0; X < out; ++X) {
    cell[X] = cell_of(X, scale, ncell);
    lambda[X] = lambda_of(X, scale, cell[X]);
  }
  for (int c = 0; c <= ncell; ++c) first_px[c] = first_px_of_cell(c, scale, ncell, out);
  return 0;
}

}  // extern "C"
