// Integer eval histograms for B200 (sm_100a): confusion matrix, class histogram and the
// raw-id -> train-id LUT.  Bit-exact replacements for
//   tools/compute_iou.py:9-11        fast_hist(a, b, n)            (+ :18-22 label_mapping fused as a LUT)
//   tools/compute_ConfusionMatrix.py:54-56   fast_hist(a, b, n33, n19)
//   tools/compute_ClassDistribution.py:52-54 fast_hist(a, n)
//
// These ARE memory-bound (2 B/pixel resp. 1 B/pixel).  Design of the uint8 fast path:
//   * persistent grid, every lane streams 16 pixels per 128-bit load (ld.global.nc,
//     L1::no_allocate), UNROLL loads in flight per lane;
//   * level 0: a 16-pixel group whose (a, b) pairs are all equal (the common case on
//     real label maps) extends a register run (bin, count) -- no shared-memory traffic;
//   * level 1: a 4-pixel word whose pairs are all equal adds 4 at once;
//   * level 2: single pixels.
//   Levels 1/2 and finished runs go to the CTA's shared-memory histogram with NATIVE u32 shared
//   atomics (measured on B200: ~4 lane-atomics/clk/SM even on random bins, which beat
//   lane-private packed counters by 2x).  The histogram is replicated kRep times (copy = lane %
//   kRep, copies on different banks) so that a hot class -- road is 41 % of real labels -- is not
//   a 32-way same-address conflict.  At the end the CTA adds its copies to the caller's int64
//   table with one global atomic per non-zero bin.
// int64 inputs (the dtype label_mapping returns) take a generic, slower kernel.
#include <mutex>
#include "common.cuh"

namespace simt {

// Benchmark tuning hook (simt_hist_set_tuning) and the lazily allocated scratch word of simt_class_hist: process-global,
// guarded by one mutex so that the entry points may be called from several host threads.
struct HistTuning { int mode, warps, unroll; };
static HistTuning g_hist_tuning = {0, 0, 0};
static std::mutex g_hist_mutex;
static HistTuning current_hist_tuning() {
  std::lock_guard<std::mutex> lock(g_hist_mutex);
  return g_hist_tuning;
}

// ------------------------------------------------------------------------------------------
// uint8 fast path
// ------------------------------------------------------------------------------------------
struct HistArgs {
  const uint8_t* a;
  const uint8_t* b;      // null for the 1-D class histogram
  long long n;           // pixels
  const uint8_t* lut;    // null = identity
  int n_rows, n_cols, nbins;
  int probe;             // benchmark probe: only stream the inputs (load-path ceiling), counts are garbage
  unsigned long long* hist;
  int* err;
};

static constexpr int kRep = 8;  // shared histogram copies

struct Accum {
  unsigned* my_hist;  // this lane's copy of the CTA histogram ([nbins] u32, shared atomics)
  const uint8_t* lut_s;
  int n_rows, n_cols, nbins;
  int probe;             // benchmark probe: only stream the inputs (load-path ceiling), counts are garbage
  bool bad;

  __device__ __forceinline__ int bin_of(unsigned araw, unsigned b) {
    const unsigned a = lut_s[araw];
    if (a >= (unsigned)n_rows) return -1;  // (a >= 0) & (a < n) mask of the reference
    const int idx = (int)a * n_cols + (int)b;
    if (idx >= nbins) { bad = true; return -1; }  // numpy would raise on reshape
    return idx;
  }
  __device__ __forceinline__ void add(int idx, unsigned count) {
    if (idx >= 0) atomicAdd(&my_hist[idx], count);
  }
};

template <bool HAS_B, int UNROLL>
__global__ void __launch_bounds__(512) hist_u8_kernel(const HistArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* lut_s = smem_raw;                                          // 256 B
  unsigned* cta_hist = reinterpret_cast<unsigned*>(smem_raw + 256);   // kRep copies, stride hstride
  const int hstride = A.nbins | 1;  // odd stride: the copies of one bin sit in different banks
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;

  for (int i = tid; i < 256; i += blockDim.x) lut_s[i] = A.lut ? A.lut[i] : (uint8_t)i;
  for (int i = tid; i < kRep * hstride; i += blockDim.x) cta_hist[i] = 0;
  __syncthreads();

  Accum acc;
  acc.my_hist = cta_hist + (lane % kRep) * hstride; acc.lut_s = lut_s;
  acc.n_rows = A.n_rows; acc.n_cols = A.n_cols; acc.nbins = A.nbins; acc.bad = false;

  // a lane-iteration is one 128-bit load of a and one of b (16 pixels), or -- single-array class histogram --
  // two consecutive 128-bit loads of a (32 pixels), so both cases move 32 B per lane per iteration
  constexpr int GPX = HAS_B ? 16 : 32;
  const long long ngroups = A.n / GPX;  // (pointers are 16-byte aligned)
  const uint4* a4 = reinterpret_cast<const uint4*>(A.a);
  const uint4* b4 = reinterpret_cast<const uint4*>(A.b);
  const long long gwarp = (long long)blockIdx.x * nwarps + warp;
  const long long total_warps = (long long)gridDim.x * nwarps;
  // every warp iteration covers UNROLL consecutive chunks of 32 groups
  const long long chunk = 32LL * UNROLL;
  const long long nchunks = (ngroups + chunk - 1) / chunk;

  int run_bin = -1;          // level-0 register run
  unsigned run_cnt = 0;

  for (long long c = gwarp; c < nchunks; c += total_warps) {
    uint4 va[UNROLL], vb[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long g = c * chunk + (long long)u * 32 + lane;
      ok[u] = g < ngroups;
      if (ok[u]) {
        va[u] = ldg_stream_u4(HAS_B ? a4 + g : a4 + 2 * g);
        vb[u] = ldg_stream_u4(HAS_B ? b4 + g : a4 + 2 * g + 1);
      }
    }
    if (A.probe) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (ok[u]) run_cnt ^= va[u].x ^ va[u].y ^ va[u].z ^ va[u].w ^ vb[u].x ^ vb[u].y ^ vb[u].z ^ vb[u].w;
      continue;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (!ok[u]) continue;
      const unsigned aw[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
      const unsigned bw[4] = {vb[u].x, vb[u].y, vb[u].z, vb[u].w};   // b pixels, or the next 16 a pixels
      const unsigned a0 = aw[0] & 0xffu, b0 = HAS_B ? (bw[0] & 0xffu) : 0u;
      const unsigned ar = a0 * 0x01010101u, br = HAS_B ? b0 * 0x01010101u : ar;
      // one LOP3 per word pair: (x ^ r) | (y ^ r)
      const unsigned da = ((aw[0] ^ ar) | (aw[1] ^ ar)) | ((aw[2] ^ ar) | (aw[3] ^ ar));
      const unsigned db = ((bw[0] ^ br) | (bw[1] ^ br)) | ((bw[2] ^ br) | (bw[3] ^ br));
      if ((da | db) == 0u) {
        const int idx = acc.bin_of(a0, b0);
        if (idx == run_bin) {
          run_cnt += GPX;
        } else {
          acc.add(run_bin, run_cnt);  // finished run (rare)
          run_bin = idx; run_cnt = GPX;
        }
      } else if (HAS_B) {
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
          const unsigned x = aw[wd], y = bw[wd];
          const unsigned xa = x & 0xffu, yb = y & 0xffu;
          if (x == xa * 0x01010101u && y == yb * 0x01010101u) {
            acc.add(acc.bin_of(xa, yb), 4u);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.add(acc.bin_of((x >> (8 * q)) & 0xffu, (y >> (8 * q)) & 0xffu), 1u);
          }
        }
      } else {
#pragma unroll
        for (int wd = 0; wd < 8; ++wd) {
          const unsigned x = wd < 4 ? aw[wd & 3] : bw[wd & 3];
          const unsigned xa = x & 0xffu;
          if (x == xa * 0x01010101u) {
            acc.add(acc.bin_of(xa, 0u), 4u);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.add(acc.bin_of((x >> (8 * q)) & 0xffu, 0u), 1u);
          }
        }
      }
    }
  }
  acc.add(run_bin, run_cnt);

  // tail pixels (n % GPX) by one thread of CTA 0
  if (blockIdx.x == 0 && tid == 0) {
    for (long long i = ngroups * GPX; i < A.n; ++i) acc.add(acc.bin_of(A.a[i], HAS_B ? A.b[i] : 0u), 1u);
  }
  if (acc.bad) atomicOr(A.err, SIMT_ERRBIT_PRED_RANGE);
  __syncthreads();
  for (int i = tid; i < A.nbins; i += blockDim.x) {
    unsigned long long v = 0;
#pragma unroll
    for (int r = 0; r < kRep; ++r) v += cta_hist[r * hstride + i];
    if (v) atomicAdd(&A.hist[i], v);
  }
}

// ------------------------------------------------------------------------------------------
// generic path: any mix of uint8 / int64 inputs, any alignment, any table size
// ------------------------------------------------------------------------------------------
template <typename TA, typename TB>
__global__ void __launch_bounds__(256) hist_generic_kernel(const TA* __restrict__ a, const TB* __restrict__ b,
                                                            long long n, const uint8_t* __restrict__ lut, int n_rows,
                                                            int n_cols, long long nbins, int smem_bins,
                                                            unsigned long long* __restrict__ hist, int* err) {
  extern __shared__ unsigned sh[];
  for (int i = threadIdx.x; i < smem_bins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  bool bad = false;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long av = (long long)a[i];
    if (lut) av = lut[(unsigned)av & 0xffu];
    if (av < 0 || av >= n_rows) continue;
    long long bv = b ? (long long)b[i] : 0;
    long long idx = av * n_cols + bv;
    if (idx < 0 || idx >= nbins) { bad = true; continue; }
    if (smem_bins) atomicAdd(&sh[idx], 1u); else atomicAdd(&hist[idx], 1ULL);
  }
  if (bad) atomicOr(err, SIMT_ERRBIT_PRED_RANGE);
  __syncthreads();
  for (int i = threadIdx.x; i < smem_bins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

__global__ void __launch_bounds__(256) label_map_kernel(const uint8_t* __restrict__ in, long long n,
                                                        const uint8_t* __restrict__ lut, long long* __restrict__ out) {
  __shared__ uint8_t lut_s[256];
  lut_s[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const long long n4 = aligned ? (n >> 2) : 0;
  const unsigned* in4 = reinterpret_cast<const unsigned*>(in);
  longlong2* out2 = reinterpret_cast<longlong2*>(out);
  for (long long i = i0; i < n4; i += stride) {
    const unsigned v = __ldg(in4 + i);
    out2[2 * i] = make_longlong2(lut_s[v & 0xff], lut_s[(v >> 8) & 0xff]);
    out2[2 * i + 1] = make_longlong2(lut_s[(v >> 16) & 0xff], lut_s[v >> 24]);
  }
  for (long long i = n4 * 4 + i0; i < n; i += stride) out[i] = lut_s[in[i]];
}

template <bool HAS_B>
static int launch_u8(const HistArgs& A, int warps, int unroll, size_t smem, int grid, cudaStream_t st) {
#define SIMT_LAUNCH_U(U)                                                                                         \
  {                                                                                                              \
    auto k = hist_u8_kernel<HAS_B, U>;                                                                           \
    prof_begin(st);                                                                                              \
    k<<<grid, warps * 32, smem, st>>>(A);                                                                        \
    prof_end(st);                                                                                                \
    return (int)cudaGetLastError();                                                                              \
  }
  if (unroll >= 4) SIMT_LAUNCH_U(4)
  if (unroll >= 2) SIMT_LAUNCH_U(2)
  SIMT_LAUNCH_U(1)
#undef SIMT_LAUNCH_U
}

static constexpr long long kMaxPerLaunch = 1LL << 31;  // keeps the u32 CTA histogram exact

static int run_hist(const void* a, int a_bytes, const void* b, int b_bytes, long long n, const uint8_t* lut,
                    int n_rows, int n_cols, long long* hist, int* err_flag, cudaStream_t st) {
  if (n == 0) return 0;
  if (!a || !hist || n < 0 || n_rows <= 0 || n_cols <= 0) return SIMT_EINVAL;
  if ((a_bytes != 1 && a_bytes != 8) || (b && b_bytes != 1 && b_bytes != 8)) return SIMT_EINVAL;
  if (lut && a_bytes != 1) return SIMT_EINVAL;
  if (!err_flag) return SIMT_EINVAL;
  if (n == 0) return 0;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  const long long nbins = (long long)n_rows * n_cols;
  const HistTuning tune = current_hist_tuning();
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
  const bool fast = a_bytes == 1 && (!b || b_bytes == 1) && nbins <= 1024 && n_rows <= 256 &&
                    (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (!b || (reinterpret_cast<uintptr_t>(b) & 15) == 0);
  for (long long off = 0; off < n; off += kMaxPerLaunch) {
    const long long len = (n - off < kMaxPerLaunch) ? (n - off) : kMaxPerLaunch;
    if (fast) {
      HistArgs A{};
      A.a = static_cast<const uint8_t*>(a) + off;
      A.b = b ? static_cast<const uint8_t*>(b) + off : nullptr;
      A.n = len; A.lut = lut; A.n_rows = n_rows; A.n_cols = n_cols; A.nbins = (int)nbins;
      A.hist = h; A.err = err_flag; A.probe = (tune.mode == 9);
      // more loads in flight per lane pay off once the launch is long enough to reach steady state
      int unroll = tune.unroll > 0 ? tune.unroll : (len >= (1LL << 29) ? 4 : 1);
      int warps = tune.warps > 0 ? tune.warps : 16;
      if (warps > 16) warps = 16;
      const size_t smem = 256 + (size_t)kRep * ((size_t)A.nbins | 1) * 4;   // <= 33 KB for 1024 bins
      int ctas_per_sm = 2048 / (warps * 32);
      if (ctas_per_sm < 1) ctas_per_sm = 1;
      const long long ngroups = len / (b ? 16 : 32);
      long long need = (ngroups + 32LL * unroll * warps - 1) / (32LL * unroll * warps);
      long long grid = (long long)di.sm_count * ctas_per_sm;
      if (grid > need) grid = need;
      if (grid < 1) grid = 1;
      rc = b ? launch_u8<true>(A, warps, unroll, smem, (int)grid, st) : launch_u8<false>(A, warps, unroll, smem, (int)grid, st);
      if (rc) return rc;
    } else {
      const int smem_bins = nbins <= 8192 ? (int)nbins : 0;
      long long grid = (len + 255) / 256;
      if (grid > (long long)di.sm_count * 8) grid = (long long)di.sm_count * 8;
      const size_t smem = (size_t)smem_bins * 4;
#define SIMT_GEN(TA, TB)                                                                                        \
  hist_generic_kernel<TA, TB><<<(int)grid, 256, smem, st>>>(static_cast<const TA*>(a) + off,                     \
                                                              b ? static_cast<const TB*>(b) + off : nullptr, len, \
                                                              lut, n_rows, n_cols, nbins, smem_bins, h, err_flag)
      if (a_bytes == 1 && (!b || b_bytes == 1)) SIMT_GEN(uint8_t, uint8_t);
      else if (a_bytes == 1) SIMT_GEN(uint8_t, long long);
      else if (!b || b_bytes == 8) SIMT_GEN(long long, long long);
      else SIMT_GEN(long long, uint8_t);
#undef SIMT_GEN
      SIMT_CUDA_TRY(cudaGetLastError());
    }
  }
  return 0;
}

}  // namespace simt

using namespace simt;

extern "C" {

void simt_hist_set_tuning(int mode, int warps_per_cta, int unroll) {
  std::lock_guard<std::mutex> lock(g_hist_mutex);
  g_hist_tuning = {mode, warps_per_cta, unroll};
}

int simt_confusion(const void* a, int a_bytes, const void* b, int b_bytes, long long n, const uint8_t* lut256,
                   int n_rows, int n_cols, long long* hist, int* err_flag, void* stream) {
  if (!b && n != 0) return SIMT_EINVAL;
  return run_hist(a, a_bytes, b, b_bytes, n, lut256, n_rows, n_cols, hist, err_flag, (cudaStream_t)stream);
}

int simt_class_hist(const void* a, int a_bytes, long long n, int n_bins, long long* hist, void* stream) {
  // no second operand: a pixel is either counted or masked out, there is nothing to flag;
  // a private scratch word keeps the kernel signature uniform
  static int* dummy_err[64] = {};
  int dev = 0;
  SIMT_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SIMT_EUNSUPPORTED;
  {
    std::lock_guard<std::mutex> lock(g_hist_mutex);
    if (!dummy_err[dev]) {
      SIMT_CUDA_TRY(cudaMalloc(&dummy_err[dev], sizeof(int)));
      SIMT_CUDA_TRY(cudaMemset(dummy_err[dev], 0, sizeof(int)));
    }
  }
  return run_hist(a, a_bytes, nullptr, 0, n, nullptr, n_bins, 1, hist, dummy_err[dev], (cudaStream_t)stream);
}

int simt_label_map(const uint8_t* in, long long n, const uint8_t* lut256, long long* out, void* stream) {
  if (!in || !lut256 || !out || n < 0) return SIMT_EINVAL;
  if (n == 0) return 0;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  long long grid = (n / 4 + 255) / 256;
  if (grid > (long long)di.sm_count * 16) grid = (long long)di.sm_count * 16;
  if (grid < 1) grid = 1;
  label_map_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(in, n, lut256, out);
  return (int)cudaGetLastError();
}

}  // extern "C"
