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
//   Levels 1/2 and run flushes go to a shared-memory histogram that is PRIVATE TO EACH LANE
//   (packed 8-bit counters, word index = (bin/4)*32 + lane, so bank == lane: no bank
//   conflicts, no atomics, no dependence on how contended a class is).  Before a byte can
//   overflow the warp folds its private counters into the CTA histogram (u32 shared
//   atomics, one per bin per warp), and at the end the CTA adds its histogram to the
//   caller's int64 table with one global atomic per non-zero bin.
//   A mode with plain shared atomics is kept for comparison (simt_hist_set_tuning).
// int64 inputs (the dtype label_mapping returns) take a generic, slower kernel.
#include "common.cuh"

namespace simt {

struct HistTuning { int mode, warps, unroll; };
static HistTuning g_hist_tuning = {0, 0, 0};

// ------------------------------------------------------------------------------------------
// uint8 fast path
// ------------------------------------------------------------------------------------------
struct HistArgs {
  const uint8_t* a;
  const uint8_t* b;      // null for the 1-D class histogram
  long long n;           // pixels
  const uint8_t* lut;    // null = identity
  int n_rows, n_cols, nbins, nwords;  // nwords = ceil(nbins/4)
  unsigned long long* hist;
  int* err;
};

template <bool PRIV>
struct Accum {
  unsigned* priv;      // lane-private packed counters of this warp (PRIV) -- [nwords][32]
  unsigned* cta_hist;  // [nbins] u32, shared atomics
  const uint8_t* lut_s;
  int n_rows, n_cols, nbins, lane;
  bool bad;

  __device__ __forceinline__ int bin_of(unsigned araw, unsigned b) {
    const unsigned a = lut_s[araw];
    if (a >= (unsigned)n_rows) return -1;  // (a >= 0) & (a < n) mask of the reference
    const int idx = (int)a * n_cols + (int)b;
    if (idx >= nbins) { bad = true; return -1; }  // numpy would raise on reshape
    return idx;
  }
  __device__ __forceinline__ void add(int idx, unsigned count) {
    if (idx < 0) return;
    if (PRIV) {
      priv[(idx >> 2) * 32 + lane] += count << ((idx & 3) * 8);
    } else {
      atomicAdd(&cta_hist[idx], count);
    }
  }
};

// fold the warp's lane-private byte counters into the CTA histogram and clear them
__device__ __noinline__ void fold_private(unsigned* priv, unsigned* cta_hist, int nwords, int nbins, int lane) {
  __syncwarp();
  for (int w0 = 0; w0 < nwords; w0 += 32) {
    const int w = w0 + lane;  // this lane sums word-row w over the 32 lane columns
    if (w < nwords) {
      unsigned lo = 0, hi = 0;  // 2 x 16-bit partial sums each (<= 32*255 < 65536)
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const unsigned v = priv[w * 32 + ((j + lane) & 31)];
        lo += v & 0x00ff00ffu;
        hi += (v >> 8) & 0x00ff00ffu;
      }
      const int bin = w * 4;
      const unsigned c0 = lo & 0xffffu, c1 = hi & 0xffffu, c2 = lo >> 16, c3 = hi >> 16;
      if (c0) atomicAdd(&cta_hist[bin], c0);
      if (c1 && bin + 1 < nbins) atomicAdd(&cta_hist[bin + 1], c1);
      if (c2 && bin + 2 < nbins) atomicAdd(&cta_hist[bin + 2], c2);
      if (c3 && bin + 3 < nbins) atomicAdd(&cta_hist[bin + 3], c3);
    }
  }
  __syncwarp();
  for (int i = lane; i < nwords * 32; i += 32) priv[i] = 0;
  __syncwarp();
}

template <bool HAS_B, bool PRIV, int UNROLL>
__global__ void __launch_bounds__(512) hist_u8_kernel(const HistArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* lut_s = smem_raw;                                          // 256 B
  unsigned* cta_hist = reinterpret_cast<unsigned*>(smem_raw + 256);   // nbins u32
  unsigned* priv_all = cta_hist + ((A.nbins + 3) & ~3);               // warps * nwords * 32
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;

  for (int i = tid; i < 256; i += blockDim.x) lut_s[i] = A.lut ? A.lut[i] : (uint8_t)i;
  for (int i = tid; i < A.nbins; i += blockDim.x) cta_hist[i] = 0;
  unsigned* priv = priv_all + (size_t)warp * A.nwords * 32;
  if (PRIV)
    for (int i = lane; i < A.nwords * 32; i += 32) priv[i] = 0;
  __syncthreads();

  Accum<PRIV> acc;
  acc.priv = priv; acc.cta_hist = cta_hist; acc.lut_s = lut_s;
  acc.n_rows = A.n_rows; acc.n_cols = A.n_cols; acc.nbins = A.nbins; acc.lane = lane; acc.bad = false;

  const long long ngroups = A.n >> 4;  // 16-pixel groups (pointers are 16-byte aligned)
  const uint4* a4 = reinterpret_cast<const uint4*>(A.a);
  const uint4* b4 = reinterpret_cast<const uint4*>(A.b);
  const long long gwarp = (long long)blockIdx.x * nwarps + warp;
  const long long total_warps = (long long)gridDim.x * nwarps;
  // every warp iteration covers UNROLL consecutive chunks of 32 groups
  const long long chunk = 32LL * UNROLL;
  const long long nchunks = (ngroups + chunk - 1) / chunk;

  int run_bin = -1;          // level-0 register run
  unsigned run_cnt = 0;
  int budget = 255;          // how much more a single byte counter of this warp may grow

  for (long long c = gwarp; c < nchunks; c += total_warps) {
    uint4 va[UNROLL], vb[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long g = c * chunk + (long long)u * 32 + lane;
      ok[u] = g < ngroups;
      if (ok[u]) {
        va[u] = ldg_stream_u4(a4 + g);
        if (HAS_B) vb[u] = ldg_stream_u4(b4 + g);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (PRIV) {
        // worst case this group adds 16 to one byte (plus a run flush handled below)
        if (budget < 16) { fold_private(priv, cta_hist, A.nwords, A.nbins, lane); budget = 255; }
      }
      bool slow = false;
      if (ok[u]) {
        const unsigned aw[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
        unsigned bw[4] = {0, 0, 0, 0};
        if (HAS_B) { bw[0] = vb[u].x; bw[1] = vb[u].y; bw[2] = vb[u].z; bw[3] = vb[u].w; }
        const unsigned a0 = aw[0] & 0xffu, b0 = bw[0] & 0xffu;
        const unsigned ar = a0 * 0x01010101u, br = b0 * 0x01010101u;
        const bool uni = (aw[0] == ar) & (aw[1] == ar) & (aw[2] == ar) & (aw[3] == ar) &
                         (bw[0] == br) & (bw[1] == br) & (bw[2] == br) & (bw[3] == br);
        if (uni) {
          const int idx = acc.bin_of(a0, b0);
          if (idx == run_bin) {
            run_cnt += 16;
          } else {
            // flush the finished run straight to the CTA histogram (rare)
            if (run_bin >= 0) atomicAdd(&cta_hist[run_bin], run_cnt);
            run_bin = idx; run_cnt = 16;
          }
        } else {
          slow = true;
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
        }
      }
      if (PRIV) {
        if (__any_sync(0xffffffffu, slow)) budget -= 16;
      }
    }
  }
  if (run_bin >= 0) atomicAdd(&cta_hist[run_bin], run_cnt);
  if (PRIV) fold_private(priv, cta_hist, A.nwords, A.nbins, lane);

  // tail pixels (n % 16) by one thread of CTA 0
  if (blockIdx.x == 0 && tid == 0) {
    for (long long i = ngroups << 4; i < A.n; ++i) {
      const int idx = acc.bin_of(A.a[i], HAS_B ? A.b[i] : 0u);
      if (idx >= 0) atomicAdd(&cta_hist[idx], 1u);
    }
  }
  if (acc.bad) atomicOr(A.err, SIMT_ERRBIT_PRED_RANGE);
  __syncthreads();
  for (int i = tid; i < A.nbins; i += blockDim.x) {
    const unsigned v = cta_hist[i];
    if (v) atomicAdd(&A.hist[i], (unsigned long long)v);
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

template <bool HAS_B, bool PRIV>
static int launch_u8(const HistArgs& A, int warps, int unroll, size_t smem, int grid, cudaStream_t st) {
#define SIMT_LAUNCH_U(U)                                                                                         \
  {                                                                                                              \
    auto k = hist_u8_kernel<HAS_B, PRIV, U>;                                                                     \
    SIMT_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
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
      A.nwords = ((int)nbins + 3) / 4; A.hist = h; A.err = err_flag;
      const bool priv = g_hist_tuning.mode != 2;
      int unroll = g_hist_tuning.unroll > 0 ? g_hist_tuning.unroll : 4;
      // shared memory: lut + cta hist + per-warp private counters; aim for 2 CTAs per SM
      const size_t per_warp = priv ? (size_t)A.nwords * 32 * 4 : 0;
      const size_t fixed = 256 + (size_t)((A.nbins + 3) & ~3) * 4;
      int warps = g_hist_tuning.warps > 0 ? g_hist_tuning.warps : 8;
      if (warps > 16) warps = 16;
      int ctas_per_sm = 2;
      if (priv) {
        const size_t budget2 = ((size_t)228 * 1024 - 2048) / 2;
        while (warps > 1 && fixed + per_warp * warps > budget2) --warps;
        if (fixed + per_warp * warps > (size_t)di.smem_optin) return SIMT_ENOSMEM;
      } else {
        ctas_per_sm = 2048 / (warps * 32);
        if (ctas_per_sm < 1) ctas_per_sm = 1;
      }
      const size_t smem = fixed + per_warp * warps;
      const long long ngroups = len >> 4;
      long long need = (ngroups + 32LL * unroll * warps - 1) / (32LL * unroll * warps);
      long long grid = (long long)di.sm_count * ctas_per_sm;
      if (grid > need) grid = need;
      if (grid < 1) grid = 1;
      if (b) rc = priv ? launch_u8<true, true>(A, warps, unroll, smem, (int)grid, st)
                       : launch_u8<true, false>(A, warps, unroll, smem, (int)grid, st);
      else   rc = priv ? launch_u8<false, true>(A, warps, unroll, smem, (int)grid, st)
                       : launch_u8<false, false>(A, warps, unroll, smem, (int)grid, st);
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

void simt_hist_set_tuning(int mode, int warps_per_cta, int unroll) { g_hist_tuning = {mode, warps_per_cta, unroll}; }

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
  if (!dummy_err[dev]) {
    SIMT_CUDA_TRY(cudaMalloc(&dummy_err[dev], sizeof(int)));
    SIMT_CUDA_TRY(cudaMemset(dummy_err[dev], 0, sizeof(int)));
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
