// Sharded training step: the all-reduce of the head's `stats` buffer fused into the scale kernel, over peer memory.
//
// With the batch sharded over N GPUs (one process per GPU) the only exchange of a step is the 2 + CK*C doubles
// {loss sum, valid count, raw dT} (2.9 KB at CK = C = 19): the mean is over the GLOBAL valid-pixel count and dT is
// summed over ranks (the reference's single-process loss.backward() at tools/trainV2_simt.py:408-409,428 sees the
// whole batch).  A library all-reduce for 2.9 KB costs a launch plus a protocol round trip per step on the critical
// path (kernel -> finalize -> all-reduce -> scale); here the scale kernel does the exchange itself:
//   * every rank owns a MAILBOX (cudaMalloc'd here, exported with cudaIpcGetMemHandle, opened by the peers, so all
//     ranks of a node hold a pointer to every mailbox; stores travel over NVLink / NVSwitch);
//   * block 0 of the scale kernel PUSHES the rank's stats into slot [parity][rank] of every mailbox (its own too),
//     fences at system scope and then releases flag [parity][rank] = step number in every mailbox;
//   * every block acquires the `world` flags of its OWN mailbox, sums the valid counts in rank order, scales its
//     slice of dLogits; block 0 also sums all stats in rank order (bitwise identical on every rank) into the global
//     stats / dT / loss.
// Slots are double-buffered by step parity: a peer can only write step s+2 after it has seen this rank's flag of step
// s+1, which is sent after this rank's scale kernel of step s has finished reading.  The step number lives in the
// mailbox and is advanced by the last block to finish, so the launch has no per-step host argument and the whole
// step (memset, head kernel, finalize, this kernel) replays from a CUDA graph.  Waits are bounded: a peer that never
// arrives sets SIMT_ERRBIT_XCHG_TIMEOUT instead of hanging the GPU.
#include "xchg.cuh"

namespace simt {

__global__ void __launch_bounds__(256) head_scale_xchg_kernel(float* __restrict__ dlogits, long long n, double* stats,
                                                               int nT, const float* __restrict__ grad_out,
                                                               float* __restrict__ dT, float* __restrict__ loss_mean,
                                                               int* __restrict__ err, const XchgArgs X) {
  unsigned char* own = X.mail[X.rank];
  unsigned long long* hdr = hdr_of(own);
  __shared__ double s_cnt;
  __shared__ int s_timeout;
  const int tid = threadIdx.x;
  const unsigned long long seq = step_seq(own);
  const int par = (int)(seq & 1ULL);
  if (tid == 0) s_timeout = 0;

  if (blockIdx.x == 0) {
    // ---- push this rank's stats into every mailbox, then publish ---------------------------------
    for (int r = 0; r < X.world; ++r) {
      double* dst = slot_of(X.mail[r], par, X.rank, X.n_stats);
      for (int i = tid; i < X.n_stats; i += blockDim.x) dst[i] = stats[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < X.world)
      st_release_sys(hdr_of(X.mail[tid]) + kHdrStatFlag + par * kMaxPeers + X.rank, seq);
  }
  __syncthreads();
  // ---- wait for every rank's contribution to arrive in OUR mailbox (bounded) ----------------------
  if (tid < X.world && !wait_flag(hdr + kHdrStatFlag + par * kMaxPeers + tid, seq)) s_timeout = 1;  // the peer is gone
  __syncthreads();
  if (tid == 0) {
    double c = 0.0;
    for (int r = 0; r < X.world; ++r) c += ld_volatile_f64(slot_of(own, par, r, X.n_stats) + 1);  // rank order
    s_cnt = c;
    if (s_timeout && err) atomicOr(err, SIMT_ERRBIT_XCHG_TIMEOUT);
  }
  __syncthreads();
  const double cnt = s_cnt;
  const float s = (float)((grad_out ? (double)__ldg(grad_out) : 1.0) / cnt);

  // ---- scale this block's slice of dLogits -------------------------------------------------------
  const long long i0 = (long long)blockIdx.x * blockDim.x + tid;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = ((reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) ? (n >> 2) : 0;
  float4* d4 = reinterpret_cast<float4*>(dlogits);
  for (long long i = i0; i < n4; i += stride) {
    float4 v = d4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    d4[i] = v;
  }
  for (long long i = n4 * 4 + i0; i < n; i += stride) dlogits[i] *= s;

  // ---- block 0: the all-reduced stats (fixed rank order -> the same bits on every rank), dT, loss ---
  if (blockIdx.x == 0) {
    for (int i = tid; i < X.n_stats; i += blockDim.x) {
      double t = 0.0;
      for (int r = 0; r < X.world; ++r) t += ld_volatile_f64(slot_of(own, par, r, X.n_stats) + i);
      stats[i] = t;
      if (i >= 2 && dT && i - 2 < nT) dT[i - 2] = (float)(t * (double)s);
      if (i == 0 && loss_mean) {
        float m = (float)(t / cnt);   // 0/0 -> NaN like the reference's mean over nothing
        if (err && (*err & SIMT_ERRBIT_LABEL_RANGE)) m = nanf("");
        *loss_mean = m;
      }
    }
  }
  // ---- the last block to finish advances the step counter ------------------------------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(hdr + 1, 1ULL);
    if (t == (unsigned long long)gridDim.x - 1ULL) {
      hdr[1] = 0ULL;
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(hdr) = seq;
    }
  }
}

}  // namespace simt

using namespace simt;

extern "C" {

size_t simt_xchg_bytes(int n_stats) {
  if (n_stats <= 0) return 0;
  return kHdrBytes + (size_t)2 * kMaxPeers * (size_t)n_stats * sizeof(double);
}

int simt_xchg_create(size_t bytes, void** mailbox, unsigned char* handle64) {
  if (!mailbox || !handle64 || bytes < kHdrBytes) return SIMT_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle is exchanged as 64 raw bytes");
  void* p = nullptr;
  SIMT_CUDA_TRY(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  memcpy(handle64, &h, 64);
  *mailbox = p;
  return 0;
}

int simt_xchg_open(const unsigned char* handle64, void** peer_mailbox) {
  if (!handle64 || !peer_mailbox) return SIMT_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SIMT_CUDA_TRY(cudaIpcOpenMemHandle(peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int simt_xchg_close(void* peer_mailbox) {
  if (!peer_mailbox) return SIMT_EINVAL;
  return (int)cudaIpcCloseMemHandle(peer_mailbox);
}

int simt_xchg_destroy(void* mailbox) {
  if (!mailbox) return SIMT_EINVAL;
  return (int)cudaFree(mailbox);
}

int simt_head_scale_sharded(float* dlogits, long long n_dlogits, double* stats, int CK, int C, const float* grad_out,
                            float* dT, float* loss_mean, int rank, int world, void* const* mailboxes, int* err_flag,
                            void* stream) {
  if (!stats || !mailboxes || n_dlogits < 0 || (n_dlogits > 0 && !dlogits)) return SIMT_EINVAL;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || CK <= 0 || C <= 0) return SIMT_EINVAL;
  XchgArgs X{};
  for (int r = 0; r < world; ++r) {
    if (!mailboxes[r]) return SIMT_EINVAL;
    X.mail[r] = static_cast<unsigned char*>(mailboxes[r]);
  }
  X.rank = rank; X.world = world; X.n_stats = 2 + CK * C;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  long long blocks = (n_dlogits / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (long long)di.sm_count * 4) blocks = (long long)di.sm_count * 4;   // all co-resident
  head_scale_xchg_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dlogits, n_dlogits, stats, dT ? CK * C : 0,
                                                                        grad_out, dT, loss_mean, err_flag, X);
  return (int)cudaGetLastError();
}

}  // extern "C"
