// Mailboxes of the sharded training step (CUDA-IPC peer memory; layout in xchg.cuh).
//
// With the batch sharded over N GPUs (one process per GPU) a step exchanges two tiny things: the ranks' valid-pixel
// counts (8 bytes each, BEFORE gradients are final: the mean is over the GLOBAL count) and the 2 + CK*C doubles
// {loss sum, valid count, raw dT} (2.9 KB at CK = C = 19, summed over ranks) -- the reference's single-process
// loss.backward() at tools/trainV2_simt.py:408-409,428 sees the whole batch.  Neither goes through a library collective:
//   * every rank owns a MAILBOX (cudaMalloc'd here, exported with cudaIpcGetMemHandle, opened by the peers, so all
//     ranks of a node hold a pointer to every mailbox; stores travel over NVLink / NVSwitch);
//   * head_prep_kernel (first kernel of the step) pushes the count, the fused kernel acquires the counts right before
//     its first dLogits update, head_finalize_kernel (last kernel) pushes the stats, waits for the peers' and sums
//     them in rank order (csrc/head.cu).  No pass over dLogits, no memset, no extra launch.
// Slots are double-buffered by step parity: a peer can only write step s+2 after it has seen this rank's stats flag of
// step s+1, which is sent after every kernel of this rank's step s has finished reading.  The step number lives in the
// mailbox and is advanced by the last block of finalize, so the launches take no per-step host argument and the whole
// step replays from a CUDA graph.  Waits are bounded (simt_xchg_set_timeout): a peer that never arrives poisons
// loss / dT / dLogits with NaN and raises SIMT_ERRBIT_XCHG_TIMEOUT -- a rank never continues with a partial sum.
#include <atomic>
#include <cstring>
#include "xchg.cuh"

namespace simt {

// Polls per wait before a rank gives up on a peer (each poll is a system-scope load, ~1 us with its back-off: the
// default is of the order of 10-20 s; an exchange takes microseconds).  Ranks that drift further apart than this -- a
// checkpoint on rank 0, a stalled data loader -- should raise the bound or pass 0 = wait for ever.
static std::atomic<long long> g_max_spins{1LL << 24};
long long xchg_max_spins() { return g_max_spins.load(); }

}  // namespace simt

using namespace simt;

extern "C" {

void simt_xchg_set_timeout(long long max_spins) { g_max_spins.store(max_spins); }

size_t simt_xchg_bytes(int C) {
  if (C <= 0) return 0;
  const size_t slot = 2 + (size_t)C * kXchgMaxCKP;
  return kHdrBytes + kCountBytes + (size_t)2 * kMaxPeers * (2 * slot) * sizeof(unsigned long long);
}

int simt_xchg_create(size_t bytes, void** mailbox, unsigned char* handle64) {
  if (!mailbox || !handle64 || bytes < kHdrBytes + kCountBytes) return SIMT_EINVAL;
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

}  // extern "C"
