/*
 * simt_b200.h -- C ABI of libsimt_b200.so: the B200 (sm_100a) implementation of
 * SimT's per-pixel training/eval head.
 *
 * The reference (CityU-AIM-Group/SimT) has no FFI of its own: the hot path is a
 * sequence of PyTorch / numpy calls in Python.  Each entry point below replaces
 * the reference lines cited next to it (paths relative to the reference root);
 * the Python host mirror in simt_b200/ binds them with ctypes and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller
 *     unless the name ends in _host;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and the call
 *     returns without synchronising; re-entrant across streams as long as each
 *     in-flight call has its own workspace;
 *   - return value: 0 on success, a positive cudaError_t from the runtime, or a
 *     negative SIMT_E* argument-validation code.  No C++ exceptions cross the
 *     boundary.  simt_b200_strerror() maps any code to text;
 *   - data-dependent contract violations (a label in [C, 254], a prediction
 *     outside [0, n)) cannot be returned synchronously: they set bits in the
 *     caller's `err_flag` word on the device (SIMT_ERRBIT_*), the offending
 *     pixel is skipped, and the head additionally poisons its loss with NaN.
 */
#ifndef SIMT_B200_H_
#define SIMT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMT_B200_ABI_VERSION 1

enum {
  SIMT_EINVAL      = -1,  /* bad argument (null pointer, non-positive size, ...) */
  SIMT_EUNSUPPORTED = -2, /* shape outside what the kernels are built for      */
  SIMT_EWORKSPACE  = -3,  /* workspace too small                                */
  SIMT_ENOSMEM     = -4   /* tile does not fit in shared memory                 */
};

#define SIMT_ERRBIT_LABEL_RANGE 1 /* head: label not ignore and not in [0, C)        */
#define SIMT_ERRBIT_PRED_RANGE  2 /* histograms: n_cols*a+b outside [0, rows*cols)   */
#define SIMT_ERRBIT_XCHG_TIMEOUT 4 /* sharded step: a peer's count / stats never arrived  */
#define SIMT_ERRBIT_NEXT_LABELS  8 /* sharded step: labels differ from the next_labels announced one step earlier */

int         simt_b200_abi_version(void);
const char* simt_b200_strerror(int code);

/* Measurement hook (bench.py): while enabled, every head / histogram entry point records a pair
 * of CUDA events on the caller's stream right around its dominant kernel.
 * simt_b200_profile_read synchronises on them, returns the summed device time and the number
 * of launches since the last read, and resets the counter.  Not thread-safe. */
void simt_b200_profile_enable(int on);
int  simt_b200_profile_read(double* total_ms_host, long long* launches_host);

/* ------------------------------------------------------------------------- *
 * The fused head.  Replaces, per head, tools/trainV2_simt.py:371-372 (bilinear
 * upsample, align_corners=True, built at :301), :402/:405 (same-size upsample
 * [identity], channel softmax, NHWC flatten), :403/:406 (torch.mm with T),
 * :408-409 (CrossEntropy2d(is_softmax=False), utils/loss.py:14-40) and the
 * autograd backward of those lines triggered at :428.
 *
 *   logits  [B, CK, h, w] f32 contiguous (DeepLab classifier output, CK = C + K)
 *   T       [CK, C] f32 row-major (model/deeplab_multi.py:259-263), or NULL for
 *           the identity (then C must equal CK): plain CE on the upsampled
 *           logits, i.e. seg_loss at trainV2_simt.py:394-395 /
 *           trainV1_warmup.py:222-224 and CrossEntropy2d(is_softmax=True)
 *   labels  [B, H, W]; label_bytes = 1 (uint8, fast path) or 8 (int64, the
 *           dtype the reference passes, trainV2_simt.py:348).  Valid pixel:
 *           label >= 0 and label != ignore (utils/loss.py:29)
 *   stats   [2 + CK*C] f64:  stats[0] = sum over valid pixels of -log q_y,
 *           stats[1] = number of valid pixels, stats[2 + k*C + c] = UNNORMALISED
 *           dT[k][c] = -sum_{valid, y=c} p_k / q_y  (zero-filled for simt_head_fwd).
 *           One f64 buffer so that a sharded run needs ONE all-reduce(sum).
 *   loss_mean [1] f32: stats[0] / stats[1] (NaN if no valid pixel, as the reference;
 *           NaN if a label was out of range)
 *   workspace: simt_head_workspace_bytes() bytes, zero-filled ONCE by the caller
 *           before first use; the library leaves it zeroed after every call.
 * ------------------------------------------------------------------------- */
size_t simt_head_workspace_bytes(int B, int CK, int C, int h, int w, int H, int W);

/* forward only (eval / no-grad) */
int simt_head_fwd(const float* logits, int B, int CK, int h, int w,
                  const float* T, int C,
                  const void* labels, int label_bytes, int H, int W, int ignore,
                  double* stats, float* loss_mean, int* err_flag,
                  void* workspace, size_t workspace_bytes, void* stream);

/* single pass forward + backward.  dlogits_raw [B, CK, h, w] receives the
 * UNNORMALISED gradient sum_pixels U^T (p_k - p_k T[k,y]/q_y) (U = the bilinear
 * operator, transposed in-kernel); it is zeroed by the call.  Multiply by
 * grad_out / n_valid (simt_head_scale) once n_valid is known globally. */
int simt_head_fwdbwd(const float* logits, int B, int CK, int h, int w,
                     const float* T, int C,
                     const void* labels, int label_bytes, int H, int W, int ignore,
                     float* dlogits_raw, double* stats, float* loss_mean, int* err_flag,
                     void* workspace, size_t workspace_bytes, void* stream);

/* two-pass backward with a scale already known on the host:
 * dlogits = scale * raw, dT [CK, C] f32 = scale * raw dT.  (scale = grad_out / n_valid) */
int simt_head_bwd(const float* logits, int B, int CK, int h, int w,
                  const float* T, int C,
                  const void* labels, int label_bytes, int H, int W, int ignore,
                  float scale, float* dlogits, float* dT, int* err_flag,
                  void* workspace, size_t workspace_bytes, void* stream);

/* epilogue of the single-pass variant, no host sync:
 *   s = (grad_out ? *grad_out : 1) / stats[1];  dlogits[i] *= s;  dT[k][c] = s * stats[2 + k*C + c]
 * `stats` may be the all-reduced buffer of a sharded run. */
int simt_head_scale(float* dlogits, long long n_dlogits, const double* stats, int CK, int C,
                    const float* grad_out, float* dT, void* stream);

/* tuning hook for benchmarks (process-global, mutex-guarded like the library's launch caches: the entry points may be
 * called from several host threads, one stream / device each): cell-rows per work unit, row slices per cell-row,
 * reserved, lanes per pixel-run (2 or 4); 0 = automatic. */
void simt_head_set_tuning(int tile_cells_y, int tile_cells_x, int threads, int lanes_per_run);

/* ------------------------------------------------------------------------- *
 * CrossEntropy2d(is_softmax=False) on already-mixed probabilities, for callers
 * that keep the reference's unfused lines: utils/loss.py:14-40 forward
 * (-mean log prob[b, y, i, j] over valid pixels) and its backward.
 *   prob [B, C, H, W] f32;  stats [2] f64 = {sum -log, n_valid}
 *   simt_nll2d_bwd: dprob (zeroed by the call) [b, y, i, j] = -(grad_out / n_valid) / prob
 * ------------------------------------------------------------------------- */
int simt_nll2d_fwd(const float* prob, int B, int C, int H, int W,
                   const void* labels, int label_bytes, int ignore,
                   double* stats, float* loss_mean, int* err_flag, void* stream);
int simt_nll2d_bwd(const float* prob, int B, int C, int H, int W,
                   const void* labels, int label_bytes, int ignore,
                   const double* stats, const float* grad_out, float* dprob, void* stream);

/* ------------------------------------------------------------------------- *
 * Integer eval histograms (bit-exact).
 *
 * simt_confusion: hist[n_cols*a' + b] += 1 for every pixel with 0 <= a' < n_rows,
 * where a' = lut256 ? lut256[a] : a.  Replaces tools/compute_iou.py:9-11
 * fast_hist(a, b, n) (n_rows = n_cols = n; dup tools/evaluate_cityscapes.py:81-83),
 * tools/compute_ConfusionMatrix.py:54-56 fast_hist(a, b, n33, n19), and -- through
 * the LUT -- tools/compute_iou.py:18-22 label_mapping fused in front of it.  Like
 * numpy.bincount on n*a+b, an out-of-range b that still lands inside the table is
 * counted where it lands; an index outside the table (numpy raises) sets
 * SIMT_ERRBIT_PRED_RANGE and is skipped.
 *   a, b: [n] uint8 (a_bytes/b_bytes = 1, fast path) or int64 (= 8)
 *   lut256: 256-byte table on the device or NULL (only with a_bytes = 1)
 *   hist: [n_rows*n_cols] int64, ACCUMULATED into (the reference's `hist +=`,
 *   compute_iou.py:51)
 *
 * simt_class_hist: hist[a] += 1 for 0 <= a < n_bins.  Replaces
 * tools/compute_ClassDistribution.py:52-54 fast_hist(a, n).
 *
 * simt_label_map: out[i] = lut256[in[i]] as int64 (compute_iou.py:18-22).
 * ------------------------------------------------------------------------- */
int simt_confusion(const void* a, int a_bytes, const void* b, int b_bytes, long long n,
                   const uint8_t* lut256, int n_rows, int n_cols,
                   long long* hist, int* err_flag, void* stream);
int simt_class_hist(const void* a, int a_bytes, long long n, int n_bins,
                    long long* hist, void* stream);
int simt_label_map(const uint8_t* in, long long n, const uint8_t* lut256, long long* out, void* stream);

/* ------------------------------------------------------------------------- *
 * T regularisers (tools/trainV2_simt.py:412-421) and anchor statistics (:375-384).
 *
 * simt_t_regularizers: ONE single-CTA launch for one head.
 *   T [CK, C], W [CK, CK] f32 (W may be NULL: volume only)
 *   out2[0] = -||W T||_F^2          (:414-415, this head's term of NTM_Convex_loss)
 *   out2[1] = 0.5 log|det(T^T T)|   (:417-418; 0 when not finite, the guard of :420-421)
 *   dT_convex = -2 W^T W T, dW_convex = -2 W T T^T, dT_volume = T (T^T T)^-1 (zeros when not finite)
 *
 * simt_anchor_stats: Anchor_index / Exist_label of :375-377 from the LOW-res logits [B, CK, h, w]:
 *   anchor_idx[k] = arg-max over the B*H*W upsampled pixels (NHWC-flatten order) of channel k,
 *   anchor_val[k] (optional) the maximum, exist_mask bit k = class k is the arg-max at some pixel
 *   (CK <= 64).  scratch: CK u64 words.  Smallest pixel / class wins ties.
 *
 * simt_bilinear_gather: rows[r][c] = upsample(src)[b, c, Y, X] at flat pixel pixel_idx[r]
 *   (labelC_flat[Anchor_index], :378) without materialising the upsampled tensor.
 * ------------------------------------------------------------------------- */
int simt_t_regularizers(const float* T, const float* W, int CK, int C, float* out2,
                        float* dT_convex, float* dT_volume, float* dW_convex, void* stream);
int simt_anchor_stats(const float* logits, int B, int CK, int h, int w, int H, int W,
                      long long* anchor_idx, float* anchor_val, unsigned long long* exist_mask,
                      unsigned long long* scratch, void* stream);
int simt_bilinear_gather(const float* src, int B, int C, int h, int w, int H, int W,
                         const long long* pixel_idx, int n, float* rows, void* stream);

/* ------------------------------------------------------------------------- *
 * Sharded step (one process per GPU, batch split over the ranks of ONE node, world <= 8).  The
 * reference is single-process: its loss.backward() (tools/trainV2_simt.py:408-409,428) sees the
 * whole batch, so the mean is over the GLOBAL valid-pixel count and dT is summed over ranks.  The
 * two tiny exchanges of a step (8-byte valid counts; the 2 + CK*C doubles of `stats`) travel over
 * peer memory (NVLink / NVSwitch P2P stores into CUDA-IPC mapped mailboxes) from inside the step's
 * own kernels -- no library collective, no extra launch, no pass over dLogits.
 *
 * simt_xchg_create: cudaMalloc + zero a mailbox of simt_xchg_bytes(C) bytes on the current
 *   device and export its 64-byte CUDA IPC handle.  The caller exchanges the handles (any
 *   transport: torch.distributed all-gather, MPI, a pipe) and opens every peer's mailbox with
 *   simt_xchg_open (current device = the opener's GPU).  close / destroy undo open / create.
 * simt_head_step_sharded: simt_head_step (below) on this rank's batch shard, with
 *   dlogits = raw * g / N_global, dT = sum_ranks raw dT * g / N_global, loss_mean = the global mean
 *   and `stats` = the all-reduced buffer (summed in rank order: the same bits on every rank).
 *     1. head_prep_kernel zeroes dlogits, counts this rank's valid labels and pushes the count into
 *        every peer's mailbox;
 *     2. the fused kernel acquires the `world` counts right before its first dLogits update (a
 *        late peer costs nothing until then) and applies g / N_global itself;
 *     3. finalize pushes this rank's stats, waits for the peers' and writes loss / dT / stats.
 *   mailboxes: HOST array [world] of device pointers, mailboxes[rank] = this rank's own.
 *   Every rank must call it once per step, in the same order.  No host-side step argument: the
 *   call sequence can be captured in a CUDA graph.
 *   Pipelined form (no rank ever waits inside a step): `next_labels` (same dtype / shape as labels, or
 *   NULL) are the labels the NEXT call will be given -- their count is exchanged now, one step early
 *   (a next call with other labels sets SIMT_ERRBIT_NEXT_LABELS); `defer` != 0 only pushes this
 *   step's stats: loss_mean / dT / stats become final in the prologue of the next call on the same
 *   workspace + mailboxes, or in simt_head_finish_sharded (same output pointers), by which time the
 *   peers' words have long arrived.  dlogits are always final on return (stream order).
 * simt_xchg_set_timeout: bound of every in-kernel wait, in polls of ~1 us (default 2^24, of the
 *   order of 10-20 s; <= 0 waits for ever).  A peer that does not arrive in time poisons this rank's
 *   loss_mean, dT, stats (and, when the count is missing, dlogits) with NaN and sets
 *   SIMT_ERRBIT_XCHG_TIMEOUT in err_flag: a rank never continues with a partial sum.  Ranks must
 *   not drift further apart than the bound (raise it around checkpoints / evaluation on one rank).
 * ------------------------------------------------------------------------- */
size_t simt_xchg_bytes(int C);
int simt_xchg_create(size_t bytes, void** mailbox, unsigned char* handle64);
int simt_xchg_open(const unsigned char* handle64, void** peer_mailbox);
int simt_xchg_close(void* peer_mailbox);
int simt_xchg_destroy(void* mailbox);
void simt_xchg_set_timeout(long long max_spins);
int simt_head_step_sharded(const float* logits, int B, int CK, int h, int w, const float* T, int C,
                           const void* labels, int label_bytes, int H, int W, int ignore,
                           const float* grad_out, float* dlogits, float* dT, double* stats, float* loss_mean,
                           int* err_flag, void* workspace, size_t workspace_bytes,
                           int rank, int world, void* const* mailboxes,
                           const void* next_labels, int defer, void* stream);
int simt_head_finish_sharded(int CK, int C, const float* grad_out, float* dT, double* stats, float* loss_mean,
                             int* err_flag, void* workspace, size_t workspace_bytes,
                             int rank, int world, void* const* mailboxes, void* stream);

/* ------------------------------------------------------------------------- *
 * simt_head_step: one whole training step of the head on ONE GPU, the form a training loop calls every
 * iteration (the same reference lines as simt_head_fwdbwd + simt_head_scale, tools/trainV2_simt.py:
 * 371-372,402-409,428), with the 1/N_valid scale resolved ON THE DEVICE BEFORE the fused kernel runs:
 *   1. head_prep_kernel zeroes dlogits and counts the valid labels in one pass;
 *   2. the fused kernel applies grad_out / N_valid to every dLogits contribution itself;
 *   3. finalize writes loss, stats and the scaled dT.
 * Three launches, no pass over dLogits after the kernel, no host sync, CUDA-graph capturable.
 * grad_out: device scalar or NULL (= 1).  dlogits / dT / loss_mean are FINAL on return (stream order).
 * (Batch-sharded runs: simt_head_step_sharded above.)
 * ------------------------------------------------------------------------- */
int simt_head_step(const float* logits, int B, int CK, int h, int w, const float* T, int C,
                   const void* labels, int label_bytes, int H, int W, int ignore,
                   const float* grad_out, float* dlogits, float* dT, double* stats, float* loss_mean,
                   int* err_flag, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * Placeholder_loss (tools/trainV2_simt.py:202-230, called at :398-399 on the
 * upsampled prediction of :371-372), fused with the bilinear upsample and its backward like the head:
 * the [B, CK, H, W] tensor, its one-hot, `predict`, `predict_open` and their gradients are never
 * materialised.  Per upsampled pixel, with a = arg-max channel (first on ties):
 *   valid   iff a < C (a known class) and (thres < 0 or max softmax prob > thres)            (:213-216)
 *   known   = -log softmax(z)_a                                                                (:217)
 *   z'      = z with z'_a replaced by the constant 0 (`ones` at :208 is zeros_like)            (:206-209)
 *   y       = first best open-set channel (k >= C) if its logit is > 0, else class 0           (:220-222)
 *   unknown = -log softmax(z')_y                                                               (:229)
 *   loss    = mean_valid(known) + lambda_place * mean_valid(unknown)                           (:230)
 *   logits [B, CK, h, w] f32 LOW-res; C = num_classes, CK = num_classes + open_classes
 *   dlogits_raw [B, CK, h, w]: SUM over valid pixels of dLoss_pixel/dlogits (not yet divided by the count)
 *   stats f64[2]: {loss sum, valid count}; loss_mean f32 (NaN when no pixel is valid, like the reference)
 *   finish with simt_head_scale(dlogits_raw, n, stats, 0, 0, grad_out, NULL).
 *   workspace: simt_head_workspace_bytes(...) bytes, zero-filled once (shared with the head calls).
 * ------------------------------------------------------------------------- */
int simt_placeholder_fwdbwd(const float* logits, int B, int CK, int h, int w, int C, int H, int W,
                            float thres, float lambda_place, float* dlogits_raw, double* stats,
                            float* loss_mean, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * Inner W optimisation (tools/trainV2_simt.py:326-339), one head, ONE single-CTA launch.
 * Runs n_steps rounds of: W = softmax(weight with diag := -1e4, dim 1) - I
 * (model/deeplab_multi.py:277-286), loss = sum((W T)^2) (:336), backward (:337), and a
 * torch.optim.Adam step (no weight decay, no amsgrad) on `weight` (:338-339).
 *   weight, exp_avg, exp_avg_sq [CK, CK] f32: updated IN PLACE (the optimiser's state tensors)
 *   T [CK, C] f32: constant during the loop
 *   step0: Adam steps already taken on `weight` (state['step']); lr/beta1/beta2/eps: the param group's
 *   dT_accum [CK, C] f32 or NULL: += sum over the rounds of dLoss/dT (the reference's backward at
 *     :337 accumulates into NTM.grad ten times before optimizer_t.step() at :432)
 *   losses [n_steps] f32 or NULL: the objective at the start of each round
 * ------------------------------------------------------------------------- */
int simt_w_fit(float* weight, float* exp_avg, float* exp_avg_sq, const float* T, int CK, int C,
               int n_steps, long long step0, double lr, double beta1, double beta2, double eps,
               float* dT_accum, float* losses, void* stream);

/* ------------------------------------------------------------------------- *
 * Pseudo-label generation.  Replaces tools/trainV2_simt.py:354-365 (softmax of the
 * frozen model's output2, upsample, max / arg-max, high / low confidence thresholds, and the
 * .cpu().numpy() round trip of :362) and the class-posterior relabel of :387-393, producing the uint8
 * label map the fused head consumes directly.
 *   fixed_logits_lo [B, C, h, w] f32; pred2_lo [B, CK, h, w] f32 (student head 2, LOW-res) or NULL
 *   label = arg-max class if max prob > thres_high; 255 if in between; if max prob < thres_low: the
 *   student's arg-max class when it is an open-set class (>= C), else 255 (without pred2_lo: C itself)
 *   probs_scratch: B*C*h*w floats.  labels_out [B, H, W] uint8.
 * ------------------------------------------------------------------------- */
int simt_pseudo_labels(const float* fixed_logits_lo, const float* pred2_lo, int B, int C, int CK, int h, int w,
                       int H, int W, float thres_high, float thres_low, float* probs_scratch,
                       uint8_t* labels_out, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused eval prediction map.  Replaces tools/evaluate_cityscapes.py:127-138
 * (evaluate_simt: upsample the closed-set channels of output2 of the 1024x512 pass and of the 1280x640
 * pass to the label size, two 159 MB device->host copies per image, add, np.argmax) and :186-196
 * (evaluate_warmup, one scale: logits_b = NULL).  pred_out [B, H, W] uint8 feeds simt_confusion.
 *   logits_a [B, CKa, ha, wa], logits_b [B, CKb, hb, wb] or NULL; the first C channels are used.
 * ------------------------------------------------------------------------- */
int simt_eval_argmax(const float* logits_a, int CKa, int ha, int wa, const float* logits_b, int CKb, int hb, int wb,
                     int B, int C, int H, int W, uint8_t* pred_out, void* stream);

/* benchmark hook (process-global, mutex-guarded): warps per CTA and 128-bit loads in flight per lane; 0 = automatic.
 * mode 0 = normal.  mode 9 = LOAD-ONLY PROBE for bandwidth measurements: the kernels stream their inputs with the
 * production access pattern but skip the counting, so the histogram they return is meaningless -- never leave it on. */
void simt_hist_set_tuning(int mode, int warps_per_cta, int unroll);

/* test hook, HOST ONLY (no GPU needed): the pixel -> cell tables the kernels derive for one axis of the
 * align_corners=True bilinear resize in -> out (torch ATen/native/UpSample.h:271-296,442-476), from
 * the same host/device functions the kernels call (csrc/common.cuh).  Pixel X reads sources cell[X] and
 * min(cell[X] + 1, in - 1) with weights 1 - lambda[X] and lambda[X]; first_px[c] (n_cells + 1 entries, n_cells =
 * max(in - 1, 1)) is the first pixel of cell c.  tests/test_index_math_cpu.py compares them with torch's own weights. */
int simt_debug_resize_tables(int in, int out, int* cell, float* lambda, int* first_px);

#ifdef __cplusplus
}
#endif
#endif /* SIMT_B200_H_ */
