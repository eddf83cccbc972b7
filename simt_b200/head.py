"""The fused SimT head as a torch.autograd.Function over the C ABI.

Replaces, per head, the reference lines tools/trainV2_simt.py:371-372 (upsample),
:402-403 / :405-406 (identity upsample, softmax, NHWC flatten, mm with T) and
:408-409 (``CrossEntropy2d(is_softmax=False)``, utils/loss.py:14-40), plus their
autograd backward (:428):

    loss_y = simt_head(pred_lo, T, label_target, (H, W))

PyTorch is used for device memory, streams and torch.distributed only; all
arithmetic is in libsimt_b200.so.  There is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

_WORKSPACES = {}
_ERRFLAGS = {}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.index, _stream_ptr())
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def error_flag(dev) -> torch.Tensor:
    """Per-device int32 word the kernels OR their SIMT_ERRBIT_* bits into."""
    dev = torch.device(dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    f = _ERRFLAGS.get(idx)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", idx))
        _ERRFLAGS[idx] = f
    return f


def check_errors(dev=None) -> None:
    """Synchronise and raise if a kernel flagged a contract violation (the reference raises
    IndexError / a device-side assert for a label in [C, 254], and numpy raises ValueError for an
    out-of-table prediction).  Clears the flag."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else torch.device(dev)
    f = error_flag(dev)
    v = int(f.item())
    if v:
        f.zero_()
        what = []
        if v & 1:
            what.append("label outside [0, C) that is not the ignore label")
        if v & 2:
            what.append("n_cols*a+b outside the histogram table")
        if v & 8:
            what.append("sharded step: the labels differ from the next_labels announced one step earlier")
        if v & 4:
            what.append("sharded step: a peer rank's count / stats never arrived (bounded wait expired; this rank's "
                        "loss, dT and possibly dlogits of that step are NaN)")
        raise IndexError("simt_b200: " + "; ".join(what))


def _check_inputs(logits, T, labels, out_size):
    if not (isinstance(logits, torch.Tensor) and logits.is_cuda):
        raise RuntimeError("simt_b200 runs on CUDA (sm_100a) only: logits must be a CUDA tensor; there is no CPU fallback")
    if logits.dtype != torch.float32 or logits.dim() != 4:
        raise TypeError("logits must be float32 [B, CK, h, w]")
    if labels.dim() != 3 or labels.size(0) != logits.size(0):
        raise ValueError(f"labels must be [B, H, W] with B={logits.size(0)}, got {tuple(labels.shape)}")
    if labels.dtype not in (torch.uint8, torch.int64):
        raise TypeError("labels must be uint8 (fast path) or int64 (the reference's dtype)")
    if labels.device != logits.device:
        raise RuntimeError("labels and logits must be on the same device")
    H, W = int(out_size[0]), int(out_size[1])
    if (labels.size(1), labels.size(2)) != (H, W):
        raise ValueError(f"{labels.size(1)}x{labels.size(2)} labels vs out_size {H}x{W}")
    CK = logits.size(1)
    if T is not None:
        if T.dim() != 2 or T.size(0) != CK:
            raise ValueError(f"T must be [CK={CK}, C], got {tuple(T.shape)}")
        if T.dtype != torch.float32 or T.device != logits.device:
            raise TypeError("T must be float32 on the logits' device")
    return H, W


def head_forward_raw(logits, T, labels, out_size, ignore=255, need_grad=True):
    """One launch of the fused kernel.  Returns (stats f64[2+CK*C], loss_mean f32[], dlogits_raw or None)."""
    lib = _lib.load()
    H, W = _check_inputs(logits, T, labels, out_size)
    logits = logits.contiguous()
    labels = labels.contiguous()
    Tc = None if T is None else T.detach().contiguous()
    B, CK, h, w = logits.shape
    C = CK if Tc is None else Tc.size(1)
    dev = logits.device
    with torch.cuda.device(dev):             # the workspace is sized from this device's SM count
        nws = lib.simt_head_workspace_bytes(B, CK, C, h, w, H, W)
        ws = _workspace(dev, nws)
    stats = torch.empty(2 + CK * C, dtype=torch.float64, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    err = error_flag(dev)
    lb = 1 if labels.dtype == torch.uint8 else 8
    tptr = None if Tc is None else Tc.data_ptr()
    with torch.cuda.device(dev):
        if need_grad:
            dl = torch.empty_like(logits)
            rc = lib.simt_head_fwdbwd(logits.data_ptr(), B, CK, h, w, tptr, C, labels.data_ptr(), lb, H, W, int(ignore),
                                      dl.data_ptr(), stats.data_ptr(), loss.data_ptr(), err.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _stream_ptr())
            _lib.check(rc, "simt_head_fwdbwd")
        else:
            dl = None
            rc = lib.simt_head_fwd(logits.data_ptr(), B, CK, h, w, tptr, C, labels.data_ptr(), lb, H, W, int(ignore),
                                   stats.data_ptr(), loss.data_ptr(), err.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _stream_ptr())
            _lib.check(rc, "simt_head_fwd")
    return stats, loss, dl


def head_scale(dl_raw, stats, CK, C, grad_out, want_dT=True):
    """dlogits = raw * grad_out / n_valid (in place), dT = stats[2:] * grad_out / n_valid."""
    lib = _lib.load()
    dev = dl_raw.device
    dT = torch.empty(CK, C, dtype=torch.float32, device=dev) if want_dT else None
    g = None
    if grad_out is not None:
        g = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        rc = lib.simt_head_scale(dl_raw.data_ptr(), dl_raw.numel(), stats.data_ptr(), CK, C,
                                 None if g is None else g.data_ptr(), None if dT is None else dT.data_ptr(),
                                 _stream_ptr())
    _lib.check(rc, "simt_head_scale")
    return dl_raw, dT


class _SimTHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, T, labels, out_size, ignore, group):
        need_grad = logits.requires_grad or (T is not None and T.requires_grad)
        stats, loss, dl = head_forward_raw(logits.detach(), T, labels, out_size, ignore, need_grad)
        if group is not None:
            # batch-sharded run: ONE all-reduce(sum) of {loss_sum, n_valid, raw dT} (2.9 KB at CK=C=19)
            import torch.distributed as dist
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
            loss = (stats[0] / stats[1]).to(torch.float32)
        ctx.has_T = T is not None
        ctx.shape_T = None if T is None else tuple(T.shape)
        ctx.consumed = False
        ctx.save_for_backward(stats, dl if dl is not None else torch.empty(0, device=logits.device))
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        stats, dl = ctx.saved_tensors
        if dl.numel() == 0:
            return None, None, None, None, None, None
        if ctx.consumed:
            raise RuntimeError("simt_head: backward through the fused head a second time "
                               "(its raw gradient buffer is scaled in place)")
        ctx.consumed = True
        CK = dl.size(1)
        C = ctx.shape_T[1] if ctx.has_T else CK
        dlog, dT = head_scale(dl, stats, CK, C, grad_out, want_dT=ctx.has_T)
        return dlog, dT, None, None, None, None


def simt_head(logits_lo: torch.Tensor, T: Optional[torch.Tensor], labels: torch.Tensor,
              out_size: Optional[Tuple[int, int]] = None, ignore: int = 255, group=None) -> torch.Tensor:
    """T-corrected per-pixel loss of one DeepLab head, fused.

    Equals ``CrossEntropy2d(is_softmax=False)(mm(softmax(interp(interp(logits_lo))), T), labels)``
    of the reference (tools/trainV2_simt.py:371-372,402-409) to 1e-5 relative; gradients flow to
    ``logits_lo`` (low resolution) and ``T``.  ``T=None`` gives plain CE on the upsampled logits
    (:394-395).  ``labels`` uint8 or int64 [B, H, W]; ``group``: a torch.distributed process group
    for a batch-sharded run (mean over the GLOBAL valid-pixel count, dT summed over ranks).
    """
    if out_size is None:
        out_size = (labels.size(1), labels.size(2))
    return _SimTHeadFn.apply(logits_lo, T, labels, tuple(out_size), int(ignore), group)


class SimTHead(torch.nn.Module):
    """Module form: ``SimTHead((H, W))(logits_lo, T, labels)``; holds no parameters."""

    def __init__(self, out_size=None, ignore_label: int = 255, group=None):
        super().__init__()
        self.out_size = out_size
        self.ignore_label = ignore_label
        self.group = group

    def forward(self, logits_lo, T, labels):
        return simt_head(logits_lo, T, labels, self.out_size, self.ignore_label, self.group)


class HeadRunner:
    """Static-shape, allocation-free form of the fused head for training loops and benchmarks.

    All outputs are preallocated once; ``step`` enqueues a label-count + zeroing pass, the fused fwd/bwd kernel (which
    applies grad_out / N_valid itself) and finalize.  Sharded (``group``) the same three kernels also exchange the
    valid counts and the 2.9 KB stats buffer with their peers over CUDA-IPC peer memory (``simt_head_step_sharded``),
    so the mean is over the GLOBAL batch and dT is summed over ranks.  It returns views
    (loss f32[], dlogits f32[B,CK,h,w], dT f32[CK,C]) without synchronising.  ``graph_step`` is the same work
    replayed from a CUDA graph.  ``close()`` releases the peer mailboxes (also a context manager).
    """

    def __init__(self, B, CK, C, h, w, H, W, device=None, ignore=255, label_dtype=torch.uint8, group=None,
                 exchange="p2p"):
        self.lib = _lib.load()
        self.shape = (int(B), int(CK), int(C), int(h), int(w), int(H), int(W))
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.ignore = int(ignore)
        self.label_dtype = label_dtype
        self.label_bytes = 1 if label_dtype == torch.uint8 else 8
        self.group = group
        B, CK, C, h, w, H, W = self.shape
        with torch.cuda.device(self.dev):       # the workspace is sized from this device's SM count
            nws = self.lib.simt_head_workspace_bytes(B, CK, C, h, w, H, W)
        self.ws = torch.zeros(nws, dtype=torch.uint8, device=self.dev)
        self.stats = torch.zeros(2 + CK * C, dtype=torch.float64, device=self.dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.dlogits = torch.zeros(B, CK, h, w, dtype=torch.float32, device=self.dev)
        self.dT = torch.zeros(CK, C, dtype=torch.float32, device=self.dev)
        self.err = error_flag(self.dev)
        self._p = (self.ws.data_ptr(), self.ws.numel(), self.stats.data_ptr(), self.loss.data_ptr(),
                   self.dlogits.data_ptr(), self.dT.data_ptr(), self.err.data_ptr())
        self._graphs = {}
        self._graph_keepalive = []
        self._checked = set()
        self._deferred = False      # sharded, pipelined mode: a deferred step is outstanding
        # sharded: the exchanges are fused into the step's kernels over CUDA-IPC peer memory when every rank can map
        # every other rank's mailbox (one node); otherwise one library all-reduce per step
        self.mailbox = None
        self._world = 1
        if group is not None:
            import torch.distributed as dist
            self._world = dist.get_world_size(group)
        if group is not None and exchange == "p2p" and self._world > 1:
            from .dist import PeerMailbox
            try:
                self.mailbox = PeerMailbox(C, group, self.dev)
            except RuntimeError:
                self.mailbox = None

    # ---- lifetime ----------------------------------------------------------------------------------------------
    def close(self):
        """Release the peer mailboxes (collective in spirit: peers that keep stepping would wait for this rank)."""
        self._graphs.clear()
        self._graph_keepalive.clear()
        if self.mailbox is not None:
            self.mailbox.close()
            self.mailbox = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- argument checks (first call per buffer set; the kernels take raw pointers) -------------------------------
    def _check(self, logits, T, labels, grad_out=None):
        key = (logits.data_ptr(), None if T is None else T.data_ptr(), labels.data_ptr(),
               None if grad_out is None else grad_out.data_ptr())
        if key in self._checked:
            return
        B, CK, C, h, w, H, W = self.shape

        def bad(what):
            raise ValueError(f"HeadRunner{self.shape}: {what}")
        if tuple(logits.shape) != (B, CK, h, w) or logits.dtype != torch.float32:
            bad(f"logits must be float32 [{B}, {CK}, {h}, {w}], got {logits.dtype} {tuple(logits.shape)}")
        if tuple(labels.shape) != (B, H, W) or labels.dtype != self.label_dtype:
            bad(f"labels must be {self.label_dtype} [{B}, {H}, {W}], got {labels.dtype} {tuple(labels.shape)}")
        if T is not None and (tuple(T.shape) != (CK, C) or T.dtype != torch.float32):
            bad(f"T must be float32 [{CK}, {C}], got {T.dtype} {tuple(T.shape)}")
        if T is None and CK != C:
            bad("T=None (plain CE) needs CK == C")
        for name, t in (("logits", logits), ("labels", labels), ("T", T), ("grad_out", grad_out)):
            if t is None:
                continue
            if not t.is_contiguous():
                bad(f"{name} must be contiguous")
            if t.device != self.dev:
                bad(f"{name} is on {t.device}, the runner on {self.dev}")
        if grad_out is not None and (grad_out.numel() != 1 or grad_out.dtype != torch.float32):
            bad("grad_out must be a float32 scalar tensor")
        if len(self._checked) < 64:
            self._checked.add(key)

    def fwdbwd(self, logits, T, labels, stream=None):
        self._check(logits, T, labels)
        B, CK, C, h, w, H, W = self.shape
        ws, nws, stats, loss, dl, _, err = self._p
        rc = self.lib.simt_head_fwdbwd(logits.data_ptr(), B, CK, h, w, None if T is None else T.data_ptr(), C,
                                       labels.data_ptr(), self.label_bytes, H, W, self.ignore, dl, stats, loss, err,
                                       ws, nws, _stream_ptr() if stream is None else stream)
        if rc:
            _lib.check(rc, "simt_head_fwdbwd")

    def scale(self, grad_out=None, stream=None):
        B, CK, C, h, w, H, W = self.shape
        _, _, stats, _, dl, dT, _ = self._p
        rc = self.lib.simt_head_scale(dl, B * CK * h * w, stats, CK, C,
                                      None if grad_out is None else grad_out.data_ptr(), dT,
                                      _stream_ptr() if stream is None else stream)
        if rc:
            _lib.check(rc, "simt_head_scale")

    def step(self, logits, T, labels, grad_out=None, next_labels=None, defer=False, out=None):
        """One training step: (loss, dlogits, dT), final and global, without synchronising.
        One GPU: ``simt_head_step`` (label count + zeroing, fused kernel applying grad_out / N itself, finalize).
        Sharded: ``simt_head_step_sharded`` -- the same three kernels exchange counts and stats over the peer
        mailboxes -- or, when the peers cannot be mapped, fwdbwd + one NCCL all-reduce + scale.
        Sharded, pipelined (no rank waits inside a step; validated on 2 GPUs): ``next_labels``
        = the labels the NEXT step will get (their valid count crosses the ranks one step early); ``defer=True``
        leaves loss / dT / stats of this step to be finalised inside the fused kernel of the step after the next, or
        by ``finish()`` (a collective: every rank calls it at the same point) -- dlogits are always final.
        ``out``: a float32 [B, CK, h, w] buffer to receive dlogits instead of ``self.dlogits``."""
        self._check(logits, T, labels, grad_out)
        stream = _stream_ptr()
        B, CK, C, h, w, H, W = self.shape
        ws, nws, stats, loss, dl, dT, err = self._p
        dl_t = self.dlogits
        if out is not None:
            if tuple(out.shape) != (B, CK, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.dev:
                raise ValueError(f"HeadRunner{self.shape}: out must be a contiguous float32 [{B}, {CK}, {h}, {w}] on {self.dev}")
            dl, dl_t = out.data_ptr(), out
        if next_labels is not None and (tuple(next_labels.shape) != (B, H, W) or next_labels.dtype != self.label_dtype
                                        or not next_labels.is_contiguous() or next_labels.device != self.dev):
            raise ValueError(f"HeadRunner{self.shape}: next_labels must look like labels")
        tp = None if T is None else T.data_ptr()
        gp = None if grad_out is None else grad_out.data_ptr()
        with torch.cuda.device(self.dev):
            if self._world > 1 and self.mailbox is None:
                import torch.distributed as dist
                self.fwdbwd(logits, T, labels, stream)
                dist.all_reduce(self.stats, op=dist.ReduceOp.SUM, group=self.group)
                self.scale(grad_out, stream)
                self.loss.copy_((self.stats[0] / self.stats[1]).to(torch.float32))   # the GLOBAL mean, as on the p2p path
                if out is not None:
                    out.copy_(self.dlogits)
                return self.loss, dl_t, self.dT
            if self._world > 1:
                mb = self.mailbox
                if not defer and self._deferred:
                    # a synchronous step pushes its stats at once: every deferred step must have been reduced first
                    # (protocol invariant, tests/test_xchg_protocol_cpu.py); every rank takes this branch together
                    self.finish()
                self._deferred = bool(defer)
                self._grad_out = grad_out
                rc = self.lib.simt_head_step_sharded(logits.data_ptr(), B, CK, h, w, tp, C, labels.data_ptr(),
                                                     self.label_bytes, H, W, self.ignore, gp, dl, dT, stats, loss, err,
                                                     ws, nws, mb.rank, mb.world, mb.ptrs,
                                                     None if next_labels is None else next_labels.data_ptr(),
                                                     1 if defer else 0, stream)
                what = "simt_head_step_sharded"
            else:
                rc = self.lib.simt_head_step(logits.data_ptr(), B, CK, h, w, tp, C, labels.data_ptr(), self.label_bytes,
                                             H, W, self.ignore, gp, dl, dT, stats, loss, err, ws, nws, stream)
                what = "simt_head_step"
        if rc:
            _lib.check(rc, what)
        return self.loss, dl_t, self.dT

    def finish(self):
        """Sharded, after ``step(..., defer=True)``: finalise loss / dT / stats of every outstanding step now (one tiny
        kernel; a no-op when nothing is outstanding).  Collective in spirit: the peers' stats of the LAST step only
        leave them when they run their next step or call ``finish()`` too."""
        if self._world <= 1 or self.mailbox is None:
            return self.loss, self.dT
        B, CK, C, h, w, H, W = self.shape
        ws, nws, stats, loss, dl, dT, err = self._p
        mb = self.mailbox
        g = getattr(self, "_grad_out", None)
        with torch.cuda.device(self.dev):
            rc = self.lib.simt_head_finish_sharded(CK, C, None if g is None else g.data_ptr(), dT, stats, loss, err, ws, nws,
                                                   mb.rank, mb.world, mb.ptrs, _stream_ptr())
        if rc:
            _lib.check(rc, "simt_head_finish_sharded")
        self._deferred = False
        return self.loss, self.dT

    def graph_step(self, logits, T, labels, grad_out=None, next_labels=None, defer=False, out=None):
        """``step`` replayed from a CUDA graph: its three launches are captured ONCE for this exact set of buffers
        (keyed by their addresses) and re-launched as one graph afterwards -- they are launch-latency-bound next to a
        90 us kernel.  The graph reads the buffers' CURRENT contents on every replay (refill ``logits`` / ``labels``
        / ``T`` in place).  Sharded runs replay too when the exchange is the fused peer-memory one (no library
        collective inside the graph); with the all-reduce fallback they take the eager path."""
        if self._world > 1 and self.mailbox is None:
            return self.step(logits, T, labels, grad_out, out=out)
        if self._world > 1 and not defer and self._deferred:
            self.finish()                                    # (see step(): a synchronous step drains the deferred ones)
        key = (logits.data_ptr(), None if T is None else T.data_ptr(), labels.data_ptr(),
               None if grad_out is None else grad_out.data_ptr(),
               None if next_labels is None else next_labels.data_ptr(), bool(defer), None if out is None else out.data_ptr())
        g = self._graphs.get(key)
        if g is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("HeadRunner.graph_step: first call for a buffer set must be outside stream capture")
            # warm-up (one-time attribute / occupancy queries): a plain synchronous step -- it must not announce
            # next_labels, because the replay below runs THESE labels once more
            self.step(logits, T, labels, grad_out, None, False, out)
            torch.cuda.current_stream(self.dev).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.step(logits, T, labels, grad_out, next_labels, defer, out)
            self._graphs[key] = g
            self._graph_keepalive.append((logits, T, labels, grad_out, next_labels, out))
        g.replay()
        if self._world > 1:
            self._deferred = bool(defer)
        return self.loss, (self.dlogits if out is None else out), self.dT

    def error_word(self) -> torch.Tensor:
        """The device's SIMT_ERRBIT_* word (int32[1], no synchronisation): copy it to the host alongside the loss, or call
        ``check()`` at a point where a synchronisation is acceptable.  A sharded step whose peer never arrived has
        already poisoned its loss / dT / dlogits with NaN when the bit is set."""
        return self.err

    def check(self) -> None:
        """Synchronise and raise if a kernel flagged a contract violation or an exchange timeout (``check_errors``)."""
        check_errors(self.dev)

    def global_loss(self):
        """loss over the all-reduced stats (sharded runs); a 0-dim f64 tensor, no sync."""
        return self.stats[0] / self.stats[1]


class HostPrefetcher:
    """Double-buffered pinned-host -> device uploader for (logits, labels) batches.

    A batch travels as ONE packed buffer (logits bytes, then label bytes): ``host_buffer()`` hands out pinned staging
    buffers with ``.logits`` / ``.labels`` views to fill, ``submit(buf)`` enqueues ONE copy of the next batch on a side
    stream, ``get()`` makes the compute stream wait for it and returns the device views.  While a step computes, the
    next step's 9 MB of inputs cross PCIe -- the usual input pipeline of a training loop.
    """

    class _Packed:
        def __init__(self, raw, B, CK, h, w, H, W, label_dtype):
            nl = B * CK * h * w * 4
            off = (nl + 15) // 16 * 16           # labels start 16-byte aligned
            self.raw = raw
            self.logits = raw[:nl].view(torch.float32).view(B, CK, h, w)
            self.labels = raw[off:].view(label_dtype).view(B, H, W)

    def __init__(self, B, CK, h, w, H, W, device=None, label_dtype=torch.uint8):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (B, CK, h, w, H, W, label_dtype)
        self.nbytes = (B * CK * h * w * 4 + 15) // 16 * 16 + B * H * W * torch.empty((), dtype=label_dtype).element_size()
        self.stream = torch.cuda.Stream(device=self.dev)
        self.bufs = [self._Packed(torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev), *self.shape)
                     for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot = 0
        self.pending = None

    def host_buffer(self):
        """A pinned staging buffer (fill ``.logits`` / ``.labels``, then ``submit`` it)."""
        return self._Packed(torch.empty(self.nbytes, dtype=torch.uint8).pin_memory(), *self.shape)

    def submit(self, host_buf) -> None:
        slot = self.slot
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])          # the compute that last used this slot is done
            self.bufs[slot].raw.copy_(host_buf.raw, non_blocking=True)
            self.events[slot].record(self.stream)
        self.pending = slot
        self.slot ^= 1

    def get(self):
        slot = self.pending
        torch.cuda.current_stream(self.dev).wait_event(self.events[slot])
        return self.bufs[slot].logits, self.bufs[slot].labels

    def release(self, slot_tensors) -> None:
        """Call after the step that consumed ``slot_tensors`` has been enqueued."""
        for i, b in enumerate(self.bufs):
            if b.logits is slot_tensors[0]:
                self.free[i].record(torch.cuda.current_stream(self.dev))
