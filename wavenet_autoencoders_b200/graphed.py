"""Inference forward replayed as one CUDA graph (additive API; the reference has no counterpart).

A teacher-forced forward of the VQ-WAE is ~75 launches (11 encoder layers, the VQ search, 3 upsampler stages, the prep kernels,
20 residual layers, the head, the NLL).  Launched eagerly from Python after a host synchronisation -- which every evaluation
step has, to read its loss -- the short frame-rate kernels at the front are issued more slowly than they execute and the GPU
idles between them.  Shapes are static per model and batch, so the whole forward is captured once and replayed; inputs are
copied into the captured buffers (from pinned host memory or from the device).  Same kernels, same results as the eager call.
Weights are read through the packed copies made at capture time: re-capture after changing parameters.
"""
from __future__ import annotations

import torch

from . import _lib
from .losses import teacher_forced_nll


class GraphedForward:
    """``GraphedForward(model, idx, mfcc, g)`` captures ``model(idx, mfcc, g)`` (+ the teacher-forced NLL of vqwae_train.py:760-766)
    for the shapes of the example inputs; calling it with new inputs of the same shapes copies them in, replays the graph and
    returns ``(logits, vq_loss, perplexity, nll)`` -- tensors owned by the graph, overwritten by the next call.  With
    ``with_logits=False`` (loss-only evaluation) the NLL is taken from the head kernel's accumulator, the logits are never
    materialised and ``logits`` is None.

    ``model`` is a ``vqvae_model.VQVAE`` (or any module with the same ``forward(x, c, g)``) in eval mode on a CUDA sm_100 device;
    ``idx`` are the (B,T) int64 mu-law classes (or the (B,O,T) one-hot tensor), ``mfcc`` the encoder input, ``g`` speaker ids."""

    def __init__(self, model, idx, mfcc, g, with_nll: bool = True, warmup: int = 2, with_logits: bool = True):
        if not (idx.is_cuda and mfcc.is_cuda and g.is_cuda):
            raise _lib.WaeError("GraphedForward: example inputs must be CUDA tensors (no CPU fallback)")
        if model.training:
            raise RuntimeError("GraphedForward captures the inference forward: call model.eval() first")
        self.model = model
        self._slots, self._next = None, 0
        self.idx, self.mfcc, self.g = idx.clone(), mfcc.clone(), g.clone()
        classes = self.idx if not torch.is_floating_point(self.idx) else None
        self.with_nll = bool(with_nll and classes is not None)
        # loss-only evaluation: the NLL comes out of the head kernel's accumulator and the (B,O,T) logits are never written
        self.loss_only = bool(self.with_nll and not with_logits and hasattr(model, "forward_nll"))

        def run():
            with torch.no_grad():
                if self.loss_only:
                    nll, vq_loss, perp = model.forward_nll(self.idx, self.mfcc, self.g, classes, 1)
                    return None, vq_loss, perp, nll
                logits, vq_loss, perp = model(self.idx, self.mfcc, self.g)
                nll = teacher_forced_nll(logits, classes) if self.with_nll else None
            return logits, vq_loss, perp, nll

        side = torch.cuda.Stream(device=idx.device)
        side.wait_stream(torch.cuda.current_stream(idx.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):              # weight packing, workspace and allocator warm-up outside the capture
                run()
        torch.cuda.current_stream(idx.device).wait_stream(side)
        torch.cuda.synchronize(idx.device)
        self.launches = None
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.logits, self.vq_loss, self.perp, self.nll = run()
        self.launches = _lib.launch_count() - n0          # this library's kernels inside one replay

    # ---- pipelined use: results read one (or more) steps behind -------------------------------------------------------------
    # A caller that reads each step's loss before it launches the next one leaves the GPU idle for the host's round trip
    # (read-back, Python, graph launch: ~0.18 ms of a 3.0 ms step at 16 x 16000, tools/e2e_overhead.py).  submit() enqueues the
    # H2D copies, the replay and the D2H copy of the step's scalars into a pinned slot and returns at once; result() waits for
    # that slot only.  An evaluation loop keeps `depth - 1` steps in flight:   t = gf.submit(...); use(gf.result(t_prev)); t_prev = t
    def submit(self, idx, mfcc, g, depth: int = 2) -> int:
        """Enqueue one step (inputs of the captured shapes, pinned host or device tensors); returns a ticket for ``result``.
        At most ``depth`` tickets may be outstanding: the oldest slot is reused."""
        if self._slots is None or len(self._slots) != depth:
            outs = [t for t in (self.nll, self.vq_loss, self.perp)]
            self._slots = [([None if t is None else torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs],
                            torch.cuda.Event()) for _ in range(depth)]
            self._next = 0
        k = self._next
        self._next = (k + 1) % depth
        host, ev = self._slots[k]
        self.idx.copy_(idx, non_blocking=True)
        self.mfcc.copy_(mfcc, non_blocking=True)
        self.g.copy_(g, non_blocking=True)
        self.graph.replay()
        for h, t in zip(host, (self.nll, self.vq_loss, self.perp)):
            if h is not None:
                h.copy_(t, non_blocking=True)
        ev.record()
        return k

    def result(self, ticket: int):
        """``(nll, vq_loss, perplexity)`` of the step ``submit`` returned ``ticket`` for, as Python floats (None where the
        model has no such output); blocks until that step's device-to-host copies have landed."""
        host, ev = self._slots[ticket]
        ev.synchronize()
        return tuple(None if h is None else float(h) for h in host)

    def __call__(self, idx, mfcc, g):
        self.idx.copy_(idx, non_blocking=True)
        self.mfcc.copy_(mfcc, non_blocking=True)
        self.g.copy_(g, non_blocking=True)
        self.graph.replay()
        return self.logits, self.vq_loss, self.perp, self.nll
