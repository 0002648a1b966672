"""Evaluation-time losses computed straight from the decoder's logits."""
from __future__ import annotations

import torch

from . import _lib


def teacher_forced_nll(logits: torch.Tensor, classes: torch.Tensor, shift: int = 1) -> torch.Tensor:
    """mean over b, t < T - shift of -log softmax(logits[b, :, t])[classes[b, t + shift]] -- the reference's training / eval
    criterion on full-length windows (vqwae_train.py:760-766: ``y_hat[:, :, :-1]`` against ``y[:, 1:]``) -- in ONE pass over
    the (B,O,T) fp32 logits (wae_nll_sum) instead of log_softmax + gather over a strided slice.  No autograd (use
    ``F.cross_entropy`` for training).  Returns a 0-dim fp32 tensor on the logits' device, no host sync."""
    if not logits.is_cuda:
        raise _lib.WaeError("teacher_forced_nll runs on CUDA sm_100 only (no CPU fallback)")
    B, O, T = logits.shape
    lg = logits.detach().float().contiguous()
    y = classes.detach().long().contiguous()
    out = torch.zeros(1, dtype=torch.float64, device=lg.device)
    _lib.check(_lib.lib().wae_nll_sum(_lib.ptr(lg), _lib.ptr(y), B, O, T, int(shift), _lib.ptr(out), _lib.stream_ptr(lg.device)),
               "wae_nll_sum")
    return (out[0] / float(B * (T - shift))).float()
