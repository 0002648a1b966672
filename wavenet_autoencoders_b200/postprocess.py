"""Waveform post-processing of synthesis.py:382-394 on the device (additive API).

The reference moves the (B,256,T) one-hot output of ``incremental_forward`` to the host and applies, in numpy / nnmnkwii:
argmax -> ``P.inv_mulaw_quantize(y, hparams.quantize_channels)`` (``P.inv_mulaw`` for input_type "mulaw") ->
``getattr(audio, hparams.postprocess)`` (``inv_preemphasis`` = ``lfilter([1], [1, -coef])``, audio.py:64-65) -> division by
``hparams.global_gain_scale``.  ``waveform_from_synthesis`` does the same in one launch of ``wae_synth_postprocess`` on the
sampled class indices (``WaveNet.last_sampled_indices`` / ``return_indices=True``), so neither the one-hot tensor nor a
device->host copy of it is needed; only the final float waveform leaves the GPU.
"""
from __future__ import annotations

import torch

from . import _lib

_KINDS = {"mulaw-quantize": 0, "mulaw": 1, "raw": 2}


def waveform_from_synthesis(y: torch.Tensor, input_type: str = "mulaw-quantize", quantize_channels: int = 256,
                            postprocess: str | None = None, preemphasis_coef: float = 0.85,
                            global_gain_scale: float = 0.0) -> torch.Tensor:
    """y: (B,T) int64 classes, or the (B,O,T) one-hot / probability output (argmax over dim 1, as the reference does), or the
    (B,1,T) / (B,T) float output of the scalar-input models  ->  (B,T) fp32 waveform on the same device.

    ``postprocess``: None / "" / "none" or "inv_preemphasis" (the only post-processing audio.py offers for synthesis);
    ``preemphasis_coef``: audio.inv_preemphasis' default 0.85; ``global_gain_scale`` <= 0: no division."""
    if input_type not in _KINDS:
        raise ValueError(f"input_type must be one of {sorted(_KINDS)}, got {input_type!r}")
    if not y.is_cuda:
        raise _lib.WaeError("waveform_from_synthesis runs on CUDA sm_100 only (no CPU fallback)")
    kind = _KINDS[input_type]
    if kind == 0:
        if y.dim() == 3:
            y = y.max(1)[1]                                     # synthesis.py:383
        y = y.long().contiguous()
    else:
        if y.dim() == 3:
            if y.size(1) != 1:
                raise ValueError("scalar-output synthesis returns (B,1,T)")
            y = y[:, 0]
        y = y.float().contiguous()
    if y.dim() != 2:
        raise ValueError(f"expected (B,T) after squeezing, got {tuple(y.shape)}")
    if postprocess in (None, "", "none"):
        coef = 0.0
    elif postprocess == "inv_preemphasis":
        coef = float(preemphasis_coef)
    else:
        raise ValueError(f"unsupported postprocess {postprocess!r} (audio.py offers inv_preemphasis for synthesis)")
    B, T = y.shape
    out = torch.empty(B, T, dtype=torch.float32, device=y.device)
    _lib.check(_lib.lib().wae_synth_postprocess(_lib.ptr(y), kind, B, T, int(quantize_channels), coef, float(global_gain_scale),
                                                _lib.ptr(out), _lib.stream_ptr(y.device)), "wae_synth_postprocess")
    return out


def inv_mulaw_quantize_table(mu: int, device) -> torch.Tensor:
    """(mu + 1,) fp32: inv_mulaw_quantize(k, mu) = inv_mulaw(2 k / mu - 1, mu) for k = 0..mu, evaluated in float64 (the table
    wae_synth_postprocess builds per block)."""
    import numpy as np
    k = np.arange(mu + 1, dtype=np.float64)
    y = 2.0 * k / float(mu) - 1.0
    m = (np.power(1.0 + float(mu), np.abs(y)) - 1.0) / float(mu)
    return torch.tensor(np.sign(y) * m, dtype=torch.float32, device=device)


def ar_post_struct(device, B, T, categorical=True, input_type: str = "mulaw-quantize", quantize_channels: int = 256,
                   postprocess: str | None = None, preemphasis_coef: float = 0.85, global_gain_scale: float = 0.0):
    """The ``wae_ar_post`` descriptor for ``wae_ar_generate_wave`` (post-processing inside the synthesis kernel) with the
    conventions of ``waveform_from_synthesis``; returns (struct, output waveform tensor (B,T), tensors to keep alive)."""
    if input_type not in _KINDS:
        raise ValueError(f"input_type must be one of {sorted(_KINDS)}, got {input_type!r}")
    if categorical != (input_type == "mulaw-quantize"):
        raise ValueError(f"input_type {input_type!r} does not match the model's sampler")
    if postprocess in (None, "", "none"):
        coef = 0.0
    elif postprocess == "inv_preemphasis":
        coef = float(preemphasis_coef)
    else:
        raise ValueError(f"unsupported postprocess {postprocess!r} (audio.py offers inv_preemphasis for synthesis)")
    wave = torch.empty(B, T, dtype=torch.float32, device=device)
    table = inv_mulaw_quantize_table(int(quantize_channels), device) if categorical else None
    post = _lib.ArPost()
    post.table = None if table is None else table.data_ptr()
    post.mu = int(quantize_channels)
    post.scalar_is_mulaw = 1 if input_type == "mulaw" else 0
    post.preemphasis_coef, post.gain = coef, float(global_gain_scale)
    post.out_wave = wave.data_ptr()
    return post, wave, (table,)
