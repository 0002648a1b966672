"""TEST INFRASTRUCTURE -- CPU restatement of the waveform post-processing of synthesis.py:382-394.  Only tests/ may import it.

PARITY UNPINNED for the mu-law part: ``P.inv_mulaw_quantize`` / ``P.inv_mulaw`` live in nnmnkwii (``pip install
wavenet_vocoder`` dependency, version not pinned by the reference, README.md:41), which is neither under /root/reference nor
installed here, so its published formulas are restated:

    inv_mulaw(y, mu)          = sign(y) * (1 / mu) * ((1 + mu) ** |y| - 1)
    inv_mulaw_quantize(k, mu) = inv_mulaw(2 * k / mu - 1, mu)

``inv_preemphasis(x, coef)`` is ``scipy.signal.lfilter([1], [1, -coef], x)`` (nnmnkwii wraps exactly that call; audio.py:64-65
passes coef = 0.85 by default), scipy is installed, so that part is the real thing.  Everything in float64.
"""
from __future__ import annotations

import numpy as np
from scipy import signal


def inv_mulaw(y, mu):
    y = np.asarray(y, np.float64)
    return np.sign(y) * (1.0 / mu) * ((1.0 + mu) ** np.abs(y) - 1.0)


def inv_mulaw_quantize(k, mu):
    return inv_mulaw(2.0 * np.asarray(k, np.float64) / mu - 1.0, mu)


def inv_preemphasis(x, coef):
    return signal.lfilter([1.0], [1.0, -coef], np.asarray(x, np.float64), axis=-1)


def waveform(y, input_type="mulaw-quantize", quantize_channels=256, postprocess=None, coef=0.85, gain=0.0):
    """y (B,T) classes / floats -> (B,T) float64, the chain of synthesis.py:382-394 per utterance."""
    if input_type == "mulaw-quantize":
        w = inv_mulaw_quantize(y, quantize_channels)
    elif input_type == "mulaw":
        w = inv_mulaw(y, quantize_channels)
    else:
        w = np.asarray(y, np.float64)
    if postprocess == "inv_preemphasis":
        w = inv_preemphasis(w, coef)
    if gain > 0:
        w = w / gain
    return w
