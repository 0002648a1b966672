"""ORACLE (test infrastructure): samplers used inside WaveNet.incremental_forward.

  categorical_from_uniform   the RNG contract of the fused AR kernel (SURVEY.md 8c "L2"): softmax -> inverse
                             CDF with a caller-supplied uniform.  The reference draws with torch.multinomial
                             (wavenet.py:335-338), whose stream a fused kernel cannot replay, so parity is defined
                             on (probabilities, uniform) -> class.  The summation ORDER below mirrors
                             csrc/wn_ar.cu exactly (lane-local sequential sums, Kogge-Stone scan over 32 lanes),
                             which makes the class bit-reproducible.
  mol_from_uniform           mixture.py:118-156 with the uniforms supplied (u' = 1e-5 + u*(1-2e-5), as Tensor.uniform_(a,b))
  gauss_from_draws           mixture.py:221-270 with the uniforms / normal draw supplied
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def softmax_unnormalised(logits: np.ndarray):
    """(ex, total) in the kernel's summation order for one row of O <= 256 logits."""
    O = logits.shape[0]
    per = (O + 31) // 32
    lg = np.full(32 * per, -np.inf, F32)
    lg[:O] = logits.astype(F32)
    mx = F32(lg.max())
    ex = np.where(np.arange(32 * per) < O, np.exp((lg - mx).astype(F32)).astype(F32), F32(0)).astype(F32)
    ex = ex.reshape(32, per)
    loc = np.zeros(32, F32)
    for i in range(per):
        loc = (loc + ex[:, i]).astype(F32)
    inc = loc.copy()
    off = 1
    while off < 32:
        shifted = np.concatenate([np.zeros(off, F32), inc[:-off]])
        inc = np.where(np.arange(32) >= off, (inc + shifted).astype(F32), inc).astype(F32)
        off <<= 1
    return ex, loc, inc, inc[31]


def categorical_from_uniform(logits: np.ndarray, u: float) -> int:
    O = logits.shape[0]
    ex, loc, inc, total = softmax_unnormalised(logits)
    per = ex.shape[1]
    thr = F32(F32(u) * total)
    for lane in range(32):
        run = F32(inc[lane] - loc[lane])
        for i in range(per):
            run = F32(run + ex[lane, i])
            o = lane * per + i
            if o < O and run > thr:
                return o
    return O - 1


def softmax_probs(logits: np.ndarray) -> np.ndarray:
    O = logits.shape[0]
    ex, _, _, total = softmax_unnormalised(logits)
    inv = F32(1.0) / total
    return (ex.reshape(-1)[:O] * inv).astype(F32)


def mol_from_uniform(y: np.ndarray, u: np.ndarray) -> float:
    """y (3*nmix,) = [logits | means | log_scales]; u (nmix+1,) uniforms in [0,1)."""
    nmix = y.shape[0] // 3
    uq = (F32(1e-5) + u[:nmix].astype(F32) * F32(1.0 - 2e-5)).astype(F32)
    g = (y[:nmix].astype(F32) - np.log(-np.log(uq))).astype(F32)
    k = int(np.argmax(g))
    ul = F32(1e-5) + F32(u[nmix]) * F32(1.0 - 2e-5)
    x = F32(y[nmix + k]) + np.exp(F32(y[2 * nmix + k])) * (np.log(ul) - np.log(F32(1.0) - ul))
    return float(np.clip(F32(x), -1.0, 1.0))


def gauss_from_draws(y: np.ndarray, u: np.ndarray) -> float:
    """y (2,), (3,) or (3*nmix,); u = nmix uniforms followed by one N(0,1) draw."""
    C = y.shape[0]
    nmix = 1 if C == 2 else C // 3
    if C == 2:
        mean, ls = y[0], y[1]
    elif nmix == 1:
        mean, ls = y[1], y[2]
    else:
        uq = (F32(1e-5) + u[:nmix].astype(F32) * F32(1.0 - 2e-5)).astype(F32)
        k = int(np.argmax((y[:nmix].astype(F32) - np.log(-np.log(uq))).astype(F32)))
        mean, ls = y[nmix + k], y[2 * nmix + k]
    x = F32(mean) + np.exp(F32(ls)) * F32(u[nmix])
    return float(np.clip(F32(x), -1.0, 1.0))
