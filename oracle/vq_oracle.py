"""ORACLE (test infrastructure): VQ nearest-codeword search + the reference's loss / perplexity bookkeeping.

search()   -> ctypes wrapper of vq_oracle.c (exact fp32 + fmaf restatement; built by `make -C oracle`)
vq_forward / sliced_vq_forward -> the full forward of VectorQuantize (vector_quantization.py:21-49) and
SlicedVectorQuantize (:75-128) in numpy on top of search(); ema_update restates :190-217 / :282-294.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_SO = HERE / "_build" / "libvq_oracle.so"
_lib = None


def build():
    subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = C.CDLL(str(_SO))
        _lib.vq_oracle_search.restype = C.c_int
        _lib.vq_oracle_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def search(x: np.ndarray, codebook: np.ndarray, d0: int = 0, sub_d: int | None = None, quant: np.ndarray | None = None):
    """x (B,D,T) fp32 -> (idx (B*T,) int64, best (B*T,), second (B*T,)); fills quant rows [d0,d0+sub_d) if given."""
    x = np.ascontiguousarray(x, np.float32)
    cb = np.ascontiguousarray(codebook, np.float32)
    B, D, T = x.shape
    K, sd = cb.shape
    sub_d = sd if sub_d is None else sub_d
    assert sub_d == sd
    idx = np.empty(B * T, np.int64)
    best = np.empty(B * T, np.float32)
    second = np.empty(B * T, np.float32)
    if quant is not None:
        assert quant.flags["C_CONTIGUOUS"] and quant.dtype == np.float32 and quant.shape == x.shape
    rc = _load().vq_oracle_search(x.ctypes.data, B, D, T, d0, sd, cb.ctypes.data, K, idx.ctypes.data,
                                  None if quant is None else quant.ctypes.data, best.ctypes.data, second.ctypes.data)
    assert rc == 0
    return idx, best, second


def perplexity(idx: np.ndarray, K: int) -> np.float32:
    p = (np.bincount(idx, minlength=K).astype(np.float32) / np.float32(idx.shape[0])).astype(np.float32)
    return np.exp(-np.sum(p * np.log(p + np.float32(1e-10)))).astype(np.float32)


def vq_forward(x: np.ndarray, codebook: np.ndarray, beta: float = 0.25):
    """VectorQuantize.forward -> (quant (B,D,T), vq_loss, perp, idx (B,T))."""
    x = np.ascontiguousarray(x, np.float32)
    quant = np.empty_like(x)
    idx, _, _ = search(x, codebook, 0, None, quant)
    q = codebook[idx].reshape(x.shape[0], x.shape[2], -1).transpose(0, 2, 1)
    mse = np.mean((q.astype(np.float64) - x) ** 2)
    loss = np.float32(beta * mse + mse)
    return quant, loss, perplexity(idx, codebook.shape[0]), idx.reshape(x.shape[0], x.shape[2])


def sliced_vq_forward(x: np.ndarray, cb1: np.ndarray, cb2: np.ndarray, beta: float = 0.25):
    """SlicedVectorQuantize.forward (n_d=2) -> (quant, vq_loss, perp1+perp2, idx (B,T,2))."""
    x = np.ascontiguousarray(x, np.float32)
    sd = cb1.shape[1]
    quant = np.empty_like(x)
    i1, _, _ = search(x, cb1, 0, sd, quant)
    i2, _, _ = search(x, cb2, sd, sd, quant)
    B, D, T = x.shape
    q = np.concatenate([cb1[i1], cb2[i2]], axis=1).reshape(B, T, D).transpose(0, 2, 1)
    mse = np.mean((q.astype(np.float64) - x) ** 2)
    loss = np.float32(mse + beta * mse)
    perp = perplexity(i1, cb1.shape[0]) + perplexity(i2, cb2.shape[0])
    return quant, loss, perp, np.stack([i1.reshape(B, T), i2.reshape(B, T)], -1)


def ema_update(x_slice: np.ndarray, idx: np.ndarray, K: int, size: np.ndarray, w: np.ndarray, decay: float):
    """One EMA step for one codebook: x_slice (N,sd).  Returns (new_size, new_w, new_codebook)."""
    counts = np.bincount(idx, minlength=K).astype(np.float32)
    size = (size * np.float32(decay) + np.float32(1.0 - decay) * counts).astype(np.float32)
    n = np.sum(size)
    size = ((size + np.float32(1e-5)) / (n + K * np.float32(1e-5)) * n).astype(np.float32)
    dw = np.zeros_like(w)
    np.add.at(dw, idx, x_slice)
    w = (w * np.float32(decay) + np.float32(1 - decay) * dw).astype(np.float32)
    return size, w, (w / size[:, None]).astype(np.float32)
