"""ORACLE (test infrastructure, never shipped, never imported by the product package).

A numpy restatement of the reference's frame-rate encoder, written from the algorithm, working from a plain ``state_dict``
of numpy arrays (no torch modules):

    conv_relu_res     vqvae_model.py:9-23   Conv1d(k, stride, padding = k // 2, bias) -> ReLU -> (+ x when stride == 1 and
                                             dim_in == dim_out)
    encoder_forward   vqvae_model.py:25-51  ten ConvReLURes blocks (k = 3,3,5,5,3,3,1,1,1,1; strides 1,1,2,2,1,...) then Linear
                                             over the channel axis

Pinned (tests/test_oracle_cpu.py) against the latents the REAL reference produced (goldens ``vqvae_tiny`` and
``vqvae_vqwae``, tools/make_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this package.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def conv1d(x: np.ndarray, w: np.ndarray, b: np.ndarray | None, stride: int) -> np.ndarray:
    """x (B, Cin, T), w (Cout, Cin, k), zero padding k // 2 on both sides -> (B, Cout, (T - 1) // stride + 1), fp32 products
    accumulated in float64 per output (the reference's library sums in an order of its own; this bounds it from the exact side)."""
    B, cin, T = x.shape
    cout, _, k = w.shape
    pad = k // 2
    xp = np.zeros((B, cin, T + 2 * pad), np.float64)
    xp[:, :, pad:pad + T] = x
    Tout = (T - 1) // stride + 1
    out = np.zeros((B, cout, Tout), np.float64)
    w64 = w.astype(np.float64)
    for j in range(k):
        seg = xp[:, :, j:j + (Tout - 1) * stride + 1:stride]            # (B, Cin, Tout): input at position o * stride + j - pad
        out += np.einsum("oc,bct->bot", w64[:, :, j], seg)
    if b is not None:
        out += b.astype(np.float64)[None, :, None]
    return out.astype(F32)


def conv_relu_res(x: np.ndarray, w: np.ndarray, b: np.ndarray | None, stride: int) -> np.ndarray:
    y = np.maximum(conv1d(x, w, b, stride), F32(0))
    if stride == 1 and w.shape[0] == w.shape[1]:
        y = (y + x).astype(F32)
    return y


def encoder_forward(sd: dict, x: np.ndarray, prefix: str = "encoder.") -> np.ndarray:
    """sd: state_dict entries ``{prefix}net.{i}.conv.weight / .bias`` and ``{prefix}lin.weight / .bias`` as numpy arrays;
    x (B, c_in, frames) -> latents (B, c_out, frames')."""
    h = np.ascontiguousarray(x, F32)
    i = 0
    while f"{prefix}net.{i}.conv.weight" in sd:
        w = np.asarray(sd[f"{prefix}net.{i}.conv.weight"], F32)
        b = sd.get(f"{prefix}net.{i}.conv.bias")
        k = w.shape[2]
        stride = 2 if k == 5 else 1                                      # vqvae_model.py:33-40: the two k = 5 blocks have stride 2
        h = conv_relu_res(h, w, None if b is None else np.asarray(b, F32), stride)
        i += 1
    wl = np.asarray(sd[f"{prefix}lin.weight"], np.float64)               # (c_out, hid)
    bl = sd.get(f"{prefix}lin.bias")
    out = np.einsum("dc,bct->bdt", wl, h.astype(np.float64))
    if bl is not None:
        out += np.asarray(bl, np.float64)[None, :, None]
    return out.astype(F32)
