"""ORACLE (test infrastructure, never shipped, never imported by the product package).

A CPU/numpy restatement of the reference's WaveNet-decoder hot path, written from the algorithm, working
from a plain ``state_dict`` of numpy arrays (no torch modules):

    fold_weight_norm        wavenet_vocoder/modules.py:18  (torch._weight_norm, dim=0)
    upsample_conditioning   wavenet_vocoder/upsample.py:18-20, 42, 53-64, 78-85
    stack_forward           wavenet_vocoder/wavenet.py:203-212 + modules.py:115-163
    incremental_forward     wavenet_vocoder/wavenet.py:299-339 + conv.py:17-46 (per-step, ring history)

Pinned (tests/test_oracle_cpu.py) against golden vectors produced by the REAL reference modules
(tools/make_golden.py, run where /root/reference exists; fixtures in tests/golden/).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def fold_weight_norm(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """w = g * v / ||v||, norm over every dim but 0."""
    v = v.astype(F32)
    norm = np.sqrt(np.sum(v.astype(np.float64) ** 2, axis=tuple(range(1, v.ndim)), keepdims=True)).astype(F32)
    return (v * (g.astype(F32) / norm)).astype(F32)


def _weight(sd, name):
    if name + ".weight_v" in sd:
        return fold_weight_norm(np.asarray(sd[name + ".weight_v"]), np.asarray(sd[name + ".weight_g"]))
    return np.asarray(sd[name + ".weight"], dtype=F32)


def _bias(sd, name, n):
    return np.asarray(sd[name + ".bias"], dtype=F32) if name + ".bias" in sd else np.zeros(n, F32)


def extract_params(sd: dict, layers: int, stacks: int, prefix: str = "") -> dict:
    """Fold weight norm and collect the decoder's matrices from a state_dict (keys as in SURVEY.md 3.4)."""
    p = {"layers": []}
    per = layers // stacks
    wf = _weight(sd, prefix + "first_conv")
    p["wf"], p["bf"] = wf[:, :, 0], _bias(sd, prefix + "first_conv", wf.shape[0])
    for l in range(layers):
        n = f"{prefix}conv_layers.{l}."
        w = _weight(sd, n + "conv")
        lay = {"w": w, "b": _bias(sd, n + "conv", w.shape[0]), "dilation": 2 ** (l % per)}
        lay["wc"] = _weight(sd, n + "conv1x1c")[:, :, 0] if (n + "conv1x1c.weight_v" in sd or n + "conv1x1c.weight" in sd) else None
        lay["wg"] = _weight(sd, n + "conv1x1g")[:, :, 0] if (n + "conv1x1g.weight_v" in sd or n + "conv1x1g.weight" in sd) else None
        lay["wo"] = _weight(sd, n + "conv1x1_out")[:, :, 0]
        lay["bo"] = _bias(sd, n + "conv1x1_out", lay["wo"].shape[0])
        lay["ws"] = _weight(sd, n + "conv1x1_skip")[:, :, 0]
        lay["bs"] = _bias(sd, n + "conv1x1_skip", lay["ws"].shape[0])
        p["layers"].append(lay)
    w3 = _weight(sd, prefix + "last_conv_layers.1")
    w4 = _weight(sd, prefix + "last_conv_layers.3")
    p["w3"], p["b3"] = w3[:, :, 0], _bias(sd, prefix + "last_conv_layers.1", w3.shape[0])
    p["w4"], p["b4"] = w4[:, :, 0], _bias(sd, prefix + "last_conv_layers.3", w4.shape[0])
    if prefix + "embed_speakers.weight" in sd:
        p["embed"] = np.asarray(sd[prefix + "embed_speakers.weight"], dtype=F32)
    if prefix + "upsample_net.conv_in.weight" in sd:
        p["conv_in"] = np.asarray(sd[prefix + "upsample_net.conv_in.weight"], dtype=F32)
        ups, i = [], 1
        while f"{prefix}upsample_net.upsample.up_layers.{i}.weight_v" in sd or f"{prefix}upsample_net.upsample.up_layers.{i}.weight" in sd:
            ups.append(_weight(sd, f"{prefix}upsample_net.upsample.up_layers.{i}").reshape(-1))
            i += 2
        p["up"] = ups
    return p


def upsample_conditioning(p: dict, c: np.ndarray) -> np.ndarray:
    """ConvInUpsampleNetwork: 1x1 conv_in (cin_pad=0), then per stage nearest stretch by s and a zero-padded
    (2s+1)-tap smoothing filter along time, one channel at a time."""
    w = p["conv_in"]
    assert w.shape[2] == 1, "oracle restates cin_pad=0 (every preset)"
    c = np.einsum("oi,bit->bot", w[:, :, 0], c.astype(F32)).astype(F32)
    for k in p["up"]:
        s = (len(k) - 1) // 2
        c = np.repeat(c, s, axis=-1)
        padded = np.pad(c, ((0, 0), (0, 0), (s, s)))
        out = np.zeros_like(c)
        for j in range(2 * s + 1):
            out += k[j] * padded[:, :, j:j + c.shape[-1]]
        c = out.astype(F32)
    return c


def speaker_vectors(p: dict, g) -> np.ndarray | None:
    if g is None:
        return None
    g = np.asarray(g)
    if "embed" in p and np.issubdtype(g.dtype, np.integer):
        return p["embed"][g.reshape(-1)]
    return g.reshape(g.shape[0], -1).astype(F32)


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(F32)))).astype(F32)


def layer_forward(lay: dict, x: np.ndarray, c: np.ndarray | None, gvec: np.ndarray | None):
    """One ResidualConv1dGLU on (B,R,T): returns (x_out, skip)."""
    B, R, T = x.shape
    w, d = lay["w"], lay["dilation"]
    kw = w.shape[2]
    pad = (kw - 1) * d
    xp = np.pad(x, ((0, 0), (0, 0), (pad, 0)))
    z = np.broadcast_to(lay["b"][None, :, None], (B, w.shape[0], T)).astype(F32).copy()
    for j in range(kw):                                  # tap j multiplies x[t - (kw-1-j) d]
        z += np.einsum("gr,brt->bgt", w[:, :, j], xp[:, :, j * d:j * d + T]).astype(F32)
    if c is not None:
        z += np.einsum("gc,bct->bgt", lay["wc"], c).astype(F32)
    if gvec is not None:
        z += (gvec @ lay["wg"].T)[:, :, None].astype(F32)
    H = z.shape[1] // 2
    h = (np.tanh(z[:, :H]) * _sigmoid(z[:, H:])).astype(F32)
    s = (np.einsum("sh,bht->bst", lay["ws"], h) + lay["bs"][None, :, None]).astype(F32)
    o = (np.einsum("rh,bht->brt", lay["wo"], h) + lay["bo"][None, :, None]).astype(F32)
    return ((o + x) * F32(math.sqrt(0.5))).astype(F32), s


def stack_forward(p: dict, x: np.ndarray, c_up: np.ndarray | None, gvec: np.ndarray | None) -> np.ndarray:
    """first_conv -> L gated residual layers with skip sum * sqrt(1/L) -> ReLU,1x1,ReLU,1x1.  (B,Oin,T)->(B,O,T)."""
    x = (np.einsum("ro,bot->brt", p["wf"], x.astype(F32)) + p["bf"][None, :, None]).astype(F32)
    skips = None
    for lay in p["layers"]:
        x, s = layer_forward(lay, x, c_up, gvec)
        skips = s if skips is None else (skips + s).astype(F32)
    a = np.maximum(skips * F32(math.sqrt(1.0 / len(p["layers"]))), 0).astype(F32)
    a = np.maximum(np.einsum("os,bst->bot", p["w3"], a) + p["b3"][None, :, None], 0).astype(F32)
    return (np.einsum("os,bst->bot", p["w4"], a) + p["b4"][None, :, None]).astype(F32)


def forward(p: dict, x, c=None, g=None) -> np.ndarray:
    """WaveNet.forward(x, c, g) with un-upsampled c and speaker ids."""
    c_up = None if c is None else (upsample_conditioning(p, c) if "conv_in" in p else np.asarray(c, F32))
    if c_up is not None and c_up.shape[-1] != x.shape[-1]:
        raise Exception(f"c {c_up.shape} x {x.shape}")
    return stack_forward(p, np.asarray(x, F32), c_up, speaker_vectors(p, g))


class IncrementalState:
    """Per-layer history of the last (kw-1)*d inputs (what conv.py:34-41 keeps in its shift buffer)."""

    def __init__(self, p: dict, B: int):
        self.hist = []
        for lay in p["layers"]:
            kw, R = lay["w"].shape[2], lay["w"].shape[1]
            self.hist.append(np.zeros((B, (kw - 1) * lay["dilation"] + 1, R), F32))
        self.t = 0


def step(p: dict, st: IncrementalState, x_in: np.ndarray, c_t: np.ndarray | None, gvec: np.ndarray | None):
    """One autoregressive step: x_in (B,Oin) -> logits (B,O)."""
    x = (x_in.astype(F32) @ p["wf"].T + p["bf"]).astype(F32)
    skips = None
    for lay, hist in zip(p["layers"], st.hist):
        w, d = lay["w"], lay["dilation"]
        kw = w.shape[2]
        ns = hist.shape[1]
        hist[:, st.t % ns] = x
        z = np.broadcast_to(lay["b"], (x.shape[0], w.shape[0])).astype(F32).copy()
        for j in range(kw):
            ts = st.t - (kw - 1 - j) * d
            if ts >= 0:
                z += (hist[:, ts % ns] @ w[:, :, j].T).astype(F32)
        if c_t is not None:
            z += (c_t @ lay["wc"].T).astype(F32)
        if gvec is not None:
            z += (gvec @ lay["wg"].T).astype(F32)
        H = z.shape[1] // 2
        h = (np.tanh(z[:, :H]) * _sigmoid(z[:, H:])).astype(F32)
        s = (h @ lay["ws"].T + lay["bs"]).astype(F32)
        x = ((h @ lay["wo"].T + lay["bo"] + x) * F32(math.sqrt(0.5))).astype(F32)
        skips = s if skips is None else (skips + s).astype(F32)
    st.t += 1
    a = np.maximum(skips * F32(math.sqrt(1.0 / len(p["layers"]))), 0).astype(F32)
    a = np.maximum(a @ p["w3"].T + p["b3"], 0).astype(F32)
    return (a @ p["w4"].T + p["b4"]).astype(F32)


def incremental_forward(p: dict, T: int, c=None, g=None, initial_input=None, test_inputs=None,
                        sampler=None) -> np.ndarray:
    """The loop of WaveNet.incremental_forward.  ``test_inputs`` (B,Tf,Oin) teacher-forces steps t<Tf;
    ``sampler(t, logits) -> next input (B,Oin)`` produces the fed-back value (default: feed logits back,
    i.e. softmax=False, quantize=False).  Returns the per-step sampler outputs / logits as (B,T,*)."""
    c_up = None if c is None else (upsample_conditioning(p, c) if "conv_in" in p else np.asarray(c, F32))
    B = (test_inputs.shape[0] if test_inputs is not None else (c_up.shape[0] if c_up is not None else initial_input.shape[0]))
    gvec = speaker_vectors(p, g)
    st = IncrementalState(p, B)
    cur = initial_input
    outs = []
    for t in range(T):
        if test_inputs is not None and t < test_inputs.shape[1]:
            cur = test_inputs[:, t]
        c_t = None if c_up is None else c_up[:, :, t]
        logits = step(p, st, np.asarray(cur, F32), c_t, gvec)
        out = logits if sampler is None else sampler(t, logits)
        outs.append(out)
        cur = out
    return np.stack(outs, axis=1)
