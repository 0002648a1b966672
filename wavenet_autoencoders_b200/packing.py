"""Weight packing for libwae_b200.so.

Everything here reads a WaveNet-shaped ``nn.Module`` purely through the attribute names the reference
defines (``first_conv``, ``conv_layers[i].{conv,conv1x1c,conv1x1g,conv1x1_out,conv1x1_skip}``,
``last_conv_layers[1|3]`` -- wavenet_vocoder/wavenet.py:119-141, modules.py:88-107), so it packs both
this package's modules and an imported reference model (the GPU parity tests use that).

Weight norm (modules.py:18, old-style ``weight_g``/``weight_v``) is folded here with the same ATen
primitive the reference's pre-forward hook uses (torch._weight_norm), so folded weights are bit-identical.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import torch

from . import _lib


def folded_weight(m: torch.nn.Module) -> torch.Tensor:
    """w = g * v / ||v||  (norm over all dims but 0), or the plain weight after remove_weight_norm."""
    if hasattr(m, "weight_g") and hasattr(m, "weight_v"):
        return torch._weight_norm(m.weight_v.detach(), m.weight_g.detach(), 0)
    return m.weight.detach()


def _bias(m, n, device):
    b = getattr(m, "bias", None)
    if b is None:
        return torch.zeros(n, dtype=torch.float32, device=device)
    return b.detach().float()


@dataclass
class StackShape:
    layers: int
    kernel_size: int
    R: int
    G: int
    S: int
    C: int
    Gi: int
    O: int
    Oin: int
    dilations: list

    @property
    def H(self):
        return self.G // 2

    def dims(self) -> _lib.StackDims:
        d = _lib.StackDims()
        d.layers, d.kernel_size = self.layers, self.kernel_size
        d.R, d.G, d.S, d.C, d.Gi, d.O, d.Oin = self.R, self.G, self.S, self.C, self.Gi, self.O, self.Oin
        for i, v in enumerate(self.dilations):
            d.dilation[i] = int(v)
        return d


def stack_shape(wn) -> StackShape:
    l0 = wn.conv_layers[0]
    conv = l0.conv
    G, R, kw = conv.out_channels, conv.in_channels, conv.kernel_size[0]
    S = l0.conv1x1_skip.out_channels
    C = l0.conv1x1c.in_channels if getattr(l0, "conv1x1c", None) is not None else 0
    Gi = l0.conv1x1g.in_channels if getattr(l0, "conv1x1g", None) is not None else 0
    O = wn.last_conv_layers[3].out_channels
    Oin = wn.first_conv.in_channels
    dil = [f.conv.dilation[0] for f in wn.conv_layers]
    if len(dil) > _lib.WAE_MAX_LAYERS:
        raise ValueError(f"at most {_lib.WAE_MAX_LAYERS} layers supported")
    return StackShape(len(dil), kw, R, G, S, C, Gi, O, Oin, dil)


_generation = 0


def bump_generation() -> None:
    """Parameters were rewritten behind autograd's back (through raw pointers, by a kernel or a CUDA-graph replay:
    train_step.FlatAdam.step, GraphedTrainStep.__call__) -- neither ``data_ptr`` nor ``_version`` moves then, so every cache of
    packed / transposed weights keys on this counter as well."""
    global _generation
    _generation += 1


def generation() -> int:
    return _generation


def params_fingerprint(wn) -> tuple:
    """Changes whenever any parameter of the module is modified in place, replaced, or rewritten through raw pointers by this
    package's optimiser (bump_generation)."""
    # called on every forward: walk a cached module list and the modules' own parameter dicts instead of Module.parameters()
    # (whose generator / de-duplication machinery costs ~0.3 ms for the 316 parameters of the vqwae decoder)
    mods = wn.__dict__.get("_fp_modules")
    if mods is None or sum(len(m._modules) for m in mods) != len(mods) - 1:      # a sub-module was added / removed: rebuild
        mods = list(wn.modules())
        wn.__dict__["_fp_modules"] = mods
    fp = [_generation]
    for m in mods:
        for p in m._parameters.values():
            if p is not None:
                fp.append(p.data_ptr())
                fp.append(p._version)
    return tuple(fp)


@dataclass
class FrontendPack:
    struct: "_lib.CondFrontend"
    total_scale: int
    keep: list


def pack_frontend(wn):
    """The conditioning front-end as plain arrays for wae_stack_forward_bf16_lat (upsample.py:29-85): conv_in weight transposed
    to [in][out], one folded (2s+1)-tap filter per stage.  None when the upsampler is not expressible that way (activation,
    frequency-axis kernel, cin_pad > 0, another interpolation mode): the caller then runs the stages one by one."""
    from .wavenet_vocoder import upsample as U
    net = wn.upsample_net
    conv_in = None
    if isinstance(net, U.ConvInUpsampleNetwork):
        conv_in, up = net.conv_in, net.upsample
    elif isinstance(net, U.UpsampleNetwork):
        up = net
    else:
        return None
    if (up.freq_axis_kernel_size != 1 or up.has_activation or up.mode != "nearest" or up.indent != 0
            or not 1 <= len(up.scales) <= 8):
        return None
    if conv_in is not None and (conv_in.kernel_size[0] != 1 or conv_in.bias is not None or conv_in.weight.shape[0] != conv_in.weight.shape[1]):
        return None
    lo, hi = 0, 127           # latent frames one 128-sample tile touches (the kernel's recursion): its frame buffer holds 24
    for sc in reversed(up.scales):
        lo, hi = (lo // sc if lo >= 0 else -1) - 1, hi // sc + 1
    if hi - lo + 1 + 2 > 24:
        return None
    fe = _lib.CondFrontend()
    keep = []
    if conv_in is not None:
        wt = folded_weight(conv_in).float()[:, :, 0].t().contiguous()      # (in, out)
        keep.append(wt)
        fe.conv_in_w_t = wt.data_ptr()
    else:
        fe.conv_in_w_t = None
    convs = [m for m in up.up_layers if isinstance(m, torch.nn.Conv2d)]
    fe.n_stages = len(up.scales)
    total = 1
    for i, (s, conv) in enumerate(zip(up.scales, convs)):
        w = folded_weight(conv).float().reshape(-1).contiguous()
        assert w.numel() == 2 * s + 1
        keep.append(w)
        fe.scale[i] = int(s)
        fe.filter[i] = w.data_ptr()
        total *= int(s)
    # the 3-coefficient form of every stage (what the kernels evaluate), summed once here in the kernels' order and precision
    import numpy as np
    coefs = []
    for s_, w in zip(up.scales, keep[-len(up.scales):]):
        wn_ = w.detach().cpu().numpy().astype(np.float32)
        a = np.zeros(s_, np.float32); b = np.zeros(s_, np.float32); c = np.zeros(s_, np.float32)
        for p in range(s_):
            x = np.float32(0.0)
            for j in range(0, s_ - p):
                x = np.float32(x + wn_[j])
            y = np.float32(0.0)
            for j in range(s_ - p, 2 * s_ - p):
                y = np.float32(y + wn_[j])
            z = np.float32(0.0)
            for j in range(2 * s_ - p, 2 * s_ + 1):
                z = np.float32(z + wn_[j])
            a[p], b[p], c[p] = x, y, z
        coefs += [a, b, c]
    coef = torch.tensor(np.concatenate(coefs), dtype=torch.float32, device=keep[-1].device)
    keep.append(coef)
    fe.coef = coef.data_ptr()
    return FrontendPack(fe, total, keep)


def _ru(x, m):
    return (x + m - 1) // m * m


_lane_streams = {}


class Lanes:
    """Fork / join of per-layer weight preparation over a few side streams.  A training step prepares ~100 small tensors per
    pass (weight-norm folds, packing permutations, gradient reshapes) -- chains of 2 us launches that a single stream (and the
    CUDA graph captured from it) executes one after the other; spread over lanes they become parallel branches of the graph.
    ``with lanes.lane(i): ...`` runs a block on side stream i (forked from the caller's stream on first use), ``join()`` makes the
    caller's stream wait for all of them.  Autograd replays each op's backward on the stream of its forward, so the weight-norm
    backward and the gradient accumulation of a layer land on its lane as well.  Same kernels, same results."""

    def __init__(self, device, n=8, also_wait=()):
        device = torch.device(device)
        self.also_wait = tuple(also_wait)        # streams every lane additionally waits for at its fork
        self.main = torch.cuda.current_stream(device)
        key = (device.index if device.index is not None else torch.cuda.current_device(), n)
        if key not in _lane_streams:
            _lane_streams[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
        self.side = _lane_streams[key]
        self.used = []

    def lane(self, i):
        st = self.side[i % len(self.side)]
        if st not in self.used:
            st.wait_stream(self.main)
            for other in self.also_wait:
                st.wait_stream(other)
            self.used.append(st)
        return torch.cuda.stream(st)

    def join(self):
        for st in self.used:
            self.main.wait_stream(st)
        self.used = []


_wgrad_streams = {}


def wgrad_stream(device, which="wgrad"):
    """The side streams of the two-stream decoder backward (one each per device): "wgrad" for its weight-gradient GEMMs,
    "bias" for the bias-gradient column sums."""
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), which)
    if key not in _wgrad_streams:
        _wgrad_streams[key] = torch.cuda.Stream(device=device)
    return _wgrad_streams[key]


def join_lane_streams_into_current(device):
    """Make the CURRENT stream wait for every lane stream of ``device`` that takes part in the ongoing work (all of them in
    eager mode, the ones inside the capture while a CUDA graph is being captured).  For code that runs inside autograd hooks --
    on whatever lane the last gradient of a bucket was accumulated -- and needs the gradients of ALL lanes."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    cur = torch.cuda.current_stream(device)
    capturing = torch.cuda.is_current_stream_capturing()
    for (dev_i, _), streams in _lane_streams.items():
        if dev_i != idx:
            continue
        for st in streams:
            if st == cur:
                continue
            if capturing:
                with torch.cuda.stream(st):
                    if not torch.cuda.is_current_stream_capturing():
                        continue
            cur.wait_stream(st)


class _NoLanes:
    def lane(self, i):
        import contextlib
        return contextlib.nullcontext()

    def join(self):
        pass


def _gather_layer_mats(wn, sh: StackShape, folded=None, lanes=None):
    """Per-layer folded fp32 matrices in natural layout.  ``folded``: the already folded weights of training.live_weights (8 per
    layer: conv, bias, conv1x1c, conv1x1g, conv1x1_out, bias, conv1x1_skip, bias) -- weight norm is then not evaluated again."""
    out = []
    dev = next(wn.parameters()).device
    lanes = lanes or _NoLanes()
    for l, f in enumerate(wn.conv_layers):
        with lanes.lane(l):
            if folded is not None:
                fw = folded[8 * l: 8 * l + 8]
                w, wc, wg, wo, ws = fw[0], fw[2], fw[3], fw[4], fw[6]
                w = w.detach().float()
                wc = wc.detach().float()[:, :, 0] if sh.C else None
                wg = wg.detach().float()[:, :, 0] if sh.Gi else None
                wo, ws = wo.detach().float()[:, :, 0], ws.detach().float()[:, :, 0]
            else:
                w = folded_weight(f.conv).float()                       # (G, R, kw)
                wc = folded_weight(f.conv1x1c).float()[:, :, 0] if sh.C else None   # (G, C)
                wg = folded_weight(f.conv1x1g).float()[:, :, 0] if sh.Gi else None  # (G, Gi)
                wo = folded_weight(f.conv1x1_out).float()[:, :, 0]      # (R, H)
                ws = folded_weight(f.conv1x1_skip).float()[:, :, 0]     # (S, H)
            out.append(dict(w=w, wc=wc, wg=wg, wo=wo, ws=ws,
                            b=_bias(f.conv, sh.G, dev), bo=_bias(f.conv1x1_out, sh.R, dev),
                            bs=_bias(f.conv1x1_skip, sh.S, dev)))
    return out


def _w1_kmajor(m, sh: StackShape, cpad: int) -> torch.Tensor:
    """(G, kw*R + cpad): taps oldest first (conv weight index j multiplies x[t-(kw-1-j)d], conv.py:56-61), then c."""
    parts = [m["w"][:, :, j] for j in range(sh.kernel_size)]
    if sh.C:
        wc = m["wc"]
        if cpad > sh.C:
            wc = torch.nn.functional.pad(wc, (0, cpad - sh.C))
        parts.append(wc)
    return torch.cat(parts, dim=1).contiguous()


def _gate_rows_bf16(w1: torch.Tensor, H: int) -> torch.Tensor:
    """(G, K) natural gate rows [tanh(0:H); sigmoid(0:H)] -> the row order of the tcgen05 layer kernels: each half zero
    padded to Hh = H rounded up to 16, and, when a half exceeds 128 channels (more than one UMMA N of 256 rows), split into
    two passes: [a(0:Ha); b(0:Ha); a(Ha:Hh); b(Ha:Hh)] with Ha = min(Hh, 128)  (include/wae_b200.h, wae_stack_bf16)."""
    Hh = _ru(H, 16)
    Ha = min(Hh, 128)
    a = torch.nn.functional.pad(w1[:H], (0, 0, 0, Hh - H))
    b = torch.nn.functional.pad(w1[H:2 * H], (0, 0, 0, Hh - H))
    blocks = [a[:Ha], b[:Ha]]
    if Hh > Ha:
        blocks += [a[Ha:], b[Ha:]]
    return torch.cat(blocks, dim=0)


class Packed:
    """Holds the packed tensors (keeps them alive) and the ctypes struct pointing at them."""

    def __init__(self):
        self.t = {}
        self.struct = None
        self.shape: StackShape | None = None


def pack_f32(wn) -> Packed:
    sh = stack_shape(wn)
    H = sh.H
    if sh.G % 8 or sh.R % 16 or sh.S % 8:
        raise _lib.WaeError(f"fp32 stack needs G%8==0, R%16==0, S%8==0 (got G={sh.G} R={sh.R} S={sh.S})")
    mats = _gather_layer_mats(wn, sh)
    dev = mats[0]["w"].device
    # pair-permuted gate columns: col 8q+i (i<4) = tanh channel 4q+i, col 8q+4+i = sigmoid channel 4q+i
    q = torch.arange(H // 4, device=dev).repeat_interleave(8)
    i = torch.arange(8, device=dev).repeat(H // 4)
    perm = torch.where(i < 4, 4 * q + i, H + 4 * q + (i - 4))
    p = Packed()
    p.shape = sh
    t = p.t
    t["w1"] = torch.stack([_w1_kmajor(m, sh, sh.C)[perm].t().contiguous() for m in mats])      # [L][K1][G]
    t["b1"] = torch.stack([m["b"][perm] for m in mats]).contiguous()
    t["wg"] = torch.stack([m["wg"][perm].t().contiguous() for m in mats]) if sh.Gi else None   # [L][Gi][G]
    t["w2"] = torch.stack([torch.cat([m["wo"], m["ws"]], 0).t().contiguous() for m in mats])   # [L][H][R+S]
    t["b2"] = torch.stack([torch.cat([m["bo"], m["bs"]]) for m in mats]).contiguous()
    t["wf"] = folded_weight(wn.first_conv).float()[:, :, 0].t().contiguous()                   # [Oin][R]
    t["bf"] = _bias(wn.first_conv, sh.R, dev).contiguous()
    l1, l3 = wn.last_conv_layers[1], wn.last_conv_layers[3]
    t["w3"] = folded_weight(l1).float()[:, :, 0].t().contiguous()                              # [S][S]
    t["b3"] = _bias(l1, sh.S, dev).contiguous()
    Opad = _ru(sh.O, 8)
    w4 = folded_weight(l3).float()[:, :, 0].t()                                                # [S][O]
    t["w4"] = torch.nn.functional.pad(w4, (0, Opad - sh.O)).contiguous()
    t["b4"] = torch.nn.functional.pad(_bias(l3, sh.O, dev), (0, Opad - sh.O)).contiguous()
    s = _lib.StackF32()
    s.d = sh.dims()
    for name in ("wf", "bf", "w1", "b1", "wg", "w2", "b2", "w3", "b3", "w4", "b4"):
        setattr(s, name, _lib.ptr(t[name]))
    p.struct = s
    return p


def pack_bf16(wn, folded=None, lanes=None) -> Packed:
    """``folded`` / ``lanes`` (training): reuse training.live_weights' folded weights and spread the per-layer permutations
    over side streams (Lanes); the caller joins the lanes before launching kernels that read the result."""
    sh = stack_shape(wn)
    H = sh.H
    mats = _gather_layer_mats(wn, sh, folded, lanes)
    dev = mats[0]["w"].device
    Hp, Cp, Op = _ru(H, 64), _ru(sh.C, 64) if sh.C else 0, _ru(sh.O, 16)
    bf = torch.bfloat16
    p = Packed()
    p.shape = sh
    t = p.t
    ln = lanes or _NoLanes()
    w1_rows, wo_rows, ws_rows = [], [], []
    for l, m in enumerate(mats):
        with ln.lane(l):
            w1_rows.append(_gate_rows_bf16(_w1_kmajor(m, sh, Cp), H))
            wo_rows.append(torch.nn.functional.pad(m["wo"], (0, Hp - H)))
            ws_rows.append(torch.nn.functional.pad(m["ws"], (0, Hp - H)))
    ln.join()
    t["w1"] = torch.stack(w1_rows).to(bf).contiguous()         # [L][2*Hh][K1p]
    t["wo"] = torch.stack(wo_rows).to(bf).contiguous()         # [L][R][Hp]
    t["ws"] = torch.stack(ws_rows).to(bf).contiguous()         # [L][S][Hp]
    l1, l3 = wn.last_conv_layers[1], wn.last_conv_layers[3]
    base = 8 * len(mats)
    fold = (lambda m, i: folded[base + i].detach()) if folded is not None else (lambda m, i: folded_weight(m))
    t["w3"] = fold(l1, 2).float()[:, :, 0].to(bf).contiguous()                                                    # [S][S]
    w4 = fold(l3, 4).float()[:, :, 0]                                                                             # [O][S]
    t["w4"] = torch.nn.functional.pad(w4, (0, 0, 0, Op - sh.O)).to(bf).contiguous()                               # [Op][S]
    t["b1"] = torch.stack([m["b"] for m in mats]).contiguous()
    t["wg"] = torch.stack([m["wg"].t().contiguous() for m in mats]) if sh.Gi else None                            # [L][Gi][G]
    t["bo"] = torch.stack([m["bo"] for m in mats]).contiguous()
    t["bs_sum"] = torch.stack([m["bs"] for m in mats]).sum(0).contiguous()
    t["b3"] = _bias(l1, sh.S, dev).contiguous()
    t["b4"] = torch.nn.functional.pad(_bias(l3, sh.O, dev), (0, Op - sh.O)).contiguous()
    t["wf"] = fold(wn.first_conv, 0).float()[:, :, 0].t().contiguous()
    t["bf"] = _bias(wn.first_conv, sh.R, dev).contiguous()
    # class-index input: first conv = one row of bf16(wf + bf) per sample (the same fp32 add + rounding the kernels did per sample)
    t["wfb"] = torch.cat([t["wf"] + t["bf"][None, :], t["bf"][None, :]], dim=0).to(bf).contiguous() if sh.Oin > 1 else None
    s = _lib.StackBF16()
    s.d = sh.dims()
    for name in ("w1", "wo", "ws", "w3", "w4", "b1", "wg", "bo", "bs_sum", "b3", "b4", "wf", "bf", "wfb"):
        setattr(s, name, _lib.ptr(t[name]))
    p.struct = s
    return p


def part(n: int, r: int, cs: int) -> int:
    return (n * r) // cs


def pack_ar(wn, cluster: int = 8, wtype: str = "bf16", utts_per_cluster: int = 2) -> Packed:
    """Per-(stage, rank) row-sliced blobs for the cluster AR kernel (see include/wae_b200.h)."""
    sh = stack_shape(wn)
    H, L = sh.H, sh.layers
    mats = _gather_layer_mats(wn, sh)
    dev = mats[0]["w"].device
    Hp, Cp = _ru(H, 64), _ru(sh.C, 64) if sh.C else 0
    mma = (wtype == "bf16mma")      # tensor-core AR kernel: plain row-major [rows padded to 16][K + 8] bf16 blobs
    dt = torch.float32 if wtype == "fp32" else torch.bfloat16
    esz = 4 if wtype == "fp32" else 2
    l1, l3 = wn.last_conv_layers[1], wn.last_conv_layers[3]
    w3 = folded_weight(l1).float()[:, :, 0]
    w4 = folded_weight(l3).float()[:, :, 0]
    chunks, offs, off = [], [], 0

    def add(mat: torch.Tensor, lanes: int):
        """mat [rows][K] -> [K/(4*lanes)][rows][4*lanes]: `lanes` lanes share one row in the AR kernel's mat-vecs (4
        consecutive elements per lane), so the 32 lanes of a warp read one contiguous run of shared memory."""
        nonlocal off
        rows, K = mat.shape
        if mma:
            mat = torch.nn.functional.pad(mat, (0, 8, 0, (-rows) % 16))   # 16-byte row padding -> conflict-free ldmatrix
        else:
            ch = 4 * lanes
            assert K % ch == 0
            mat = mat.reshape(rows, K // ch, ch).permute(1, 0, 2)
        raw = mat.to(dt).contiguous().view(torch.uint8).flatten()
        n = raw.numel()
        pad = (-n) % 16
        if pad:
            raw = torch.cat([raw, torch.zeros(pad, dtype=torch.uint8, device=dev)])
        offs.append(off)
        chunks.append(raw)
        off += n + pad

    for m in mats:
        w1 = _w1_kmajor(m, sh, Cp)                                           # (G, K1p)
        for r in range(cluster):                                             # stage 2l
            p0, p1 = part(H, r, cluster), part(H, r + 1, cluster)
            rows = torch.stack([w1[p0:p1], w1[H + p0:H + p1]], dim=1).reshape(-1, w1.shape[1])  # a_p, b_p interleaved
            add(rows, 8)
        wo = torch.nn.functional.pad(m["wo"], (0, Hp - H))
        ws = torch.nn.functional.pad(m["ws"], (0, Hp - H))
        for r in range(cluster):                                             # stage 2l+1
            add(torch.cat([wo[part(sh.R, r, cluster):part(sh.R, r + 1, cluster)],
                           ws[part(sh.S, r, cluster):part(sh.S, r + 1, cluster)]], 0), 4)
    for r in range(cluster):
        add(w3[part(sh.S, r, cluster):part(sh.S, r + 1, cluster)], 8)
    for r in range(cluster):
        add(w4[part(sh.O, r, cluster):part(sh.O, r + 1, cluster)], 8)
    p = Packed()
    p.shape = sh
    t = p.t
    t["blob"] = torch.cat(chunks) if chunks else torch.zeros(16, dtype=torch.uint8, device=dev)
    t["layer_off"] = torch.tensor(offs, dtype=torch.int64, device=dev)
    t["b1"] = torch.stack([m["b"] for m in mats]).contiguous()
    t["wg"] = torch.stack([m["wg"].t().contiguous() for m in mats]) if sh.Gi else None
    t["bo"] = torch.stack([m["bo"] for m in mats]).contiguous()
    t["bs"] = torch.stack([m["bs"] for m in mats]).contiguous()
    t["b3"] = _bias(l1, sh.S, dev).contiguous()
    t["b4"] = _bias(l3, sh.O, dev).contiguous()
    t["wf"] = folded_weight(wn.first_conv).float()[:, :, 0].t().contiguous()
    t["bf"] = _bias(wn.first_conv, sh.R, dev).contiguous()
    s = _lib.ArWeights()
    s.d = sh.dims()
    s.wtype = {"fp32": 0, "bf16": 1, "bf16mma": 2}[wtype]
    s.cluster = cluster
    s.utts_per_cluster = utts_per_cluster
    s.blob = _lib.ptr(t["blob"])
    s.layer_off = _lib.ptr(t["layer_off"])
    for name in ("b1", "wg", "bo", "bs", "b3", "b4", "wf", "bf"):
        setattr(s, name, _lib.ptr(t[name]))
    p.struct = s
    assert esz * 0 == 0
    return p


class WorkspaceCache:
    """Grow-only scratch buffer per device (the C ABI never allocates)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != torch.device(device):
            self.buf = None
            self.buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        return self.buf
