"""Training path of the decoder stack (BASELINE configs[2], SURVEY 8(b) "autograd-differentiable forward").

Forward: the same tcgen05 kernels as inference (``wae_stack_forward_bf16_save``), which additionally keep every layer
input ``x_all``, the gated activations ``h_all`` and the channels-last conditioning for the backward pass.

Backward: the layer equations of ``ResidualConv1dGLU._forward`` (modules.py:115-163) differentiated by hand on those saved
bf16 channels-last tensors.  On the GPU every contraction runs on this library's tcgen05 kernels (``wae_stack_backward_bf16``,
csrc/wn_bwd.cu: dgrad-shaped GEMMs with the dilated taps as TMA boxes and the gate derivative / residual / ReLU masks fused in
the epilogues, MN-major split-K wgrads) -- ``_stack_backward_tc`` below only packs the transposed weights and scatters the
fp32 gradients back to the parameters.  ``stack_backward`` keeps the same derivation as torch expressions: it is what the
float64 CPU test differentiates against autograd, and the fp32 twin the GPU test compares the kernels with.  Per layer:

    Xcat = [x(t-2d) | x(t-d) | x(t) | c(t)]                 (B,T,kw*R+Cp)   gathered once per layer
    z    = Xcat W1cat^T + gb            (recomputed: only h = tanh*sigmoid was stored, 256 B/sample/layer instead of 768)
    dh   = dS Ws + (dx' sqrt(.5)) Wo                        skip term for all layers comes from ONE GEMM over K = L*H
    dz   = [dh sig (1-tanh^2) | dh tanh sig (1-sig)]
    dW1cat = dz^T Xcat ,  dXcat = dz W1cat  ->  dx (taps shifted back) , dc
    dWo  = (dx' sqrt(.5))^T h

The result is a custom ``torch.autograd.Function`` whose inputs are the weight-norm-folded weights, so the gradients flow on
to ``weight_g`` / ``weight_v``, the speaker embedding and the upsampling network through ordinary autograd.
"""
from __future__ import annotations

import math

import os

import torch
from torch.nn import functional as F

from . import _lib, packing

BF = torch.bfloat16


def _fw(m):
    """Weight-norm folded weight WITH autograd history (packing.folded_weight is the detached twin)."""
    if hasattr(m, "weight_g") and hasattr(m, "weight_v"):
        return torch._weight_norm(m.weight_v, m.weight_g, 0)
    return m.weight


def live_weights(wn, lanes=None):
    """Flat list of the stack's folded weights and biases, fixed order (None where a conv has no bias / is absent).
    ``lanes`` (packing.Lanes): the folds of layer l run on side stream l -- and so will their backward."""
    out = []
    ln = lanes or packing._NoLanes()
    for l, f in enumerate(wn.conv_layers):
        with ln.lane(l):
            out += [_fw(f.conv), f.conv.bias,
                    _fw(f.conv1x1c) if f.conv1x1c is not None else None,
                    _fw(f.conv1x1g) if f.conv1x1g is not None else None,
                    _fw(f.conv1x1_out), f.conv1x1_out.bias, _fw(f.conv1x1_skip), f.conv1x1_skip.bias]
    l1, l3 = wn.last_conv_layers[1], wn.last_conv_layers[3]
    with ln.lane(len(wn.conv_layers)):          # on a lane as well: their gradients come from the wgrad stream (two-stream backward)
        out += [_fw(wn.first_conv), wn.first_conv.bias, _fw(l1), l1.bias, _fw(l3), l3.bias]
    return out


def _lanes_for(wn, x):
    """Side-stream lanes for the weight preparation of a CUDA training step (``wn.prep_lanes``, default 8; 0 disables)."""
    n = int(getattr(wn, "prep_lanes", os.environ.get("WAE_PREP_LANES", "8")))
    if not (x.is_cuda and n > 0):
        return None
    # gradient accumulators of parameters first used on the main stream (an earlier eager step) now receive gradients produced
    # on a lane: intended, autograd synchronises the two streams -- silence its per-step warning about it
    quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
    if quiet is not None:
        quiet(False)
    return packing.Lanes(x.device, n)


PER_LAYER = 8


def _shift(x, s):
    """x[b, t - s] with zeros for t < s (the causal left padding, modules.py:80-85)."""
    return x if s == 0 else F.pad(x, (0, 0, s, 0))[:, : x.shape[1]]


def _unshift(y, s):
    """adjoint of _shift: y[b, t + s], zeros past the end."""
    return y if s == 0 else F.pad(y, (0, 0, 0, s))[:, s:]


def _mm_acc(a, b, adt):
    """a (..., K) @ b (K, N) with the result kept in the accumulation dtype: the ReLU masks of the recomputed head must be
    decided on the fp32 sums the forward kernel saw in TMEM, not on sums rounded to bf16 first."""
    if a.dtype == adt:
        return a @ b
    a2 = a.reshape(-1, a.shape[-1])
    try:
        out = torch.mm(a2, b, out_dtype=adt)
    except (TypeError, RuntimeError, NotImplementedError):
        out = (a2 @ b).to(adt)
    return out.reshape(*a.shape[:-1], b.shape[-1])


_ONES = {}
_COLSUM_MM = [True]        # cleared on the first failure so that a CUDA-graph capture never retries a failing call


def _colsum(t, adt=torch.float32):
    """Sum over all but the last dimension, accumulated in adt.  For the tall bf16 (B*T, N) activations-gradients on the GPU the
    sum is a 1 x (B*T) ones-row GEMM with an fp32 result: torch's column reduction of a row-major tall matrix takes 77 us per
    call at 61440 x 256 (23 calls = 1.8 ms of the 15 ms training step, tools/train_profile.py); the GEMM reads the 31 MB once."""
    if _COLSUM_MM[0] and t.is_cuda and t.dtype == BF and adt == torch.float32 and t.dim() >= 2 and t.numel() >= (1 << 18):
        t2 = t.reshape(-1, t.shape[-1])
        key = (t2.shape[0], t.device)
        ones = _ONES.get(key)
        if ones is None:
            ones = _ONES[key] = torch.ones(1, t2.shape[0], dtype=BF, device=t.device)
        try:
            return torch.mm(ones, t2, out_dtype=adt).reshape(-1)
        except (TypeError, RuntimeError, NotImplementedError):
            _COLSUM_MM[0] = False
    return t.sum(dim=tuple(range(t.dim() - 1)), dtype=adt)          # accumulates in adt without materialising a converted copy


def _wgrad(dy, x, adt=torch.float32):
    """dy (B,T,N), x (B,T,K) -> dy^T x (N,K), one GEMM over K = B*T."""
    return (dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])).to(adt)


class StackTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wn, x, c_up, gvec, *weights):
        sh = packing.stack_shape(wn)
        B, T = x.shape[0], x.shape[-1]
        dev = x.device
        L, R, H = sh.layers, sh.R, sh.H
        Hp, Cp = packing._ru(H, 64), (packing._ru(sh.C, 64) if sh.C else 0)
        prep = getattr(wn, "_prep", None)
        if prep is not None:                                # begin_weight_prep ran at the start of the step: just wait for its stream
            lanes, pk = prep["lanes"], prep["pk"]
            torch.cuda.current_stream(dev).wait_stream(prep["stream"])
            ctx.bwp = prep["bwp"]
        else:
            lanes = getattr(wn, "_lanes", None)             # forked by stack_*_train around live_weights; joined here
            pk = packing.pack_bf16(wn, folded=weights, lanes=lanes)
            if lanes is not None:
                lanes.join()
            ctx.bwp = None
        lib = _lib.lib()
        x_is_index = not torch.is_floating_point(x)         # (B,T) classes of a one-hot-input model: the one-hot tensor never exists
        xf = x.detach().long().contiguous() if x_is_index else x.detach().float().contiguous()
        cf = None if c_up is None else c_up.detach().float().contiguous()
        gf = None if gvec is None else gvec.detach().float().contiguous()
        logits = torch.empty(B, sh.O, T, dtype=torch.float32, device=dev)
        x_all = torch.empty(L, B, T, R, dtype=BF, device=dev)
        h_all = torch.empty(L, B, T, Hp, dtype=BF, device=dev)
        c_cl = torch.empty(B, T, Cp, dtype=BF, device=dev) if sh.C else None
        r1 = torch.empty(B, T, sh.S, dtype=BF, device=dev)           # the head's two hidden activations (ReLU outputs)
        r2 = torch.empty(B, T, sh.S, dtype=BF, device=dev)
        # the gate factors (tanh, sigmoid of every layer's gate pre-activations, bf16) kept by the version-4 layer kernel: the
        # backward then skips the recompute of the gate GEMM, a third of its FLOPs (WAE_SAVE_GATE=0: recompute as before)
        gate = None
        if (os.environ.get("WAE_SAVE_GATE", "1") != "0" and tc_backward_supported(sh)
                and lib.wae_stack_gate_save_supported(pk.struct.d) == 1):
            gate = torch.empty(L * B * T * packing._ru(H, 16) * 4, dtype=torch.uint8, device=dev)
        save = _lib.StackSaved(_lib.ptr(x_all), _lib.ptr(h_all), _lib.ptr(c_cl), _lib.ptr(r1), _lib.ptr(r2), _lib.ptr(gate))
        n = lib.wae_stack_workspace_bf16(pk.struct.d, B, T)
        ws = wn._ws.get(n, dev)
        fwd = lib.wae_stack_forward_bf16_save_idx if x_is_index else lib.wae_stack_forward_bf16_save
        _lib.check(fwd(pk.struct, _lib.ptr(xf), _lib.ptr(cf), _lib.ptr(gf), B, T, _lib.ptr(logits), save, _lib.ptr(ws), ws.numel(),
                       _lib.stream_ptr(dev)), "wae_stack_forward_bf16_save")
        ctx.sh, ctx.dil = sh, list(sh.dilations)
        ctx.x_needs_grad = (not x_is_index) and x.requires_grad
        ctx.c_present, ctx.g_present = c_up is not None, gvec is not None
        if not hasattr(wn, "_ws_bwd"):
            wn._ws_bwd = packing.WorkspaceCache()
        ctx.pk, ctx.ws_cache = pk, wn._ws_bwd
        ctx.gate = gate
        ctx.two_ok = lanes is not None            # the two-stream backward relies on the folds having run on lanes
        ctx.save_for_backward(xf, gf, x_all, h_all, c_cl, r1, r2, *[None if w is None else w.detach() for w in weights])
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        xf, gf, x_all, h_all, c_cl, r1, r2, *weights = ctx.saved_tensors
        dxin, dc_up, dgvec, grads = stack_backward(ctx.sh, ctx.dil, xf, gf, x_all, h_all, c_cl, weights, dlogits, ctx.x_needs_grad,
                                                   r1=r1, r2=r2, pk=ctx.pk, ws_cache=ctx.ws_cache, gate=ctx.gate, bwp=getattr(ctx, "bwp", None))
        return (None, dxin, dc_up if ctx.c_present else None, dgvec.to(gf.dtype) if ctx.g_present else None, *grads)


class StackNLLFunction(torch.autograd.Function):
    """Decoder stack + teacher-forced cross-entropy in one autograd node (training side): the forward returns
    the scalar loss of vqwae_train.py:760-766 (mask of ones) from one pass over the logits (wae_nll_sum); the backward writes
    softmax - onehot straight into the (B,T,O) bf16 operand of the tensor-core backward (wae_train_ce_grad) -- no log_softmax
    forward / backward kernels, no (B,O,T) gradient tensor, no transposing cast."""

    @staticmethod
    def forward(ctx, wn, x, c_up, gvec, target, shift, *weights):
        logits = StackTrainFunction.forward(ctx, wn, x, c_up, gvec, *weights)
        B, O, T = logits.shape
        tgt = target.detach().long().contiguous()
        out = torch.zeros(1, dtype=torch.float64, device=logits.device)
        _lib.check(_lib.lib().wae_nll_sum(_lib.ptr(logits), _lib.ptr(tgt), B, O, T, int(shift), _lib.ptr(out), _lib.stream_ptr(logits.device)),
                   "wae_nll_sum")
        ctx.logits, ctx.target, ctx.shift = logits, tgt, int(shift)
        return (out[0] / float(B * (T - shift))).float()

    @staticmethod
    def backward(ctx, dloss):
        xf, gf, x_all, h_all, c_cl, r1, r2, *weights = ctx.saved_tensors
        logits, B, O, T = ctx.logits, *ctx.logits.shape
        if not tc_backward_supported(ctx.sh):
            raise _lib.WaeError("forward_nll needs the tensor-core backward (R, S multiples of 64, G <= 256, O % 16 == 0)")
        dy = torch.empty(B, T, O, dtype=BF, device=logits.device)
        gs = dloss.detach().float().reshape(1).contiguous()
        _lib.check(_lib.lib().wae_train_ce_grad(_lib.ptr(logits), _lib.ptr(ctx.target), B, O, T, ctx.shift, _lib.ptr(gs),
                                                1.0 / float(B * (T - ctx.shift)), _lib.ptr(dy), _lib.stream_ptr(logits.device)),
                   "wae_train_ce_grad")
        dxin, dc_up, dgvec, grads = _stack_backward_tc(ctx.sh, xf, gf, x_all, h_all, c_cl, r1, r2, weights, None, ctx.x_needs_grad,
                                                       ctx.pk, ctx.ws_cache, dy=dy, two_ok=getattr(ctx, "two_ok", False),
                                                       gate=getattr(ctx, "gate", None), bwp=getattr(ctx, "bwp", None))
        return (None, dxin, dc_up if ctx.c_present else None, dgvec.to(gf.dtype) if ctx.g_present else None, None, None, *grads)


def stack_nll_train(wn, x, c_up, gvec, target, shift=1):
    """Scalar teacher-forced NLL with autograd through the tcgen05 forward / backward (StackNLLFunction).  x: (B,Oin,T) float
    or, for a one-hot-input model, the (B,T) integer classes."""
    prep = getattr(wn, "_prep", None)
    if prep is not None:                      # weights prepared on a side stream since the start of the step (begin_weight_prep)
        try:
            return StackNLLFunction.apply(wn, x, c_up, gvec, target, shift, *prep["weights"])
        finally:
            wn._prep = None
    wn._lanes = _lanes_for(wn, x)
    try:
        return StackNLLFunction.apply(wn, x, c_up, gvec, target, shift, *live_weights(wn, wn._lanes))
    finally:
        if wn._lanes is not None:
            wn._lanes.join()
        wn._lanes = None


def tc_backward_supported(sh) -> bool:
    """Shapes wae_stack_backward_bf16 covers (include/wae_b200.h): every preset of the reference with G <= 256."""
    return (sh.R % 64 == 0 and sh.R <= 256 and sh.S % 64 == 0 and sh.S <= 256 and _ru(sh.H, 16) <= 128 and sh.O % 16 == 0
            and sh.O <= 256 and 1 <= sh.kernel_size <= 5)


def _ru(x, m):
    return (x + m - 1) // m * m


def pack_backward_weights(sh, weights, dev):
    """The weights in the transposed K-major forms of the backward's dgrad GEMMs (include/wae_b200.h, wae_stack_bwd): tiny
    tensors, a handful of launches for all layers.  Depends on the weights only, so a training step builds them on its
    weight-preparation stream during the forward (begin_weight_prep) instead of between the loss and the first backward GEMM."""
    L, R, H, S, C, kw = sh.layers, sh.R, sh.H, sh.S, sh.C, sh.kernel_size
    Hh, Hp, Cp = _ru(H, 16), _ru(H, 64), (_ru(C, 64) if C else 0)
    Gp = 2 * Hh
    Gq = _ru(Gp, 64)
    f32 = torch.float32

    def lw(l, k):
        return weights[l * PER_LAYER + k].detach()

    base = L * PER_LAYER
    W3, W4 = weights[base + 2].detach(), weights[base + 4].detach()
    rows = torch.cat([torch.arange(H, device=dev), Hh + torch.arange(H, device=dev)])     # natural gate row -> packed row
    W1 = torch.stack([lw(l, 0) for l in range(L)]).float()                                  # (L,G,R,kw)
    wdx = torch.zeros(L, R, kw, Gq, dtype=f32, device=dev)
    wdx[:, :, :, rows] = W1.permute(0, 2, 3, 1)
    wdx = wdx.reshape(L, R, kw * Gq).to(BF).contiguous()
    Ws = torch.stack([lw(l, 6)[:, :, 0] for l in range(L)]).float()                         # (L,S,H)
    Wo = torch.stack([lw(l, 4)[:, :, 0] for l in range(L)]).float()                         # (L,R,H)
    wdh = torch.zeros(L, Hp, S + R, dtype=f32, device=dev)
    wdh[:, :H, :S] = Ws.transpose(1, 2)
    wdh[:, :H, S:] = Wo.transpose(1, 2)
    wdh = wdh.to(BF).contiguous()
    wct = None
    if C:
        Wc = torch.stack([lw(l, 2)[:, :, 0] for l in range(L)]).float()                     # (L,G,C)
        wct = torch.zeros(Cp, L, Gq, dtype=f32, device=dev)
        wct[:C][:, :, rows] = Wc.permute(2, 0, 1)
        wct = wct.reshape(Cp, L * Gq).to(BF).contiguous()
    w4t = W4[:, :, 0].float().t().to(BF).contiguous()                                        # (S,O)
    w3t = W3[:, :, 0].float().t().to(BF).contiguous()                                        # (S,S)
    return dict(wdx=wdx, wdh=wdh, wct=wct, w4t=w4t, w3t=w3t)


def begin_weight_prep(wn, device):
    """Start the decoder's per-step weight preparation (weight-norm folds, forward packing, the backward's transposed packs) on a
    side stream forked from the current one -- called at the very start of a training step, so that it runs beside the encoder /
    VQ / upsampler instead of after them (it depends on the parameters only).  stack_nll_train / stack_forward_train pick the
    result up from ``wn._prep`` and make the current stream wait for it."""
    device = torch.device(device)
    if device.type != "cuda":
        return None
    prep = packing.wgrad_stream(device, "prep")
    prep.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(prep):
        lanes = _lanes_for(wn, next(wn.parameters()))
        weights = live_weights(wn, lanes)
        pk = packing.pack_bf16(wn, folded=weights, lanes=lanes)            # joins the lanes into the prep stream
        sh = packing.stack_shape(wn)
        bwp = pack_backward_weights(sh, weights, device) if tc_backward_supported(sh) else None
    wn._prep = dict(stream=prep, lanes=lanes, weights=weights, pk=pk, bwp=bwp)
    return wn._prep


def _stack_backward_tc(sh, xf, gf, x_all, h_all, c_cl, r1, r2, weights, dlogits, x_needs_grad, pk, ws_cache, dy=None, two_ok=False,
                       gate=None, bwp=None):
    """The backward on the tensor cores: pack the transposed weights, one call of wae_stack_backward_bf16, scatter the packed
    fp32 gradients to the parameter shapes.  Same return value as stack_backward."""
    lib = _lib.lib()
    L, R, G, H, S, C, O, kw, Gi = sh.layers, sh.R, sh.G, sh.H, sh.S, sh.C, sh.O, sh.kernel_size, sh.Gi
    _, B, T, _ = x_all.shape
    dev = x_all.device
    Hh, Hp, Cp = _ru(H, 16), _ru(H, 64), (_ru(C, 64) if C else 0)
    Gp = 2 * Hh
    Gq, K1p = _ru(Gp, 64), kw * R + Cp
    f32 = torch.float32

    def lw(l, k):
        return weights[l * PER_LAYER + k]

    base = L * PER_LAYER
    Wf, W3, W4 = weights[base], weights[base + 2], weights[base + 4]
    rows = torch.cat([torch.arange(H, device=dev), Hh + torch.arange(H, device=dev)])     # natural gate row -> packed row
    if bwp is None:
        bwp = pack_backward_weights(sh, weights, dev)
    wdx, wdh, wct, w4t, w3t = bwp["wdx"], bwp["wdh"], bwp["wct"], bwp["w4t"], bwp["w3t"]
    # ---- outputs ----
    sizes = dict(dw1=L * Gp * K1p, dwo=L * R * Hp, dws=S * L * Hp, dw3=S * S, dw4=O * S, dgb=L * B * Gp, dbo=L * R, dbs=S, db3=S, db4=O)
    flat = torch.empty(sum(sizes.values()), dtype=f32, device=dev)                          # zeroed by the call
    o, off = {}, 0
    for k, n in sizes.items():
        o[k] = flat[off:off + n]
        off += n
    dc = torch.empty(B, T, Cp, dtype=f32, device=dev) if C else None
    dx0 = torch.empty(B, T, R, dtype=BF, device=dev)
    bw = _lib.StackBwd()
    for name, t in (("wdh", wdh), ("wdx", wdx), ("wct", wct), ("w4t", w4t), ("w3t", w3t), ("x_all", x_all), ("h_all", h_all),
                    ("c_cl", c_cl), ("r1", r1), ("r2", r2), ("gemb", gf if (Gi and gf is not None) else None), ("dc", dc), ("dx0", dx0)):
        setattr(bw, name, _lib.ptr(t))
    for k in sizes:
        setattr(bw, k, o[k].data_ptr())
    bw.dy = _lib.ptr(dy)                                 # (B,T,O) bf16 from wae_train_ce_grad, or None: transpose-cast dlogits
    bw.gate = _lib.ptr(gate)                             # kept gate factors (forward's StackSaved.gate), or None: recompute
    dl = None if dy is not None else dlogits.float().contiguous()
    # two streams (WAE_BWD_STREAMS=1 turns it off): the weight-gradient GEMMs go to a side stream nothing on this stream waits
    # for -- this function returns dc / dx0 / bias gradients in stream order, and the upsampler / VQ / encoder backward that
    # autograd runs next overlaps with the wgrads still draining
    two = two_ok and os.environ.get("WAE_BWD_STREAMS", "2") != "1" and dev.type == "cuda"
    sw = packing.wgrad_stream(dev) if two else None
    if two:
        n = lib.wae_stack_backward_workspace_bf16_2s(pk.struct.d, B, T)
        ws = ws_cache.get(n, dev) if ws_cache is not None else torch.empty(n, dtype=torch.uint8, device=dev)
        sc = packing.wgrad_stream(dev, "bias") if os.environ.get("WAE_BWD_BIAS_STREAM", "1") != "0" else None
        _lib.check(lib.wae_stack_backward_bf16_2s(pk.struct, bw, _lib.ptr(dl), B, T, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev),
                                                  sw.cuda_stream, sc.cuda_stream if sc is not None else None), "wae_stack_backward_bf16_2s")
        if sc is not None:
            torch.cuda.current_stream(dev).wait_stream(sc)       # the column sums (dgb, dbo, ...) are consumed on this stream below
        # Everything the side streams still read or write after this function has returned was allocated on THIS stream: without a
        # note to the caching allocator, the blocks of the saved activations (released by autograd as soon as this node is done), of
        # dy and of the gradient buffer could be handed to the next allocation on this stream -- the upsampler / encoder backward
        # -- while a lagging weight-gradient GEMM is still using them.  (During graph capture the allocator then keeps such blocks
        # until the capture ends.)
        for t in (x_all, h_all, c_cl, r1, r2, dy, flat, xf):
            if t is not None and t.is_cuda:
                t.record_stream(sw)
                if sc is not None:
                    t.record_stream(sc)
    else:
        n = lib.wae_stack_backward_workspace_bf16(pk.struct.d, B, T)
        ws = ws_cache.get(n, dev) if ws_cache is not None else torch.empty(n, dtype=torch.uint8, device=dev)
        _lib.check(lib.wae_stack_backward_bf16(pk.struct, bw, _lib.ptr(dl), B, T, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                   "wae_stack_backward_bf16")
    # ---- scatter to the parameter shapes ----
    grads = [None] * len(weights)
    dw1 = o["dw1"].view(L, Gp, K1p)[:, rows]                                                # (L,G,K1p) natural gate rows
    dW1 = dw1[:, :, :kw * R].reshape(L, G, kw, R).permute(0, 1, 3, 2)                       # (L,G,R,kw)
    dWc = dw1[:, :, kw * R: kw * R + C] if C else None
    dgb = o["dgb"].view(L, B, Gp)[:, :, rows]                                               # (L,B,G)
    dwo, dws = o["dwo"].view(L, R, Hp), o["dws"].view(S, L, Hp)
    dgvec = torch.zeros_like(gf, dtype=f32) if gf is not None else None
    if Gi and gf is not None:
        Wg = torch.stack([lw(l, 3)[:, :, 0] for l in range(L)]).float()                     # (L,G,Gi)
        dWg = (dgb.unsqueeze(-1) * gf.float()[None, :, None, :]).sum(1)                     # (L,G,Gi): tiny, element-wise
        dgvec = (dgb.unsqueeze(-1) * Wg.unsqueeze(1)).sum(2).sum(0)                         # (B,Gi)
    db1 = dgb.sum(1)
    for l in range(L):
        k = l * PER_LAYER
        grads[k + 0] = dW1[l]
        if lw(l, 1) is not None:
            grads[k + 1] = db1[l]
        if C:
            grads[k + 2] = dWc[l].unsqueeze(-1)
        if Gi and gf is not None and lw(l, 3) is not None:
            grads[k + 3] = dWg[l].unsqueeze(-1)
        if l < L - 1:                       # the last layer's residual branch is dead: no gradient, as under autograd in the reference
            grads[k + 4] = dwo[l, :, :H].unsqueeze(-1)
            if lw(l, 5) is not None:
                grads[k + 5] = o["dbo"].view(L, R)[l]
        grads[k + 6] = dws[:, l, :H].unsqueeze(-1)
        if lw(l, 7) is not None:
            grads[k + 7] = o["dbs"]
    grads[base + 2] = o["dw3"].view(S, S).unsqueeze(-1)
    if weights[base + 3] is not None:
        grads[base + 3] = o["db3"]
    grads[base + 4] = o["dw4"].view(O, S).unsqueeze(-1)
    if weights[base + 5] is not None:
        grads[base + 5] = o["db4"]
    # ---- first conv: x0 = Wf x + bf.  dWf = dx0^T x^T as one more MN-major wgrad over the transposed-cast input ----
    x_is_index = not torch.is_floating_point(xf)
    Oin = sh.Oin if x_is_index else xf.shape[1]
    st = _lib.stream_ptr(dev)
    dWf = torch.zeros(R, _ru(Oin, 16), dtype=f32, device=dev)
    if x_is_index:                                                                           # (B,T) classes: bf16 one-hot rows in one launch
        xT = torch.empty(B, T, _ru(Oin, 16), dtype=BF, device=dev)
        _lib.check(lib.wae_onehot_bf16(_lib.ptr(xf), B * T, _ru(Oin, 16), _lib.ptr(xT), st), "wae_onehot_bf16")
        _lib.check(lib.wae_gemm_bf16_nt(_lib.ptr(dx0), _lib.ptr(xT), _lib.ptr(dWf), R, _ru(Oin, 16), B * T, st), "wae_gemm_bf16_nt")
        grads[base] = dWf[:, :Oin].unsqueeze(-1)
    elif Oin % 16 == 0 and B <= 65535:
        xT = torch.empty(B, T, Oin, dtype=BF, device=dev)                                   # exact for the one-hot input of every preset
        _lib.check(lib.wae_train_transpose_cast(_lib.ptr(xf), B, Oin, T, _lib.ptr(xT), st), "wae_train_transpose_cast")
        _lib.check(lib.wae_gemm_bf16_nt(_lib.ptr(dx0), _lib.ptr(xT), _lib.ptr(dWf), R, Oin, B * T, st), "wae_gemm_bf16_nt")
        grads[base] = dWf[:, :Oin].unsqueeze(-1)
    else:                                                                                    # scalar input (Oin = 1): a weighted column sum
        grads[base] = (dx0.float() * xf.transpose(1, 2).float()).sum((0, 1)).reshape(R, Oin, 1) if Oin == 1 else \
            torch.einsum("btr,bot->ro", dx0.float(), xf.float()).unsqueeze(-1)
    if weights[base + 1] is not None:
        dbf = torch.zeros(R, dtype=f32, device=dev)
        _lib.check(lib.wae_colsum_bf16(_lib.ptr(dx0), B, T, R, 0, _lib.ptr(dbf), st), "wae_colsum_bf16")
        grads[base + 1] = dbf
    dxin = None
    if x_needs_grad:
        dxin = (dx0.float() @ Wf[:, :, 0].float()).transpose(1, 2).contiguous()
    dc_up = dc[..., :C].transpose(1, 2).contiguous() if C else None
    # the copies that bring permuted views into the parameter shapes: layer l on lane l (the weight-norm backward that consumes
    # them runs there too); joined before returning, autograd assumes this node's outputs live on its own stream
    # Two-stream backward: the tensors that come out of the wgrad stream (dW1, dWc, dWo, dWs, dW3, dW4) are only touched on
    # lanes that waited for it, and layer l's lane is the stream its weight-norm fold ran on in the forward, i.e. the stream
    # autograd runs the fold's backward on -- so nothing on THIS stream ever waits for the wgrads.  Bias-like gradients (column
    # sums, made on this stream) stay here: their accumulators live on this stream.
    n_l = int(os.environ.get("WAE_PREP_LANES", "8")) or 1
    lanes = packing.Lanes(dev, n_l, also_wait=(sw,) if two else ()) if dev.type == "cuda" else packing._NoLanes()
    out = []
    for i, (g, w) in enumerate(zip(grads, weights)):
        if g is None or w is None:
            out.append(g)
            continue
        weight_like = (i % PER_LAYER in (0, 2, 3, 4, 6)) if i < base else ((i - base) % 2 == 0)      # the weight-normed tensors
        if weight_like:
            with lanes.lane(i // PER_LAYER if i < base else L):
                if two and g.is_cuda:
                    g.record_stream(torch.cuda.current_stream(dev))      # the lane reads a block that belongs to this node's stream
                out.append(g.to(w.dtype).reshape(w.shape))
        else:
            out.append(g.to(w.dtype).reshape(w.shape))
    if not two:
        lanes.join()
    return dxin, dc_up, dgvec, out


def stack_backward(sh, dil, xf, gf, x_all, h_all, c_cl, weights, dlogits, x_needs_grad=False, cdt=BF, adt=torch.float32,
                   r1=None, r2=None, pk=None, ws_cache=None, gate=None, bwp=None):
    """Hand-derived backward of the decoder stack on saved channels-last activations.  cdt: GEMM operand dtype (bf16 on the
    GPU), adt: accumulation / element-wise dtype.  tests/test_host_cpu.py runs it in float64 against torch autograd.  With
    bf16 CUDA tensors, the head's saved hidden activations (r1, r2) and the packed forward weights (pk) it runs on the
    tensor-core kernels (_stack_backward_tc); shapes those do not cover keep the library-GEMM composite below."""
    if cdt == BF and x_all.is_cuda and r1 is not None and r2 is not None and pk is not None and tc_backward_supported(sh):
        return _stack_backward_tc(sh, xf, gf, x_all, h_all, c_cl, r1, r2, weights, dlogits, x_needs_grad, pk, ws_cache, gate=gate, bwp=bwp)
    if True:
        L, R, G, H, S, C, O, kw = sh.layers, sh.R, sh.G, sh.H, sh.S, sh.C, sh.O, sh.kernel_size
        _, B, T, _ = x_all.shape
        Hp = h_all.shape[-1]
        Cp = c_cl.shape[-1] if c_cl is not None else 0
        scale = math.sqrt(1.0 / L)
        rs = math.sqrt(0.5)
        grads = [None] * len(weights)

        def lw(l, k):
            return weights[l * PER_LAYER + k]

        base = L * PER_LAYER
        Wf, W3, W4 = weights[base], weights[base + 2], weights[base + 4]
        b3 = weights[base + 3]

        # ---- head: logits = W4 relu(W3 relu(s) + b3) + b4,  s = (sum_l Ws_l h_l + bs_l) * sqrt(1/L) ----
        dY = torch.empty(dlogits.shape[0], dlogits.shape[2], dlogits.shape[1], dtype=cdt, device=dlogits.device)
        if cdt == BF and dlogits.is_cuda and dlogits.dtype == torch.float32 and O % 2 == 0 and dlogits.shape[0] <= 65535:
            dl = dlogits.contiguous()                                                      # (B,T,O): tiled transpose + cast, one launch
            _lib.check(_lib.lib().wae_train_transpose_cast(_lib.ptr(dl), dl.shape[0], O, dl.shape[2], _lib.ptr(dY),
                                                           _lib.stream_ptr(dl.device)), "wae_train_transpose_cast")
        else:
            dY.copy_(dlogits.transpose(1, 2))                                              # (B,T,O): transpose + cast in one pass
        Hcat = h_all.permute(1, 2, 0, 3).reshape(B, T, L * Hp)                            # (B,T,L*Hp)
        Wscat = torch.cat([F.pad(lw(l, 6)[:, :, 0], (0, Hp - H)) for l in range(L)], dim=1).to(cdt)   # (S, L*Hp)
        bs_sum = torch.zeros(S, dtype=adt, device=dY.device)
        for l in range(L):
            if lw(l, 7) is not None:
                bs_sum = bs_sum + lw(l, 7).to(adt)
        s = (_mm_acc(Hcat, Wscat.t(), adt) + bs_sum) * scale
        r1 = torch.relu(s).to(cdt)
        p2 = _mm_acc(r1, W3[:, :, 0].to(cdt).t(), adt) + (b3.to(adt) if b3 is not None else 0.0)
        r2 = torch.relu(p2).to(cdt)
        grads[base + 4] = _wgrad(dY, r2, adt).unsqueeze(-1)
        if weights[base + 5] is not None:
            grads[base + 5] = _colsum(dY, adt)
        dp2 = ((dY @ W4[:, :, 0].to(cdt)).to(adt) * (p2 > 0)).to(cdt)
        grads[base + 2] = _wgrad(dp2, r1, adt).unsqueeze(-1)
        if b3 is not None:
            grads[base + 3] = _colsum(dp2, adt)
        dS = ((dp2 @ W3[:, :, 0].to(cdt)).to(adt) * (s > 0) * scale).to(cdt)               # d loss / d (skip sum), the same for every layer
        del s, r1, p2, r2, dp2
        dWscat = _wgrad(dS, Hcat, adt)                                                         # (S, L*Hp): every layer's skip wgrad in one GEMM
        dbs = _colsum(dS, adt)
        dHskip = dS @ Wscat                                                               # (B,T,L*Hp)
        del Hcat

        # ---- residual layers, last to first ----
        # On the GPU (bf16) the gathers and element-wise chains between the GEMMs are three CUDA kernels (csrc/wn_train.cu);
        # the float64 CPU check of the same derivation runs them as torch expressions.
        fused = (cdt == BF and x_all.is_cuda)
        lib = _lib.lib() if fused else None
        st = _lib.stream_ptr(x_all.device) if fused else None
        dev = dS.device
        dx = None                                    # fused: bf16 d loss / d x_{l+1} ALREADY scaled by sqrt(.5) (= dxo); else unscaled, adt
        dC = torch.zeros(B, T, C, dtype=adt, device=dev) if C else None
        dgvec = torch.zeros_like(gf, dtype=adt) if gf is not None else None
        K = kw * R + Cp
        for l in reversed(range(L)):
            d = dil[l]
            W1, b1, Wc, Wg, Wo = lw(l, 0), lw(l, 1), lw(l, 2), lw(l, 3), lw(l, 4)
            grads[l * PER_LAYER + 6] = dWscat[:, l * Hp: l * Hp + H].unsqueeze(-1).contiguous()
            if lw(l, 7) is not None:
                grads[l * PER_LAYER + 7] = dbs
            dh_skip = dHskip[..., l * Hp: l * Hp + H]                       # view, row stride L*Hp
            h_l = h_all[l][..., :H]
            dh_res = None
            if dx is not None:
                dxo = dx if fused else (dx * rs).to(cdt)
                dh_res = dxo @ Wo[:, :, 0].to(cdt)                          # (B,T,H)
                grads[l * PER_LAYER + 4] = _wgrad(dxo, h_l, adt).unsqueeze(-1)
                if lw(l, 5) is not None:
                    grads[l * PER_LAYER + 5] = _colsum(dxo, adt)
            else:                         # last layer: its residual branch is dead -- no gradient, as under autograd in the reference
                dxo = None
            wparts = [W1[:, :, j] for j in range(kw)]
            if C:
                wparts.append(F.pad(Wc[:, :, 0], (0, Cp - C)))
            W1cat = torch.cat(wparts, dim=1).to(cdt)                        # (G,K)
            gb = b1.to(adt) if b1 is not None else torch.zeros(G, dtype=adt, device=dev)
            gb = gb[None, :].expand(B, G)
            if Wg is not None and gf is not None:
                gb = gb + gf.to(adt) @ Wg[:, :, 0].to(adt).t()
            # recompute the gate pre-activations from the saved layer input
            X = x_all[l]
            if fused:
                Xcat = torch.empty(B, T, K, dtype=BF, device=dev)
                _lib.check(lib.wae_train_im2col(_lib.ptr(X), _lib.ptr(c_cl), B, T, R, Cp, kw, d, _lib.ptr(Xcat), st), "wae_train_im2col")
                z = Xcat @ W1cat.t()                                        # (B,T,G) bf16, bias added inside the gate kernel
                dz = torch.empty(B, T, G, dtype=BF, device=dev)
                dgb = torch.zeros(B, G, dtype=torch.float32, device=dev)
                gbc = gb.float().contiguous()
                _lib.check(lib.wae_train_gate_bwd(_lib.ptr(z), _lib.ptr(gbc), dh_skip.data_ptr(), L * Hp, _lib.ptr(dh_res), B, T, H,
                                                  _lib.ptr(dz), _lib.ptr(dgb), st), "wae_train_gate_bwd")
                del z
            else:
                parts = [_shift(X, (kw - 1 - j) * d) for j in range(kw)] + ([c_cl] if C else [])
                Xcat = torch.cat(parts, dim=-1)                             # (B,T,K)
                z = (Xcat @ W1cat.t()).to(adt) + gb[:, None, :]
                th, sg = torch.tanh(z[..., :H]), torch.sigmoid(z[..., H:])
                dhf = dh_skip.to(adt) + (dh_res.to(adt) if dh_res is not None else 0.0)
                dz = torch.cat([dhf * sg * (1 - th * th), dhf * th * sg * (1 - sg)], dim=-1)
                del z, th, sg, dhf
                dgb = dz.sum(1)                                             # (B,G)
                dz = dz.to(cdt)
            if b1 is not None:
                grads[l * PER_LAYER + 1] = dgb.sum(0)
            if Wg is not None and gf is not None:
                grads[l * PER_LAYER + 3] = (dgb.to(adt).t() @ gf.to(adt)).unsqueeze(-1)
                dgvec += dgb.to(adt) @ Wg[:, :, 0].to(adt)
            dW1cat = _wgrad(dz, Xcat, adt)                                  # (G,K)
            grads[l * PER_LAYER + 0] = torch.stack([dW1cat[:, j * R:(j + 1) * R] for j in range(kw)], dim=-1)
            if C:
                grads[l * PER_LAYER + 2] = dW1cat[:, kw * R: kw * R + C].unsqueeze(-1).contiguous()
            dXcat = dz @ W1cat                                              # (B,T,K)
            del Xcat, dz
            if fused:
                dxn = torch.empty(B, T, R, dtype=BF, device=dev)
                _lib.check(lib.wae_train_dx_accum(_lib.ptr(dXcat), _lib.ptr(dxo), B, T, R, C, Cp, kw, d, rs if l > 0 else 1.0,
                                                  _lib.ptr(dxn), _lib.ptr(dC), st), "wae_train_dx_accum")
            else:
                dxn = dxo.to(adt) if dxo is not None else 0.0
                for j in range(kw):
                    dxn = dxn + _unshift(dXcat[..., j * R:(j + 1) * R], (kw - 1 - j) * d).to(adt)
                if C:
                    dC += dXcat[..., kw * R: kw * R + C].to(adt)
            dx = dxn
            del dXcat

        # ---- first conv: x0 = Wf x + bf ----
        dx0 = dx.to(cdt)
        xb = xf.to(cdt)                                                                    # (B,Oin,T); exact for one-hot input
        grads[base] = torch.bmm(xb, dx0).to(adt).sum(0).t().unsqueeze(-1).contiguous()    # (R,Oin,1)
        if weights[base + 1] is not None:
            grads[base + 1] = _colsum(dx0, adt)
        dxin = None
        if x_needs_grad:
            dxin = (dx0 @ Wf[:, :, 0].to(cdt)).to(adt).transpose(1, 2).contiguous()
        dc_up = dC.transpose(1, 2).contiguous() if C else None
        grads = [g if g is None or w is None else g.to(w.dtype).reshape(w.shape) for g, w in zip(grads, weights)]
        return dxin, dc_up, dgvec, grads


def stack_forward_train(wn, x, c_up, gvec):
    """(B,O,T) logits with autograd through the tcgen05 forward + GEMM backward.  x one-hot / dense (B,Oin,T); c_up
    (B,C,T) already upsampled; gvec (B,Gi) speaker vectors (with autograd history back to the embedding)."""
    wn._lanes = _lanes_for(wn, x)
    try:
        return StackTrainFunction.apply(wn, x, c_up, gvec, *live_weights(wn, wn._lanes))
    finally:
        if wn._lanes is not None:
            wn._lanes.join()
        wn._lanes = None
