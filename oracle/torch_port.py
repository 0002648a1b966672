"""ORACLE / CPU BASELINE (test infrastructure): the reference's hot path re-expressed with the SAME ATen calls
the reference makes (F.conv1d, F.linear, torch.addmm, tanh/sigmoid ...), as plain functions over folded
weights.  The reference is pure Python over PyTorch and cannot travel to the GPU box (/root/reference does not
exist there), so this port is what bench.py times on the host cores as ``cpu_baseline`` / ``--impl reference``
(kind "port").  It is the reference's best case: weight norm is folded once (= after make_generation_fast_,
wavenet.py:358-364), the per-step Python overhead of nn.Module dispatch is gone.

    stack_forward         wavenet.py:203-212, modules.py:115-163   (op for op: padded dilated conv1d, slice, 1x1s)
    ar_generate           wavenet.py:299-339, conv.py:17-46        (clone-shift input buffers, strided gather, F.linear)
    vq_forward            vector_quantization.py:21-49             (addmm, argmin, one-hot scatter, one-hot @ codebook)

Checked against the reference's golden vectors in tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import math

import torch
from torch.nn import functional as F


def params_from_numpy(p: dict, device="cpu") -> dict:
    """oracle.wavenet_oracle.extract_params(...) -> torch tensors."""
    def t(a):
        return None if a is None else torch.as_tensor(a, dtype=torch.float32, device=device)
    out = {k: t(v) for k, v in p.items() if k not in ("layers", "up")}
    out["layers"] = [{k: (v if k == "dilation" else t(v)) for k, v in lay.items()} for lay in p["layers"]]
    if "up" in p:
        out["up"] = [t(u) for u in p["up"]]
    return out


def upsample(p: dict, c: torch.Tensor) -> torch.Tensor:
    c = F.conv1d(c, p["conv_in"])
    c = c.unsqueeze(1)
    for k in p["up"]:
        s = (k.numel() - 1) // 2
        c = F.interpolate(c, scale_factor=(1, s), mode="nearest")
        c = F.conv2d(c, k.view(1, 1, 1, -1), padding=(0, s))
    return c.squeeze(1)


def stack_forward(p: dict, x: torch.Tensor, c_up, gvec) -> torch.Tensor:
    B, _, T = x.shape
    g_bct = None if gvec is None else gvec.unsqueeze(-1).expand(B, -1, T).contiguous()
    x = F.conv1d(x, p["wf"].unsqueeze(-1), p["bf"])
    skips = 0
    for lay in p["layers"]:
        w, d = lay["w"], lay["dilation"]
        residual = x
        z = F.conv1d(x, w, lay["b"], padding=(w.shape[2] - 1) * d, dilation=d)[:, :, :T]
        a, b = z.split(z.size(1) // 2, dim=1)
        if c_up is not None:
            ca, cb = F.conv1d(c_up, lay["wc"].unsqueeze(-1)).split(z.size(1) // 2, dim=1)
            a, b = a + ca, b + cb
        if g_bct is not None:
            ga, gb = F.conv1d(g_bct, lay["wg"].unsqueeze(-1)).split(z.size(1) // 2, dim=1)
            a, b = a + ga, b + gb
        h = torch.tanh(a) * torch.sigmoid(b)
        s = F.conv1d(h, lay["ws"].unsqueeze(-1), lay["bs"])
        x = (F.conv1d(h, lay["wo"].unsqueeze(-1), lay["bo"]) + residual) * math.sqrt(0.5)
        skips = skips + s
    skips = skips * math.sqrt(1.0 / len(p["layers"]))
    x = F.relu(skips)
    x = F.relu(F.conv1d(x, p["w3"].unsqueeze(-1), p["b3"]))
    return F.conv1d(x, p["w4"].unsqueeze(-1), p["b4"])


def ar_generate(p: dict, T: int, c_btc, gvec, init: torch.Tensor, test_inputs=None, sample=None) -> torch.Tensor:
    """The reference's incremental algorithm: per layer a (B, (kw-1)d+1, R) buffer shifted by clone every step."""
    B = init.shape[0]
    lin = [lay["w"].permute(0, 2, 1).reshape(lay["w"].shape[0], -1).contiguous() for lay in p["layers"]]
    bufs = [None] * len(p["layers"])
    outs = []
    cur = init
    for t in range(T):
        if test_inputs is not None and t < test_inputs.shape[1]:
            cur = test_inputs[:, t]
        x = F.linear(cur, p["wf"], p["bf"])
        skips = 0
        for i, lay in enumerate(p["layers"]):
            kw, d = lay["w"].shape[2], lay["dilation"]
            if bufs[i] is None:
                bufs[i] = x.new_zeros(B, kw + (kw - 1) * (d - 1), x.shape[1])
            else:
                bufs[i][:, :-1, :] = bufs[i][:, 1:, :].clone()
            bufs[i][:, -1, :] = x
            inp = bufs[i][:, 0::d, :].contiguous() if d > 1 else bufs[i]
            z = F.linear(inp.view(B, -1), lin[i], lay["b"])
            a, b = z.split(z.size(-1) // 2, dim=-1)
            if c_btc is not None:
                ca, cb = F.linear(c_btc[:, t], lay["wc"]).split(z.size(-1) // 2, dim=-1)
                a, b = a + ca, b + cb
            if gvec is not None:
                ga, gb = F.linear(gvec, lay["wg"]).split(z.size(-1) // 2, dim=-1)
                a, b = a + ga, b + gb
            h = torch.tanh(a) * torch.sigmoid(b)
            s = F.linear(h, lay["ws"], lay["bs"])
            x = (F.linear(h, lay["wo"], lay["bo"]) + x) * math.sqrt(0.5)
            skips = skips + s
        y = F.relu(skips * math.sqrt(1.0 / len(p["layers"])))
        y = F.linear(F.relu(F.linear(y, p["w3"], p["b3"])), p["w4"], p["b4"])
        out = y if sample is None else sample(t, y)
        outs.append(out)
        cur = out
    return torch.stack(outs, dim=1)


def vq_forward(x: torch.Tensor, codebook: torch.Tensor, beta: float = 0.25):
    """VectorQuantize.forward, op for op (including the (N,K) one-hot and the one-hot @ codebook gather)."""
    inputs = x.permute(0, 2, 1).contiguous()
    B, T, D = inputs.shape
    flat = inputs.view(-1, D)
    in_sqr = torch.sum(flat ** 2, dim=1, keepdim=True)
    e_sqr = torch.sum(codebook ** 2, dim=1)
    dis = torch.addmm(e_sqr + in_sqr, flat, codebook.t(), alpha=-2.0, beta=1.0)
    ind = torch.argmin(dis, dim=1).unsqueeze(1)
    enc = torch.zeros(B * T, codebook.shape[0]).scatter_(1, ind, 1)
    quant = torch.matmul(enc, codebook).view(B, T, D)
    loss = beta * torch.mean((quant - inputs) ** 2) + torch.mean((quant - inputs) ** 2)
    quant = inputs + (quant - inputs)
    avg = torch.mean(enc, dim=0)
    perp = torch.exp(-torch.sum(avg * torch.log(avg + 1e-10)))
    return quant.permute(0, 2, 1).contiguous(), loss, perp, ind.view(B, T)
