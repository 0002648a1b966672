"""Conditioning upsamplers with the reference's names and parameters (wavenet_vocoder/upsample.py).

Inference runs each stage (nearest stretch by s + 1x(2s+1) smoothing conv, upsample.py:18-20,42) as one
CUDA kernel (wae_upsample_stage); under autograd the same stage is expressed with torch ops.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .. import _lib


class Stretch2d(nn.Module):
    def __init__(self, x_scale, y_scale, mode="nearest"):
        super().__init__()
        self.x_scale, self.y_scale, self.mode = x_scale, y_scale, mode

    def forward(self, x):
        return F.interpolate(x, scale_factor=(self.y_scale, self.x_scale), mode=self.mode)


class UpsampleStageFunction(torch.autograd.Function):
    """One stage (nearest stretch by s + 1 x (2s+1) smoothing conv, upsample.py:37-49) with both directions on this library's
    kernels: wae_upsample_stage forward, wae_upsample_stage_backward for d input and d filter.  The training step of the
    reference differentiates F.interpolate + Conv2d(1, 1, (1, 2s+1)) at up to sample rate here (cuDNN: layout conversions and
    single-channel wgrad kernels, ~1.3 ms of a 7 ms step at 8 x 7680 samples)."""

    @staticmethod
    def forward(ctx, x, w, s):
        B, C, Tin = x.shape
        x = x.contiguous().float()
        w = w.contiguous().float()
        out = torch.empty(B, C, Tin * s, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().wae_upsample_stage(_lib.ptr(x), B * C, Tin, s, _lib.ptr(w), _lib.ptr(out), _lib.stream_ptr(x.device)),
                   "wae_upsample_stage")
        ctx.save_for_backward(x, w)
        ctx.s = s
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, C, Tin = x.shape
        s, lib = ctx.s, _lib.lib()
        dy = dy.contiguous().float()
        din = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        ws = torch.empty(lib.wae_upsample_stage_backward_workspace(B * C, s), dtype=torch.uint8, device=x.device)
        _lib.check(lib.wae_upsample_stage_backward(_lib.ptr(dy), _lib.ptr(x), B * C, Tin, s, _lib.ptr(w), _lib.ptr(din), _lib.ptr(dw),
                                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr(x.device)), "wae_upsample_stage_backward")
        return din, dw, None


class UpsampleNetwork(nn.Module):
    def __init__(self, upsample_scales, upsample_activation="none", upsample_activation_params={},
                 mode="nearest", freq_axis_kernel_size=1, cin_pad=0, cin_channels=80):
        super().__init__()
        self.up_layers = nn.ModuleList()
        self.indent = cin_pad * int(np.prod(upsample_scales))
        self.scales = [int(s) for s in upsample_scales]
        self.freq_axis_kernel_size = freq_axis_kernel_size
        self.has_activation = upsample_activation != "none"
        self.mode = mode
        for scale in self.scales:
            k_size = (freq_axis_kernel_size, scale * 2 + 1)
            conv = nn.Conv2d(1, 1, kernel_size=k_size, padding=((freq_axis_kernel_size - 1) // 2, scale), bias=False)
            conv.weight.data.fill_(1.0 / np.prod(k_size))
            self.up_layers.append(Stretch2d(scale, 1, mode))
            self.up_layers.append(nn.utils.weight_norm(conv))
            if self.has_activation:
                self.up_layers.append(getattr(nn, upsample_activation)(**upsample_activation_params))

    def _kernel_ok(self, c):
        return (c.is_cuda and not (torch.is_grad_enabled() and (c.requires_grad or any(p.requires_grad for p in self.parameters())))
                and self.freq_axis_kernel_size == 1 and not self.has_activation and self.mode == "nearest")

    def _train_kernel_ok(self, c):
        """Differentiable kernel path (``train_impl`` = "kernels", the default; "autograd" keeps the torch composite)."""
        return (c.is_cuda and c.dim() == 3 and c.dtype == torch.float32 and getattr(self, "train_impl", "kernels") == "kernels"
                and self.freq_axis_kernel_size == 1 and not self.has_activation and self.mode == "nearest"
                and all(1 <= s <= 128 for s in self.scales) and c.numel() > 0)

    def forward(self, c, defer_last=False):
        """defer_last=True (inference, CUDA kernels usable, no indent): run every stage but the last and return
        ``(frames, filter, scale)`` of the last one, which wae_stack_forward_bf16_up fuses into the decoder stack."""
        if defer_last and not (self._kernel_ok(c) and self.indent == 0 and len(self.scales) >= 1):
            return None
        if self._kernel_ok(c):
            from ..packing import folded_weight
            B, C, T = c.shape
            cur = c.contiguous().float()
            convs = [m for m in self.up_layers if isinstance(m, nn.Conv2d)]
            for i, (s, conv) in enumerate(zip(self.scales, convs)):
                w = folded_weight(conv).float().reshape(-1).contiguous()
                if defer_last and i == len(self.scales) - 1:
                    return cur, w, s
                out = torch.empty(B, C, cur.shape[-1] * s, dtype=torch.float32, device=c.device)
                _lib.check(_lib.lib().wae_upsample_stage(_lib.ptr(cur), B * C, cur.shape[-1], s, _lib.ptr(w),
                                                         _lib.ptr(out), _lib.stream_ptr(c.device)), "wae_upsample_stage")
                cur = out
            c = cur
        elif self._train_kernel_ok(c):
            # under autograd on the GPU: the same stage kernel with its own backward (UpsampleStageFunction); the weight-norm
            # fold stays a torch op, so the gradient reaches weight_g / weight_v as in the reference
            convs = [m for m in self.up_layers if isinstance(m, nn.Conv2d)]
            cur = c
            for s, conv in zip(self.scales, convs):
                if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
                    w = torch._weight_norm(conv.weight_v, conv.weight_g, 0)
                else:
                    w = conv.weight
                cur = UpsampleStageFunction.apply(cur, w.reshape(-1), s)
            c = cur
        else:
            c = c.unsqueeze(1)
            for f in self.up_layers:
                c = f(c)
            c = c.squeeze(1)
        if self.indent > 0:
            c = c[:, :, self.indent:-self.indent]
        return c


class ConvInUpsampleNetwork(nn.Module):
    def __init__(self, upsample_scales, upsample_activation="none", upsample_activation_params={},
                 mode="nearest", freq_axis_kernel_size=1, cin_pad=0, cin_channels=80):
        super().__init__()
        self.conv_in = nn.Conv1d(cin_channels, cin_channels, kernel_size=2 * cin_pad + 1, bias=False)
        self.upsample = UpsampleNetwork(upsample_scales, upsample_activation, upsample_activation_params,
                                        mode, freq_axis_kernel_size, cin_pad=0, cin_channels=cin_channels)

    def forward(self, c, defer_last=False):
        # conv_in is a frame-rate C x C conv (1x1 for cin_pad=0).  cuDNN convolutions default to TF32 on sm_80+
        # (torch.backends.cudnn.allow_tf32), which costs ~1e-3 relative accuracy against the reference's fp32 CPU
        # path, so it is evaluated as a plain fp32 matmul / with TF32 disabled.
        w = self.conv_in.weight
        if w.shape[2] == 1:
            c = torch.matmul(w[:, :, 0], c)
        else:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                c = self.conv_in(c)
        return self.upsample(c, defer_last=defer_last) if defer_last else self.upsample(c)
