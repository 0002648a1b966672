"""Parameter containers with the reference's names/initialisation (wavenet_vocoder/modules.py).

The arithmetic of ``ResidualConv1dGLU`` (modules.py:115-163) lives in the fused CUDA kernels
(csrc/wn_stack_f32.cu, wn_stack_bf16.cu, wn_ar.cu); these classes keep constructor signature,
sub-module names and the old-style weight-norm parametrisation (``weight_g``/``weight_v``) so that
reference checkpoints load unchanged.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from . import conv


def Conv1d(in_channels, out_channels, kernel_size, dropout=0, **kwargs):
    m = conv.Conv1d(in_channels, out_channels, kernel_size, **kwargs)
    nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)
    return nn.utils.weight_norm(m)


def Embedding(num_embeddings, embedding_dim, padding_idx, std=0.01):
    m = nn.Embedding(num_embeddings, embedding_dim, padding_idx=padding_idx)
    m.weight.data.normal_(0, std)
    return m


def Conv1d1x1(in_channels, out_channels, bias=True):
    return Conv1d(in_channels, out_channels, kernel_size=1, padding=0, dilation=1, bias=bias)


class ResidualConv1dGLU(nn.Module):
    """One gated residual layer: dilated causal conv -> (+c, +g) -> tanh*sigmoid -> skip / residual 1x1s."""

    def __init__(self, residual_channels, gate_channels, kernel_size, skip_out_channels=None,
                 cin_channels=-1, gin_channels=-1, dropout=1 - 0.95, padding=None, dilation=1,
                 causal=True, bias=True, *args, **kwargs):
        super().__init__()
        self.dropout = dropout
        if skip_out_channels is None:
            skip_out_channels = residual_channels
        if padding is None:
            padding = (kernel_size - 1) * dilation if causal else (kernel_size - 1) // 2 * dilation
        self.causal = causal
        self.conv = Conv1d(residual_channels, gate_channels, kernel_size, padding=padding,
                           dilation=dilation, bias=bias, *args, **kwargs)
        self.conv1x1c = Conv1d1x1(cin_channels, gate_channels, bias=False) if cin_channels > 0 else None
        self.conv1x1g = Conv1d1x1(gin_channels, gate_channels, bias=False) if gin_channels > 0 else None
        gate_out_channels = gate_channels // 2
        self.conv1x1_out = Conv1d1x1(gate_out_channels, residual_channels, bias=bias)
        self.conv1x1_skip = Conv1d1x1(gate_out_channels, skip_out_channels, bias=bias)

    def forward(self, x, c=None, g=None):
        """Differentiable single-layer evaluation with torch ops: the reference composite (modules.py:115-163).

        Test infrastructure / fp32 debugging only -- reached solely through ``WaveNet.train_impl = "autograd"``.  Inference and
        the default training path go through WaveNet.forward -> libwae_b200 (wn_stack_*.cu, wn_bwd.cu) and never get here.
        """
        T = x.size(-1)
        z = self.conv(F.dropout(x, p=self.dropout, training=self.training))
        z = z[:, :, :T] if self.causal else z
        if c is not None:
            z = z + self.conv1x1c(c)
        if g is not None:
            z = z + self.conv1x1g(g)
        a, b = z.split(z.size(1) // 2, dim=1)
        h = torch.tanh(a) * torch.sigmoid(b)
        return (self.conv1x1_out(h) + x) * math.sqrt(0.5), self.conv1x1_skip(h)

    def incremental_forward(self, x, c=None, g=None):
        raise NotImplementedError(
            "per-layer incremental_forward is fused into WaveNet.incremental_forward (libwae_b200 wae_ar_generate)")

    def clear_buffer(self):
        for m in (self.conv, self.conv1x1_out, self.conv1x1_skip, self.conv1x1c, self.conv1x1g):
            if m is not None:
                m.clear_buffer()
