"""Conv1d with the reference's constructor and state_dict (wavenet_vocoder/conv.py:7-16).

The reference implements autoregressive synthesis module by module (Conv1d.incremental_forward with
a shift buffer per layer, conv.py:17-46).  Here synthesis is ONE fused kernel driven from
WaveNet.incremental_forward, so this class is only a parameter container; ``clear_buffer`` is kept
because callers invoke it.
"""
from torch import nn


class Conv1d(nn.Conv1d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.clear_buffer()

    def incremental_forward(self, input):
        if self.training:
            raise RuntimeError("incremental_forward only supports eval mode")
        raise NotImplementedError(
            "per-module incremental_forward is fused into WaveNet.incremental_forward (libwae_b200 wae_ar_generate)")

    def clear_buffer(self):
        self.input_buffer = None
