"""One VQ-WAE training step as vqwae_train.py:745-790 runs it (BASELINE configs[2]): encoder -> VQ -> decoder teacher-forced
logits -> masked cross-entropy on the next sample + vq_loss -> backward -> (data-parallel: one flat gradient all-reduce) ->
gradient clipping -> Adam.  Used by bench.py and the DP tests; the reference's script itself cannot be imported (docopt,
librosa, nnmnkwii are absent), so the three lines of its loss are restated here with their line numbers."""
from __future__ import annotations

import torch
from torch.nn import functional as F

from . import parallel


def make_optimizer(model, lr=4e-4, capturable=False):
    """hps/vqwae.json:50-55: Adam, lr 4e-4, betas (0.9, 0.999), eps 1e-8, no weight decay."""
    return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, capturable=capturable)


def train_step(model, opt, idx, mfcc, g, clip=100.0, world=1, timers=None):
    """idx (B,T) int64 mu-law classes; mfcc (B,39,frames); g (B,1) speaker ids.  Returns the loss tensor (no sync)."""
    x = F.one_hot(idx, 256).float().transpose(1, 2)                        # the collate's one-hot input (vqwae_train.py:509-520)
    opt.zero_grad(set_to_none=True)
    y_hat, vq_loss, _ = model(x, mfcc, g)                                  # vqvae_model.py:64-71
    # vqwae_train.py:760-766: y_hat[:, :, :-1] predicts y[:, 1:]; full-length synthetic windows -> the mask is all ones
    loss = F.cross_entropy(y_hat[:, :, :-1], idx[:, 1:]) + vq_loss
    loss.backward()
    if world > 1:
        if timers is not None:
            timers[0].record()
        parallel.allreduce_gradients(model)                                # replaces replicate/gather (vqwae_train.py:698-706)
        if timers is not None:
            timers[1].record()
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)               # hps/vqwae.json:63, vqwae_train.py:779-780
    opt.step()
    return loss.detach()


class GraphedTrainStep:
    """The whole step (forward kernels, backward GEMMs + kernels, all-reduce, clip, Adam) captured once into a CUDA graph and
    replayed: a step is ~3000 launches of mostly small kernels, so eager launching is CPU-bound (33 ms of host time for 16 ms
    of GPU work at 8 x 7680 samples).  Shapes are static; inputs are copied into the captured buffers."""

    def __init__(self, model, opt, idx, mfcc, g, clip=100.0, world=1, warmup=3):
        self.idx, self.mfcc, self.g = idx.clone(), mfcc.clone(), g.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                  # allocator, workspace cache, cuDNN / cuBLAS handles, NCCL warm-up
                train_step(model, opt, self.idx, self.mfcc, self.g, clip, world)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (the NCCL watchdog, a profiler) may touch the legacy stream meanwhile
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = train_step(model, opt, self.idx, self.mfcc, self.g, clip, world)

    def __call__(self, idx, mfcc, g):
        self.idx.copy_(idx, non_blocking=True)
        self.mfcc.copy_(mfcc, non_blocking=True)
        self.g.copy_(g, non_blocking=True)
        self.graph.replay()
        return self.loss
