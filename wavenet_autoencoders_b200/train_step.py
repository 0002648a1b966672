"""One VQ-WAE training step as vqwae_train.py:745-790 runs it (BASELINE configs[2]): encoder -> VQ -> decoder teacher-forced
logits -> masked cross-entropy on the next sample + vq_loss -> backward -> (data-parallel: one flat gradient all-reduce) ->
gradient clipping -> Adam.  Used by bench.py and the DP tests; the reference's script itself cannot be imported (docopt,
librosa, nnmnkwii are absent), so the three lines of its loss are restated here with their line numbers."""
from __future__ import annotations

import torch
from torch.nn import functional as F

from . import packing, parallel


def make_optimizer(model, lr=4e-4, capturable=False):
    """hps/vqwae.json:50-55: Adam, lr 4e-4, betas (0.9, 0.999), eps 1e-8, no weight decay."""
    return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, capturable=capturable)


class FlatAdam:
    """Adam + global-norm gradient clipping on ONE flat fp32 buffer (wae_sumsq / wae_adam_step).

    Every parameter becomes a view into ``flat_p`` and every ``.grad`` a view into ``flat_g`` (autograd accumulates into them
    in place), so the data-parallel exchange is a single all-reduce of ``flat_g`` with no gather/scatter copies, and the
    clip + update of all tensors is two kernel launches.  Same update rule as ``torch.optim.Adam`` (no amsgrad, no weight decay)
    after ``clip_grad_norm_`` -- tests/test_gpu_parity.py.  Parameters that never receive a gradient (the last layer's
    ``conv1x1_out``) see zeros and do not move, like parameters torch's Adam skips."""

    def __init__(self, model, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, clip=100.0, ema_decay=None):
        from . import _lib
        self._lib = _lib
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        ALIGN = 64                                   # floats: every parameter view starts 256-byte aligned (kernels use 16-byte loads)
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += -(-p.numel() // ALIGN) * ALIGN
        self.offsets = offs
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m, self.v = torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p)
        for p, off in zip(self.params, offs):
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
        self.betas, self.eps, self.clip = (float(betas[0]), float(betas[1])), float(eps), float(clip or 0.0)
        # torch.optim surface the reference's loop touches: it writes the scheduled rate into param_group['lr'] every step
        # (vqwae_train.py:730-735) and saves / restores optimizer.state_dict() (save_checkpoint / load_checkpoint)
        self.param_groups = [{"params": self.params, "lr": float(lr), "betas": self.betas, "eps": self.eps}]
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)     # what the kernel reads (graph replay safe)
        self._lr_on_dev = float(lr)
        self.step_a = torch.zeros(1, dtype=torch.float32, device=dev)       # step count lives on the device (CUDA-graph safe)
        self.step_b = torch.zeros(1, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        # the reference's ExponentialMovingAverage shadow of every parameter (vqwae_train.py:337-350), updated in the same pass
        self.ema_decay = None if ema_decay is None else float(ema_decay)
        self.ema = self.flat_p.clone() if ema_decay is not None else None

    def ema_state(self):
        """{parameter: shadow tensor} views into the flat shadow buffer (what clone_as_averaged_model copies, :353-360)."""
        return {p: self.ema[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)}

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        self.param_groups[0]["lr"] = float(value)

    def sync_lr(self):
        """Copy param_groups[0]['lr'] to the device scalar the update kernel reads (a no-op while it is unchanged).  ``step()``
        calls it; with ``GraphedTrainStep`` it runs before every replay, so a scheduled learning rate takes effect."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_dev:
            self.lr_dev.fill_(lr)
            self._lr_on_dev = lr

    def zero_grad(self, set_to_none=False):
        """Zeroes the flat gradient buffer; ``.grad`` stays a view into it whatever ``set_to_none`` says."""
        self.flat_g.zero_()

    def _reattach_grads(self):
        """``model.zero_grad()`` (set_to_none) or a fresh ``.grad`` tensor detaches a parameter from ``flat_g``; the update reads
        ``flat_g``, so bring stray gradients back (copy) and restore the views."""
        for p, off in zip(self.params, self.offsets):
            view = self.flat_g[off:off + p.numel()]
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad.reshape(-1))
            else:
                continue
            p.grad = view.view_as(p)

    def state_dict(self):
        """Adam moments, step count, EMA shadow and learning rate (what a checkpoint needs to resume: vqwae_train.py save_checkpoint)."""
        return {"m": self.m.clone(), "v": self.v.clone(), "step": self.step_a.clone(), "ema": None if self.ema is None else self.ema.clone(),
                "lr": float(self.param_groups[0]["lr"]), "betas": self.betas, "eps": self.eps, "clip": self.clip,
                "ema_decay": self.ema_decay, "numel": int(self.flat_p.numel())}

    def load_state_dict(self, sd):
        if int(sd["numel"]) != int(self.flat_p.numel()):
            raise ValueError(f"FlatAdam.load_state_dict: {sd['numel']} elements in the checkpoint, {self.flat_p.numel()} here")
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.step_a.copy_(sd["step"]); self.step_b.copy_(sd["step"])
        if self.ema is not None and sd.get("ema") is not None:
            self.ema.copy_(sd["ema"])
        self.param_groups[0]["lr"] = float(sd["lr"])
        self.sync_lr()

    def enable_overlap(self, model=None, bounds=None, group=None):
        """Bucketed all-reduce overlapped with the backward (parallel.BucketedAllReduce).  Default buckets: [the WaveNet
        decoder's parameters | everything else] when ``model`` has a ``wavenet`` whose parameters come first -- the decoder's
        gradients are complete when its backward node finishes, before the upsampler / VQ / encoder backward runs."""
        from . import parallel
        if bounds is None:
            n_dec = 0
            dec = getattr(model, "wavenet", None)
            if dec is not None:
                ids = {id(p) for p in dec.parameters()}
                while n_dec < len(self.params) and id(self.params[n_dec]) in ids:
                    n_dec += 1
            bounds = [0, n_dec, len(self.params)] if 0 < n_dec < len(self.params) else [0, len(self.params)]
        self.bucketed = parallel.BucketedAllReduce(self.params, self.flat_g, self.offsets, bounds, group)
        return self.bucketed

    def allreduce(self, world):
        if getattr(self, "bucketed", None) is not None:
            self.bucketed.finish()
            return
        import torch.distributed as dist
        dist.all_reduce(self.flat_g)
        self.flat_g.div_(world)

    def step(self):
        L, ptr = self._lib.lib(), self._lib.ptr
        st = self._lib.stream_ptr(self.flat_p.device)
        n = self.flat_p.numel()
        if not torch.cuda.is_current_stream_capturing():
            self._reattach_grads()
            self.sync_lr()
        if self.clip > 0:
            self.sumsq.zero_()
            self._lib.check(L.wae_sumsq(ptr(self.flat_g), n, ptr(self.sumsq), st), "wae_sumsq")
        self._lib.check(L.wae_adam_step_dlr(ptr(self.flat_p), ptr(self.flat_g), ptr(self.m), ptr(self.v), n, ptr(self.lr_dev), self.betas[0],
                                            self.betas[1], self.eps, self.clip, ptr(self.sumsq), ptr(self.step_a), ptr(self.step_b),
                                            ptr(self.ema), self.ema_decay or 0.0, st),
                        "wae_adam_step_dlr")
        self.step_a.copy_(self.step_b)
        packing.bump_generation()        # parameters changed through raw pointers: packed-weight caches must re-pack


def train_step(model, opt, idx, mfcc, g, clip=100.0, world=1, timers=None, fused_loss=True, index_input=True):
    """idx (B,T) int64 mu-law classes; mfcc (B,39,frames); g (B,1) speaker ids.  Returns the loss tensor (no sync)."""
    opt.zero_grad(set_to_none=True)
    if world > 1 and getattr(opt, "bucketed", None) is not None:
        opt.bucketed.start_step()
    # vqwae_train.py:760-766: y_hat[:, :, :-1] predicts y[:, 1:]; full-length synthetic windows -> the mask is all ones
    wn = getattr(model, "wavenet", None)
    index_input = (index_input and fused_loss and hasattr(model, "forward_nll") and wn is not None and idx.is_cuda
                   and getattr(wn, "fused_training_ok", lambda _x: False)(idx))
    # the collate's one-hot input (vqwae_train.py:509-520) is one_hot(classes): the fused kernel path takes the classes themselves
    x = idx if index_input else F.one_hot(idx, 256).float().transpose(1, 2)
    if fused_loss and hasattr(model, "forward_nll"):
        nll, vq_loss, _ = model.forward_nll(x, mfcc, g, idx, 1)            # decoder loss + backward fused (training.StackNLLFunction)
        loss = nll + vq_loss
    else:
        y_hat, vq_loss, _ = model(x, mfcc, g)                              # vqvae_model.py:64-71
        loss = F.cross_entropy(y_hat[:, :, :-1], idx[:, 1:]) + vq_loss
    loss.backward()
    flat = isinstance(opt, FlatAdam)
    if world > 1:
        if timers is not None:
            timers[0].record()
        if flat:
            opt.allreduce(world)                                           # one all-reduce of the flat gradient buffer, no copies
        else:
            parallel.allreduce_gradients(model)                            # replaces replicate/gather (vqwae_train.py:698-706)
        if timers is not None:
            timers[1].record()
    if not flat:                                                           # FlatAdam clips inside its update (opt.clip)
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip)           # hps/vqwae.json:63, vqwae_train.py:779-780
    opt.step()
    return loss.detach()


class GraphedTrainStep:
    """The whole step (forward kernels, backward GEMMs + kernels, all-reduce, clip, Adam) captured once into a CUDA graph and
    replayed: a step is ~3000 launches of mostly small kernels, so eager launching is CPU-bound (33 ms of host time for 16 ms
    of GPU work at 8 x 7680 samples).  Shapes are static; inputs are copied into the captured buffers."""

    def __init__(self, model, opt, idx, mfcc, g, clip=100.0, world=1, warmup=3):
        self.idx, self.mfcc, self.g = idx.clone(), mfcc.clone(), g.clone()
        self.opt = opt
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
        if hasattr(self.opt, "sync_lr"):
            self.opt.sync_lr()           # a scheduled learning rate (opt.param_groups[0]['lr']) reaches the captured update kernel
        self.graph.replay()
        packing.bump_generation()        # the replayed optimiser kernels rewrote every parameter in place
        return self.loss
