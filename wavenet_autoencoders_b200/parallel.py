"""Multi-GPU plumbing for the hot path: one process per GPU over torch.distributed.

* Teacher-forced inference, VQ and AR synthesis shard by UTTERANCE with no data-path collective
  (``shard_utterances``); random streams are indexed by global utterance id so results do not depend on the
  number of GPUs (``utterance_uniforms``).
* Training is data-parallel with ONE exchange step: an all-reduce (sum, then /world) of the flat gradient
  (``allreduce_gradients``), replacing the reference's replicate/scatter/gather (vqwae_train.py:698-706).
  Parameters that received no gradient (the last layer's conv1x1_out) contribute zeros.
  ``BucketedAllReduce`` cuts the flat buffer into buckets in the order the backward pass completes them and starts
  each bucket's all-reduce as soon as its last gradient has been accumulated, so that the exchange of the decoder's
  gradients (80 % of the bytes) runs under the rest of the backward (upsampler, VQ, encoder).
"""
from __future__ import annotations

import torch


def shard_utterances(n_utts: int, world: int, rank: int) -> range:
    """Balanced contiguous shard of utterance ids for this rank."""
    lo = (n_utts * rank) // world
    hi = (n_utts * (rank + 1)) // world
    return range(lo, hi)


def utterance_uniforms(utt_ids, T: int, n: int, seed: int, device) -> torch.Tensor:
    """(T, len(utt_ids), n) uniforms in [0,1); column u depends only on (seed, utt_ids[u]) -- not on the sharding."""
    cols = []
    for uid in utt_ids:
        g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + int(uid))
        cols.append(torch.rand(T, n, generator=g))
    out = torch.stack(cols, dim=1) if cols else torch.zeros(T, 0, n)
    return out.to(device)


def allreduce_gradients(module: torch.nn.Module, group=None) -> int:
    """Average gradients over the process group through one flat buffer. Returns the number of elements reduced."""
    import torch.distributed as dist
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        return 0
    dev, dt = params[0].device, params[0].dtype
    flat = torch.zeros(sum(p.numel() for p in params), device=dev, dtype=dt)
    off = 0
    for p in params:
        if p.grad is not None:
            flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
        off += p.numel()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()
    return flat.numel()


class BucketedAllReduce:
    """Overlapped gradient exchange on a FLAT gradient buffer whose slices are the parameters' ``.grad`` views.

    ``bounds`` are the bucket boundaries as indices into ``params`` (bucket i = params[bounds[i]:bounds[i+1]], contiguous in
    the flat buffer because the views are laid out in parameter order).  A post-accumulate-grad hook on every parameter
    counts its bucket down; the bucket's all-reduce is issued asynchronously (its own communication stream; NCCL / gloo)
    the moment the count reaches zero, i.e. while autograd is still running the remaining backward.  Which parameters
    receive a gradient at all is learnt from the first step (the last layer's dead ``conv1x1_out`` never does): that step
    reduces every bucket after the backward.  ``finish()`` issues what is still pending, waits for everything and divides
    by the world size."""

    def __init__(self, params, flat_g, offsets, bounds, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.params, self.flat_g = list(params), flat_g
        self.bounds = list(bounds)
        nb = len(self.bounds) - 1
        ends = [offsets[self.bounds[i + 1]] if self.bounds[i + 1] < len(self.params) else flat_g.numel() for i in range(nb)]
        starts = [offsets[self.bounds[i]] for i in range(nb)]
        self.slices = [flat_g[a:b] for a, b in zip(starts, ends)]
        self.bucket_of = {}
        for i in range(nb):
            for j in range(self.bounds[i], self.bounds[i + 1]):
                self.bucket_of[id(self.params[j])] = i
        self.expected = None                     # per bucket: number of parameters that fire, learnt in the first step
        self.fired = [set() for _ in range(nb)]
        self.left = [0] * nb
        self.works = [None] * nb
        self.overlapped = 0                      # buckets whose all-reduce started inside the backward (last step)
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def start_step(self):
        nb = len(self.slices)
        self.works = [None] * nb
        self.overlapped = 0
        if self.expected is not None:
            self.left = list(self.expected)
        for f in self.fired:
            f.clear()

    def _hook(self, p):
        i = self.bucket_of[id(p)]
        if id(p) in self.fired[i]:
            return
        self.fired[i].add(id(p))
        if self.expected is None:
            return
        self.left[i] -= 1
        if self.left[i] == 0 and self.works[i] is None:
            if self.flat_g.is_cuda:
                # the bucket's gradients were accumulated on several side streams (packing.Lanes: layer l's weight-norm backward
                # runs on lane l); the collective is ordered after the CURRENT stream only -- make it see all of them
                from . import packing
                packing.join_lane_streams_into_current(self.flat_g.device)
            self.works[i] = self.dist.all_reduce(self.slices[i], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.overlapped += 1

    def finish(self):
        world = self.dist.get_world_size(self.group)
        if self.expected is None:                # calibration step: nothing was issued from the hooks
            self.expected = [len(f) for f in self.fired]
        for i, sl in enumerate(self.slices):
            if self.works[i] is None:
                self.works[i] = self.dist.all_reduce(sl, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
        for w in self.works:
            w.wait()
        self.flat_g.div_(world)

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []
