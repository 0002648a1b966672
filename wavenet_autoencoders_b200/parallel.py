"""Multi-GPU plumbing for the hot path: one process per GPU over torch.distributed.

* Teacher-forced inference, VQ and AR synthesis shard by UTTERANCE with no data-path collective
  (``shard_utterances``); random streams are indexed by global utterance id so results do not depend on the
  number of GPUs (``utterance_uniforms``).
* Training is data-parallel with ONE collective: an all-reduce (sum, then /world) of the flat gradient
  (``allreduce_gradients``), replacing the reference's replicate/scatter/gather (vqwae_train.py:698-706).
  Parameters that received no gradient (the last layer's conv1x1_out) contribute zeros.
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
