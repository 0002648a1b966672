"""world_size-2 gloo tests (CPU) of the N>1 host logic: utterance sharding and the flat gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wavenet_autoencoders_b200 import parallel
from wavenet_autoencoders_b200 import testing as T


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
        torch.manual_seed(0)
        m = WaveNet(**T.CONFIGS["tiny"]).train()
        m.load_state_dict(T.synth_state_dict(m, 1))
        m.train_impl = "autograd"        # CPU / gloo: the torch-op composite, opted into explicitly (host logic under test)
        # each rank: its own utterance shard of a global batch of 4
        ids = parallel.shard_utterances(4, world, rank)
        x, _, c, g = T.synth_inputs(T.CONFIGS["tiny"], 4, 64, 0)
        sl = slice(ids.start, ids.stop)
        y = m(x[sl], c[sl], g[sl])
        (y.square().sum() / (4 * y[0].numel())).backward()
        n = parallel.allreduce_gradients(m)
        grads = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        u = parallel.utterance_uniforms(ids, 8, 1, seed=7, device="cpu")
        # numpy (pickled by value): tensors travel as shared-memory handles that die with an exiting worker
        q.put((rank, n, grads.numpy(), list(ids), u.numpy()))
    finally:
        dist.destroy_process_group()


def test_dp_allreduce_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    res = [(r, n, torch.from_numpy(gr), ids, torch.from_numpy(u)) for r, n, gr, ids, u in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: full batch, same loss normalisation; DP averages per-rank grads -> x world
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(0)
    m = WaveNet(**T.CONFIGS["tiny"]).train()
    m.load_state_dict(T.synth_state_dict(m, 1))
    m.train_impl = "autograd"
    x, _, c, g = T.synth_inputs(T.CONFIGS["tiny"], 4, 64, 0)
    y = m(x, c, g)
    (y.square().sum() / (4 * y[0].numel())).backward()
    full = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in m.parameters()])
    assert res[0][1] == full.numel() == sum(p.numel() for p in m.parameters())
    assert torch.allclose(res[0][2], res[1][2])                       # replicas agree after the all-reduce
    assert torch.allclose(res[0][2] * world, full, rtol=1e-4, atol=1e-6)
    assert res[0][3] == [0, 1] and res[1][3] == [2, 3]
    # random streams depend on the global utterance id only
    all_u = parallel.utterance_uniforms(range(4), 8, 1, seed=7, device="cpu")
    assert torch.equal(torch.cat([res[0][4], res[1][4]], dim=1), all_u)


def test_shard_utterances_is_a_partition():
    for n, w in [(256, 8), (7, 4), (3, 8), (0, 2)]:
        got = [i for r in range(w) for i in parallel.shard_utterances(n, w, r)]
        assert got == list(range(n))


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
        dead = torch.nn.Parameter(torch.zeros(7))                      # never receives a gradient (like the last layer's conv1x1_out)
        params = list(net[4].parameters()) + [dead] + list(net[2].parameters()) + list(net[0].parameters())   # backward order
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += -(-p.numel() // 4) * 4
        flat = torch.zeros(n)
        for p, o in zip(params, offs):
            p.grad = flat[o:o + p.numel()].view_as(p)
        br = parallel.BucketedAllReduce(params, flat, offs, [0, 3, 5, len(params)])
        outs = []
        for step in range(3):
            flat.zero_()
            br.start_step()
            x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100 * step + rank))
            net(x).square().mean().backward()
            br.finish()
            outs.append((flat.clone().numpy(), br.overlapped))
        q.put((rank, outs))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_overlaps_and_matches_flat_world2():
    """parallel.BucketedAllReduce: the first step calibrates (no overlap), later steps start a bucket's all-reduce from the
    hook of its last gradient (all but the last bucket start INSIDE the backward); the result equals the average of the two
    ranks' gradients, the dead parameter's slot stays zero."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import numpy as np
    for step in range(3):
        np.testing.assert_allclose(res[0][step][0], res[1][step][0], rtol=1e-6, atol=1e-7)       # replicas agree
    assert res[0][0][1] == 0 and res[0][1][1] == 3 and res[0][2][1] == 3                            # calibration, then every bucket from its hook
    # reference: average of the per-rank gradients computed locally
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
    want = None
    for rank in range(world):
        net.zero_grad()
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100 * 2 + rank))
        net(x).square().mean().backward()
        g = torch.cat([p.grad.reshape(-1) for p in net[4].parameters()])
        want = g if want is None else want + g
    np.testing.assert_allclose(res[0][2][0][:want.numel()], (want / world).numpy(), rtol=1e-5, atol=1e-7)
