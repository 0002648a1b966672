"""Where the end-to-end step's time goes beyond its kernels (GPU box only): graph replay alone vs H2D + replay + D2H per step,
with the class indices travelling as int64 (the reference's dtype) or as uint8 (mu-law classes are 8-bit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wavenet_autoencoders_b200.graphed import GraphedForward

m = bench.build_vqvae("cuda"); m.wavenet.precision = "bf16"
idx, mfcc, g = bench.synth_batch(16, 1000)
idx_p, mfcc_p, g_p = idx.pin_memory(), mfcc.pin_memory(), g.pin_memory()
gf = GraphedForward(m, idx.cuda(), mfcc.cuda(), g.cuda(), with_logits=False)


def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"replay only, back to back:            {timed(lambda: gf.graph.replay()):.4f} ms")
print(f"replay + loss read-back per step:     {timed(lambda: (gf.graph.replay(), float(gf.nll.item()))):.4f} ms")
print(f"H2D int64 + replay + read-back:       {timed(lambda: float(gf(idx_p, mfcc_p, g_p)[3].item())):.4f} ms")
pend = []


def pipe():
    t = gf.submit(idx_p, mfcc_p, g_p)
    if pend:
        gf.result(pend.pop())
    pend.append(t)


print(f"submit / result one step behind:      {timed(pipe):.4f} ms")
gf.result(pend.pop())


def h2d_replay():
    gf.idx.copy_(idx_p, non_blocking=True); gf.mfcc.copy_(mfcc_p, non_blocking=True); gf.g.copy_(g_p, non_blocking=True)
    gf.graph.replay()


print(f"H2D + replay, no read-back at all:    {timed(h2d_replay):.4f} ms")
host = torch.empty((), dtype=gf.nll.dtype).pin_memory()


def h2d_replay_d2h():
    h2d_replay(); host.copy_(gf.nll, non_blocking=True)


print(f"H2D + replay + async D2H, no wait:    {timed(h2d_replay_d2h):.4f} ms")
print(f"replay only again:                    {timed(lambda: gf.graph.replay()):.4f} ms")
