"""Two-stream decoder backward (wae_stack_backward_bf16_2s: weight-gradient GEMMs on a side stream under the upsampler / VQ /
encoder backward) against the one-stream call at BASELINE configs[2] per GPU (8 x 7680): gradients of one eager step compared
tensor by tensor (the split-K wgrads add with fp32 red.add, so "equal" means equal up to summation order), then the
graph-replayed step timed both ways (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from wavenet_autoencoders_b200 import train_step as TS

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
rs = np.random.RandomState(7); Bt, Tt = 8, 7680
ti = torch.tensor(rs.randint(0, 256, size=(Bt, Tt)), dtype=torch.long, device=dev)
tmf = torch.tensor(rs.normal(size=(Bt, 39, Tt // 160)), dtype=torch.float32, device=dev)
tg = torch.tensor(rs.randint(0, 153, size=(Bt, 1)), dtype=torch.long, device=dev)


def grads(streams):
    os.environ["WAE_BWD_STREAMS"] = streams
    tm = bench.build_vqvae(dev).train(); tm.wavenet.precision = "bf16"; tm.wavenet.train_impl = "kernels"
    opt = TS.FlatAdam(tm)
    opt.step = lambda: None                                   # keep the gradients of the step
    out = []
    for _ in range(3):
        loss = TS.train_step(tm, opt, ti, tmf, tg)
        torch.cuda.synchronize()
        out.append((float(loss), opt.flat_g.clone()))
    names = [n for n, p in tm.named_parameters() if p.requires_grad]
    return out, names, opt.offsets, [p.numel() for p in opt.params]


g1, names, offs, sizes = grads("1")
g2, _, _, _ = grads("2")
for it in range(3):
    (l1, a), (l2, b) = g1[it], g2[it]
    worst, wn_ = 0.0, ""
    for n, o, k in zip(names, offs, sizes):
        x, y = a[o:o + k].double(), b[o:o + k].double()
        d = float((x - y).norm() / max(float(x.norm()), 1e-30))
        if d > worst:
            worst, wn_ = d, n
    print(f"eager step {it}: loss {l1:.6f} / {l2:.6f}; worst per-tensor rel L2 between 1-stream and 2-stream gradients {worst:.3e} ({wn_})")
    assert abs(l1 - l2) < 1e-6 and worst < 1e-4, (l1, l2, worst, wn_)
# repeatability of the one-stream call itself (the red.add order), for scale
print("1-stream run-to-run rel L2:", float((g1[1][1] - g1[2][1]).double().norm() / g1[1][1].double().norm()))

for streams in ("1", "2", "1", "2"):
    os.environ["WAE_BWD_STREAMS"] = streams
    tm = bench.build_vqvae(dev).train(); tm.wavenet.precision = "bf16"; tm.wavenet.train_impl = "kernels"
    opt = TS.FlatAdam(tm)
    gs = TS.GraphedTrainStep(tm, opt, ti, tmf, tg)
    for _ in range(3): gs(ti, tmf, tg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): loss = gs(ti, tmf, tg)
    e1.record(); torch.cuda.synchronize()
    print(f"WAE_BWD_STREAMS={streams}: graphed train step {e0.elapsed_time(e1) / 20:.3f} ms, loss after 23 steps {float(loss):.5f}")
    del gs, tm, opt
