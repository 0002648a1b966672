"""Graph-replayed VQ-WAE training step at BASELINE configs[2] per GPU (8 x 7680): fused / unfused loss, fp32 / TF32 encoder convolutions
(GPU box only)."""
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
for fused, tf32, cl in ((True, False, False), (True, False, True), (True, True, True), (True, True, False)):
    tm = bench.build_vqvae(dev).train(); tm.wavenet.precision = "bf16"; tm.wavenet.train_impl = "kernels"
    tm.encoder.train_tf32 = tf32
    tm.encoder.train_channels_last = cl
    opt = TS.FlatAdam(tm)
    orig = TS.train_step
    TS.train_step = lambda *a, **k: orig(*a, fused_loss=fused, **k)
    gs = TS.GraphedTrainStep(tm, opt, ti, tmf, tg)
    TS.train_step = orig
    for _ in range(3): gs(ti, tmf, tg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): loss = gs(ti, tmf, tg)
    e1.record(); torch.cuda.synchronize()
    print(f"fused_loss={fused} encoder_tf32={tf32} channels_last={cl}: graphed train step {e0.elapsed_time(e1) / 20:.3f} ms, loss {float(loss):.4f}")
    del gs, tm, opt
