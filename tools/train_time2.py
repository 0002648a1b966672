"""Graph-replayed VQ-WAE training step at BASELINE configs[2] per GPU (8 x 7680) under environment / attribute switches given on
the command line as name=value pairs, e.g.  `python tools/train_time2.py up=autograd up=kernels`  (GPU box only)."""
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
for spec in sys.argv[1:] or ["default"]:
    tm = bench.build_vqvae(dev).train(); tm.wavenet.precision = "bf16"; tm.wavenet.train_impl = "kernels"
    for kv in spec.split(","):
        if "=" not in kv:
            continue
        k, v = kv.split("=")
        if k == "up":                                  # upsampler stages: "kernels" (UpsampleStageFunction) or "autograd" (torch / cuDNN)
            tm.wavenet._upsample = (lambda c, un=tm.wavenet.upsample_net, v=v: (setattr(un.upsample, "train_impl", v), un(c))[1])
        elif k == "enc":
            tm.encoder.train_impl = v
        elif k == "tf32":
            tm.encoder.train_tf32 = v == "1"
        else:
            os.environ[k] = v
    opt = TS.FlatAdam(tm)
    gs = TS.GraphedTrainStep(tm, opt, ti, tmf, tg)
    for _ in range(3): gs(ti, tmf, tg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(30): loss = gs(ti, tmf, tg)
    e1.record(); torch.cuda.synchronize()
    print(f"{spec}: graphed train step {e0.elapsed_time(e1) / 30:.3f} ms, loss {float(loss):.4f}", flush=True)
    del gs, tm, opt
