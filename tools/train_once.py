"""A few EAGER VQ-WAE training steps at 8 x 7680 for ncu / compute-sanitizer (GPU box only):
ncu --set full -k regex:bwd_gemm --launch-skip 86 -c 4 python tools/train_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from wavenet_autoencoders_b200 import train_step as TS
dev = torch.device("cuda:0")
rs = np.random.RandomState(7)
Bt, Tt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 7680)
ti = torch.tensor(rs.randint(0, 256, size=(Bt, Tt)), dtype=torch.long, device=dev)
tmf = torch.tensor(rs.normal(size=(Bt, 39, Tt // 160)), dtype=torch.float32, device=dev)
tg = torch.tensor(rs.randint(0, 153, size=(Bt, 1)), dtype=torch.long, device=dev)
tm = bench.build_vqvae(dev).train(); tm.wavenet.precision = "bf16"; tm.wavenet.train_impl = "kernels"
opt = TS.FlatAdam(tm)
for _ in range(int(os.environ.get("STEPS", "3"))):
    loss = TS.train_step(tm, opt, ti, tmf, tg)
torch.cuda.synchronize()
print("loss", float(loss))
