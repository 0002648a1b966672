"""One steady-state e2e step of the bench workload (class indices -> VQVAE.forward_nll, bf16 kernels, NLL from the head kernel)
between cudaProfilerStart/Stop -- the target of the ncu launch list committed under profiles/.  GPU box only.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/step_once.py [logits]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

m = bench.build_vqvae("cuda")
m.wavenet.precision = "bf16"
idx, mfcc, g = bench.synth_batch(16, 1000)
idx, mfcc, g = idx.cuda(), mfcc.cuda(), g.cuda()
with_logits = len(sys.argv) > 1 and sys.argv[1] == "logits"


def step():
    with torch.no_grad():
        if with_logits:
            return m(idx, mfcc, g)[0].float().mean()
        return m.forward_nll(idx, mfcc, g, idx, 1)[0]


for _ in range(3):
    v = step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
v = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(v))
