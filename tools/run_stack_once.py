"""Runs the bf16 teacher-forced stack at the bench shape a few times (target of ncu captures). GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
m = bench.build_vqvae("cuda")
m.wavenet.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
idx, mfcc, g = bench.synth_batch(16, 1000)
idx, mfcc, g = idx.cuda(), mfcc.cuda(), g.cuda()
x = torch.nn.functional.one_hot(idx, 256).float().transpose(1, 2).contiguous()
with torch.no_grad():
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
        y = m(x, mfcc, g)[0]
    torch.cuda.synchronize()
    torch.cuda.profiler.start()          # ncu --profile-from-start off: exactly one steady-state forward
    y = m(x, mfcc, g)[0]
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", float(y.abs().mean()))
