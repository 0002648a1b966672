"""Host-side cost of one eager e2e step (class indices -> VQVAE.forward_nll): cProfile over 50 steps. GPU box only."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
m = bench.build_vqvae("cuda"); m.wavenet.precision = "bf16"
idx, mfcc, g = bench.synth_batch(16, 1000)
idx, mfcc, g = idx.cuda(), mfcc.cuda(), g.cuda()
def step():
    with torch.no_grad():
        return m.forward_nll(idx, mfcc, g, idx, 1)[0]
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): v = step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time per step {1e3 * (t1 - t0) / 50:.3f} ms; with final sync {1e3 * (t2 - t0) / 50:.3f} ms per step")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): v = step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
