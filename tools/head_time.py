"""Head kernel variants (1-CTA vs CTA pairs) at the bench shape: time per launch and identical logits (GPU box only)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wavenet_autoencoders_b200 import _lib
L = _lib.lib()
m = bench.build_vqvae("cuda"); m.wavenet.precision = "bf16"
idx, mfcc, g = bench.synth_batch(16, 1000)
idx, mfcc, g = idx.cuda(), mfcc.cuda(), g.cuda()
ref = None
for pair in (0, 1):
    L.wae_set_head_pair(pair)
    with torch.no_grad():
        for _ in range(3):
            y = m(idx, mfcc, g)[0]
        torch.cuda.synchronize()
        L.wae_profile_enable(1)
        for _ in range(10):
            y = m(idx, mfcc, g)[0]
        torch.cuda.synchronize()
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int32 * 4)()
        L.wae_profile_read(ms, n, 4); L.wae_profile_enable(0)
        nll = float(m.forward_nll(idx, mfcc, g, idx, 1)[0])
    if ref is None:
        ref = y.clone()
    print(f"head_pair={pair}: head {ms[2] / max(n[2], 1) * 1e3:.1f} us  layer {ms[1] / max(n[1], 1) * 1e3:.1f} us  max|dy| {float((y - ref).abs().max()):.3e}  nll {nll:.6f}")
