"""Phase clocks of the fused encoder + VQ kernel (CTA 0): build with WAE_NVCC_DEFS=WAE_EV_PROF, GPU box only.

    WAE_NVCC_DEFS=WAE_EV_PROF python -m wavenet_autoencoders_b200.build --force && python tools/enc_profile.py [B] [frames]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from wavenet_autoencoders_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
m = bench.build_vqvae("cuda")
x = torch.randn(B, 39, frames, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m._encode_quantize(x)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    _lib.check(_lib.lib().wae_encoder_vq_profile(buf), "wae_encoder_vq_profile")
t = list(buf)
t0 = t[41]
print(f"setup (zero + input) {t[0] - t0:8d} cycles   first barrier {t[1] - t[0]:8d}")
prev = t[1]
nl = len(m.encoder.net)
for l in range(nl):
    print(f"layer {l:2d}: compute+store {t[2 + 2 * l] - prev:8d}   barrier {t[3 + 2 * l] - t[2 + 2 * l]:8d}")
    prev = t[3 + 2 * l]
print(f"tail (Linear + VQ)   {t[40] - prev:8d}")
print(f"total                {t[40] - t0:8d} cycles")
for name, o in (("layer 7 (k=1)", 42), ("layer 1 (k=3, 100 frames)", 52)):
    lbl = ["prologue -> chunk 0 ready", "chunk 0 -> 1 ready", "chunk 1 -> 2 ready", "chunk 2 ready -> loop end", "loop end -> partial sums reduced"]
    print(name + ": " + "; ".join(f"{lbl[i]} {t[o + i + 1] - t[o + i]}" for i in range(5)))
