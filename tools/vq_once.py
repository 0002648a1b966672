"""One VQ search at N = 2^20 (for ncu: -k regex:vq_search_resident).  Usage: python tools/vq_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavenet_autoencoders_b200.vector_quantization import VectorQuantize  # noqa: E402

mod = VectorQuantize(256, 64).cuda()
x = torch.randn(64, 64, 1 << 14, device="cuda") * 0.05
with torch.no_grad():
    for _ in range(3):
        mod(x)
torch.cuda.synchronize()
print("ok")
