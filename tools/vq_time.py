"""Time the VQ nearest-codeword search at several sizes with both kernels of wae_vq_search (0 = automatic: the codebook-resident
persistent kernel for many vectors; 1 = the chunked kernel).  Usage (on a B200): python tools/vq_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavenet_autoencoders_b200 import _lib  # noqa: E402
from wavenet_autoencoders_b200.vector_quantization import SlicedVectorQuantize, VectorQuantize  # noqa: E402


def main():
    L = _lib.lib()
    dev = torch.device("cuda:0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, mod in (("VectorQuantize(256, 64)", VectorQuantize(256, 64)), ("SlicedVectorQuantize(256, 64)", SlicedVectorQuantize(256, 64))):
        mod = mod.to(dev)
        for N in (1 << 14, 1 << 17, 1 << 20):
            x = torch.randn(64, 64, N // 64, device=dev) * 0.05
            res = {}
            for variant in (1, 0):
                _lib.check(L.wae_vq_set_variant(variant), "wae_vq_set_variant")
                with torch.no_grad():
                    for _ in range(3):
                        mod(x)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(10):
                        mod(x)
                    e1.record()
                    torch.cuda.synchronize()
                res[variant] = e0.elapsed_time(e1) / 10
            print(f"{name:32s} N={N:8d}  chunked {res[1] * 1e3:9.1f} us  resident {res[0] * 1e3:9.1f} us  "
                  f"({N / res[0] / 1e3:8.1f} M vectors/s, {N * 2 * 256 * 64 / res[0] / 1e9:6.2f} TFLOP/s fp32)", flush=True)
    L.wae_vq_set_variant(0)


if __name__ == "__main__":
    main()
