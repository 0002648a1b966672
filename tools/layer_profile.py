"""Role-level cycle breakdown of the residual-layer kernels (needs a WAE_LAYER_PROF=1 build).  GPU box only.

    WAE_LAYER_PROF=1 python -m wavenet_autoencoders_b200.build --force; python tools/layer_profile.py [-1 | -2 | 0 | 1 | 2 | 4]

The argument is wae_set_layer_cluster's: -1 version 2 (default kernel), -2 version 2 on CTA pairs (per 256-sample super-tile).
For the version-2 kernels E1 is also split: a = waiting for the previous tile's TMA stores + barrier, b = TMEM loads + gate math +
st.shared, rest = fences, barrier, TMA store issue, barrier arrival."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_autoencoders_b200 import _lib
if os.environ.get("WAE_LIB_VARIANT"):      # experiment build (tools/build_variant.py); the product loads libwae_b200.so
    _lib.LIB_PATH = _lib.PKG / "variants" / f"libwae_{os.environ['WAE_LIB_VARIANT']}.so"
import bench
L = _lib.lib()
m = bench.build_vqvae("cuda"); m.wavenet.precision = "bf16"
idx, mfcc, g = bench.synth_batch(16, 1000)
idx, mfcc, g = idx.cuda(), mfcc.cuda(), g.cuda()
x = torch.nn.functional.one_hot(idx, 256).float().transpose(1, 2).contiguous()
_lib.check(L.wae_set_layer_cluster(int(sys.argv[1]) if len(sys.argv) > 1 else -1), "mode")
with torch.no_grad():
    for _ in range(2):
        m(x, mfcc, g)
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    L.wae_layer_set_profile_buffer(buf.data_ptr())
    m(x, mfcc, g)          # the buffer holds the LAST layer launched that has GEMM2?  (every launch overwrites) -> last layer
    torch.cuda.synchronize()
    L.wae_layer_set_profile_buffer(None)
p = buf.view(148, 16).double().cpu()
names = ["prod wait-empty", "prod total", "mma wait-full", "mma wait-epi1", "mma wait-epi2", "mma issue", "mma total", "tiles",
         "epi wait-acc1", "epi E1", "epi wait-acc2", "epi E2", "epi prefetch", "epi total", "epi E1 a (store wait)", "epi E1 b (gate)"]
tiles = p[:, 7].clamp(min=1)
print("counters of layer 5 (dilation 32), cycles")
for i, n in enumerate(names):
    col = p[:, i][p[:, 13] > 0] if i < 8 and (p[:, 6] > 0).sum() < 148 else p[:, i]   # pair kernels: MMA counters exist in leaders only
    colt = col if len(col) == 148 else p[:, i][p[:, 6] > 0]
    tl = tiles if len(colt) == 148 else tiles[p[:, 6] > 0]
    print(f"{n:22s} mean {colt.mean().item():10.0f}  max {colt.max().item():10.0f}  per tile {(colt / tl).mean().item():9.0f}")
