"""Times the bf16 stack with the residual-layer kernel in clusters of 1/2/4 CTAs (TMA weight multicast). GPU box only."""
import sys, os, ctypes
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
ref = None
for cs in ([int(a) for a in sys.argv[1:]] or (-1, -4, -3, -2)):   # -3 = version-3 kernel (TMEM ping-pong, default); -1 = version-2 kernel; -2 = version 2 on CTA pairs; 0 = first CTA-pair kernel; 1 = first 1-CTA kernel
    _lib.check(L.wae_set_layer_cluster(cs), "set cluster")
    with torch.no_grad():
        for _ in range(3):
            y = m(x, mfcc, g)[0]
        torch.cuda.synchronize()
        L.wae_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = m(x, mfcc, g)[0]
        e1.record(); torch.cuda.synchronize()
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int32 * 4)()
        L.wae_profile_read(ms, n, 4); L.wae_profile_enable(0)
    if ref is None:
        ref = y.clone()
    err = float((y - ref).abs().max() / ref.abs().max())
    tot = e0.elapsed_time(e1) / 10
    lay = ms[1] / max(n[1], 1)
    print(f"cluster={cs}: step {tot:.3f} ms  ({16*16000/tot*1e3/1e6:.1f} M samples/s)  layer kernel {lay*1e3:.1f} us avg  "
          f"= {16*16000*490701/lay/1e9:.0f} TFLOP/s   head {ms[2]/max(n[2],1)*1e3:.1f} us  prep {ms[0]/max(n[0],1)*1e3:.1f} us   max|dy| vs cs=1: {err:.2e}")
