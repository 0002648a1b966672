"""Times AR synthesis variants at the vqwae shape. GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_autoencoders_b200 import testing as T
from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet

def run(B, Tn, prec, impl, cluster, U):
    cfg = T.CONFIGS["vqwae"]
    torch.manual_seed(0)
    m = WaveNet(**cfg).eval(); m.load_state_dict(T.synth_state_dict(m, 1)); m = m.cuda()
    m.precision, m.ar_impl, m.ar_cluster, m.ar_utts_per_cluster = prec, impl, cluster, U
    lat = torch.randn(B, 64, Tn // 640, device="cuda"); g = torch.randint(0, 153, (B, 1), device="cuda")
    u = torch.rand(Tn, B, device="cuda")
    m.incremental_forward(c=lat[:, :, :1], g=g, T=640, uniforms=u[:640], return_indices=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    idx = m.incremental_forward(c=lat, g=g, T=Tn, uniforms=u, return_indices=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"B={B:3d} T={Tn} {prec}/{impl} cluster={cluster} U={U}: {1e3*ms/Tn:7.1f} us/step  {B*Tn/ms*1e3:10.0f} samples/s  RTF/utt {Tn/ms*1e3/16000:.2f}  "
          f"(classes used: {idx.unique().numel()})", flush=True)

if __name__ == "__main__":
    run(8, 1280, "bf16", "mma", 8, 8)
    run(32, 1280, "bf16", "mma", 8, 8)
    run(64, 1280, "bf16", "mma", 8, 8)
    run(32, 1280, "bf16", "mma", 8, 4)
    run(32, 1280, "bf16", "mma", 16, 8)
    run(32, 1280, "bf16", "simt", 8, 4)
    run(32, 1280, "fp32", "simt", 16, 4)
    run(32, 1280, "fp32", "simt", 16, 2)
