"""Phase breakdown of the AR cluster kernel (cycle counters of thread 0 of every CTA). GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_autoencoders_b200 import _lib, testing as T
from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet

NAMES_SIMT = ["x0/firstconv", "xr-load", "wait-W1", "gemv1", "wait-W2", "cluster-wait-1", "gemv2", "cpasync-wait", "cluster-sync-2",
              "head", "sample", "-"]
NAMES_MMA = ["x0/firstconv", "prefetch+gb", "wait-W1", "mma1+sync", "reduce+gate", "allgather-h", "barrier-1", "mma2+sync",
             "reduce2+allgather-x", "barrier-2+ring", "head", "sample"]

def run(cfg_name, B, Tn, prec, cluster, U, impl='simt'):
    cfg = T.CONFIGS[cfg_name]
    torch.manual_seed(0)
    m = WaveNet(**cfg).eval()
    m.load_state_dict(T.synth_state_dict(m, 1))
    m = m.cuda()
    m.precision, m.ar_cluster, m.ar_utts_per_cluster, m.ar_impl = prec, cluster, U, impl
    NAMES = NAMES_MMA if (impl == 'mma' and prec == 'bf16') else NAMES_SIMT
    hop = T.hop(cfg)
    lat = torch.randn(B, cfg["cin_channels"], Tn // hop, device="cuda")
    g = torch.randint(0, cfg["n_speakers"], (B, 1), device="cuda")
    u = torch.rand(Tn, B, device="cuda")
    init = torch.zeros(B, cfg["out_channels"], 1, device="cuda"); init[:, 5] = 1
    m.incremental_forward(initial_input=init, c=lat, g=g, T=Tn, uniforms=u, return_indices=True)   # warm-up
    nct = ((B + U - 1) // U) * cluster
    buf = torch.zeros(nct * 16, dtype=torch.int64, device="cuda")
    _lib.lib().wae_ar_set_profile_buffer(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    m.incremental_forward(initial_input=init, c=lat, g=g, T=Tn, uniforms=u, return_indices=True)
    e1.record(); torch.cuda.synchronize()
    _lib.lib().wae_ar_set_profile_buffer(None)
    ms = e0.elapsed_time(e1)
    p = buf.view(nct, 16)[:, :12].double().cpu()
    tot = p.sum(1)
    print(f"--- {cfg_name} B={B} T={Tn} {prec} cluster={cluster} U={U}: {ms:.2f} ms total, {1e3*ms/Tn:.1f} us/step, "
          f"{B*Tn/ms*1e3:.0f} samples/s; cycles/step (CTA0) = {tot[0].item()/Tn:.0f}, max over CTAs {tot.max().item()/Tn:.0f}")
    for i, n in enumerate(NAMES):
        print(f"    {n:16s} CTA0 {p[0, i].item()/Tn:9.0f} cyc/step   mean {p[:, i].mean().item()/Tn:9.0f}   max {p[:, i].max().item()/Tn:9.0f}")

if __name__ == "__main__":
    run("vqwae", 8, 640, "bf16", 8, 8, "mma")
    run("vqwae", 64, 640, "bf16", 8, 8, "mma")
    run("vqwae", 2, 640, "fp32", 16, 2)
