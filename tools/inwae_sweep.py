"""BASELINE configs[4] / SURVEY 8(d) C5: IN-WAE decoder (hps/inae_hp.json; 20x2 file-true and the 30x3 variant) teacher-forced
forward roofline sweep over B x T, bf16 tensor-core stack (and the fp32-faithful stack on a subset).  GPU box only.

    python tools/inwae_sweep.py [--quick] > profiles/rN_inwae_sweep.txt
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_autoencoders_b200 import testing as T
from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def flops_per_sample(cfg):
    R, G, S, C, O, L, kw = (cfg["residual_channels"], cfg["gate_channels"], cfg["skip_out_channels"], cfg["cin_channels"],
                            cfg["out_channels"], cfg["layers"], cfg["kernel_size"])
    H = G // 2
    return L * (2 * kw * R * G + 2 * C * G + 2 * H * R + 2 * H * S) + 2 * S * S + 2 * S * O


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    peak = 1402.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    print(f"# IN-WAE decoder teacher-forced forward, one B200; peak = {peak} TFLOP/s (sustained bf16)")
    print(f"{'config':10s} {'prec':5s} {'B':>3s} {'T':>6s} {'ms':>9s} {'samples/s':>12s} {'TFLOP/s':>8s} {'frac':>6s}")
    for name in ("inwae", "inwae_30x3"):
        cfg = T.CONFIGS[name]
        torch.manual_seed(0)
        m = WaveNet(**cfg).eval()
        m.load_state_dict(T.synth_state_dict(m, 1))
        m = m.cuda()
        fl = flops_per_sample(cfg)
        hop = T.hop(cfg)
        grid = [(64, 32000)] if a.quick else [(b, t) for b in (1, 4, 16, 64) for t in (8000, 16000, 32000)]
        for prec in ("bf16", "fp32"):
            for B, Tn in grid:
                if prec == "fp32" and not (B, Tn) in ((1, 16000), (16, 16000)):
                    continue
                Tn = Tn // hop * hop
                idx = torch.randint(0, 256, (B, Tn), device="cuda")
                x = torch.zeros(B, 256, Tn, device="cuda").scatter_(1, idx[:, None, :], 1.0)
                c = torch.randn(B, 64, Tn // hop, device="cuda")
                g = torch.randint(0, 153, (B, 1), device="cuda")
                m.precision = prec
                with torch.no_grad():
                    for _ in range(2):
                        y = m(x, c, g)
                    del y
                    reps = 5 if B * Tn >= 256000 else 20
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize(); e0.record()
                    for _ in range(reps):
                        y = m(x, c, g)
                        del y
                    e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                tf = B * Tn * fl / (ms * 1e-3) / 1e12
                print(f"{name:10s} {prec:5s} {B:3d} {Tn:6d} {ms:9.3f} {B * Tn / ms * 1e3:12.0f} {tf:8.1f} {tf / peak:6.3f}", flush=True)
                del x, c, idx
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
