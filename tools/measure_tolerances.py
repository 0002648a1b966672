"""Prints the measured bf16 errors the GPU parity tests bound (so that the asserted bounds can be kept at ~2x measured). GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import build_model, load_golden, rel_err
from wavenet_autoencoders_b200 import testing as T, training

for case in ["wavenet_tiny", "wavenet_tiny_k2", "wavenet_vqwae", "wavenet_inwae", "wavenet_vqwae_b2", "wavenet_inwae_b2"]:
    g = load_golden(case)
    cfg = T.CONFIGS[str(g["cfg"])]
    m = build_model(str(g["cfg"]), int(g["seed"]), "cuda")
    x, idx, c, spk = T.synth_inputs(cfg, int(g["B"]), int(g["T"]), int(g["in_seed"]))
    m.precision = "bf16"
    with torch.no_grad():
        y = m(x.cuda(), c.cuda(), spk.cuda())
    s = int(g["stride"]) if "stride" in g else 1
    key = "logits" if "logits" in g else "inc_logits"
    ref = g[key]
    yy = y[:, :, ::s].cpu().numpy()
    if ref.shape != yy.shape:
        print(case, "shape mismatch", ref.shape, yy.shape); continue
    print(f"{case}: bf16 forward rel err {rel_err(yy, ref):.3e}")

def per_tensor(cfg_name):
    cfg = T.CONFIGS[cfg_name]
    m = build_model(cfg_name, 3, "cuda").train()
    B, Tn = (3, 320) if cfg_name != "vqwae" else (2, 1280)
    x, idx, c, spk = T.synth_inputs(cfg, B, Tn, 11)
    x, idx, spk = x.cuda(), idx.cuda(), spk.cuda()
    m.precision = "bf16"
    out = {}
    for impl in ("autograd", "kernels"):
        m.train_impl = impl
        m.zero_grad(set_to_none=True)
        cc = c.cuda().clone().requires_grad_(True)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            y = m(x, cc, spk)
            loss = torch.nn.functional.cross_entropy(y[:, :, :-1], idx[:, 1:])
            loss.backward()
        out[impl] = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        out[impl]["<dc>"] = cc.grad.clone()
    worst = []
    gmax = max(float(v.abs().max()) for v in out["autograd"].values())
    for n in out["autograd"]:
        a, b = out["autograd"][n].double().flatten(), out["kernels"][n].double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
        l2 = float((a - b).norm() / (a.norm() + 1e-300))
        worst.append((cos, l2, float(a.norm()), n))
    worst.sort()
    print(cfg_name, "per-tensor end-to-end (b): lowest cosines:")
    for cos, l2, nrm, n in worst[:8]:
        print(f"   cos {cos:.4f} relL2 {l2:.3f} |g| {nrm:.3e}  {n}")
    print("   by kind: " + "; ".join(f"{k}: min cos {min(c for c, _, _, n in worst if k in n):.4f} max relL2 {max(l for _, l, _, n in worst if k in n):.3f}"
                                      for k in ("bias", "weight_g", "weight_v", "<dc>", "embed")))
for cfg_name in ("tiny", "tiny_k2", "vqwae"):
    per_tensor(cfg_name)
