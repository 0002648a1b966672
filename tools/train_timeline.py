"""Kernel timeline of ONE replay of the graphed VQ-WAE training step (8 x 7680): busy time (union over streams), idle gaps on the
device with the kernels around them, time per kernel family, and when each stream is active.  GPU box only; CUPTI traces graph
replays kernel by kernel."""
import json, os, sys, tempfile, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from wavenet_autoencoders_b200 import train_step as TS
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
rs = np.random.RandomState(7); Bt, Tt = 8, 7680
ti = torch.tensor(rs.randint(0, 256, size=(Bt, Tt)), dtype=torch.long, device=dev)
tmf = torch.tensor(rs.normal(size=(Bt, 39, Tt // 160)), dtype=torch.float32, device=dev)
tg = torch.tensor(rs.randint(0, 153, size=(Bt, 1)), dtype=torch.long, device=dev)
tm = bench.build_vqvae(dev).train(); tm.wavenet.precision = "bf16"; tm.wavenet.train_impl = "kernels"
opt = TS.FlatAdam(tm)
gs = TS.GraphedTrainStep(tm, opt, ti, tmf, tg)
for _ in range(5): gs(ti, tmf, tg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs(ti, tmf, tg)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} device activities over {(t1 - t0) / 1e3:.3f} ms")
# union busy time + gaps
busy, cur_end, gaps, last = 0.0, t0, [], None
for e in ev:
    s, d = e["ts"], e["dur"]
    if s > cur_end:
        gaps.append((s - cur_end, cur_end - t0, last["name"][:60] if last else "", e["name"][:60]))
        busy += d; cur_end = s + d
    elif s + d > cur_end:
        busy += s + d - cur_end; cur_end = s + d
    if last is None or s + d >= cur_end: last = e
print(f"device busy (union) {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms in {len(gaps)} gaps")
for g in sorted(gaps, reverse=True)[:25]:
    print(f"  gap {g[0]:7.1f} us at {g[1] / 1e3:6.3f} ms   after {g[2]}  ->  before {g[3]}")
fam = collections.defaultdict(lambda: [0.0, 0])
for e in ev:
    n = e["name"]
    for key in ("bwd_gemm", "wgrad_kernel", "layer_bf16", "head_bf16", "colsum", "enc_conv", "enc_wgrad", "enc_reduce", "upsample", "weight_norm", "vq_", "adam", "sumsq",
                "ce_grad", "nll", "first_conv", "transpose_cast", "elementwise", "Memcpy", "Memset", "CatArray", "reduce_kernel", "index"):
        if key in n:
            fam[key][0] += e["dur"]; fam[key][1] += 1; break
    else:
        fam[n[:50]][0] += e["dur"]; fam[n[:50]][1] += 1
for k, (d, c) in sorted(fam.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"  {k:52s} {d / 1e3:7.3f} ms  {c:4d} launches")
# coarse phases: first / last occurrence of marker kernels
def span(key):
    xs = [e for e in ev if key in e["name"]]
    return ((xs[0]["ts"] - t0) / 1e3, (xs[-1]["ts"] + xs[-1]["dur"] - t0) / 1e3) if xs else None
for key in ("enc_conv", "layer_bf16", "head_bf16", "bwd_gemm", "wgrad_kernel", "colsum", "upsample_stage_bwd", "enc_wgrad", "weight_norm_bwd", "adam"):
    print(f"  {key:22s} active {span(key)}")
# 100-us buckets: tensor-kernel occupancy of time
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r2_train_timeline_events.txt"), "w") as f:
    for e in ev:
        f.write(f"{(e['ts'] - t0) / 1e3:8.4f} {e['dur']:7.1f} s{e.get('args', {}).get('stream', '?'):<4} {e['name'][:110]}\n")
