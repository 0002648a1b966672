"""Where one VQ-WAE training step spends its time (GPU box only): CUDA kernel time vs wall time, top kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from wavenet_autoencoders_b200 import train_step as TS

dev = torch.device("cuda:0")
impl = sys.argv[1] if len(sys.argv) > 1 else "kernels"
torch.backends.cudnn.benchmark = True
tm = bench.build_vqvae(dev).train()
tm.wavenet.precision = "bf16"
tm.wavenet.train_impl = impl
opt = TS.FlatAdam(tm) if impl == "kernels" else TS.make_optimizer(tm)
rs = np.random.RandomState(7)
Bt, Tt = 8, 7680
ti = torch.tensor(rs.randint(0, 256, size=(Bt, Tt)), dtype=torch.long, device=dev)
tmf = torch.tensor(rs.normal(size=(Bt, 39, Tt // 160)), dtype=torch.float32, device=dev)
tg = torch.tensor(rs.randint(0, 153, size=(Bt, 1)), dtype=torch.long, device=dev)
for _ in range(3):
    TS.train_step(tm, opt, ti, tmf, tg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    TS.train_step(tm, opt, ti, tmf, tg)
e1.record(); torch.cuda.synchronize()
print(f"impl={impl}: {e0.elapsed_time(e1) / 5:.2f} ms per step (wall, GPU events)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    TS.train_step(tm, opt, ti, tmf, tg)
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.self_device_time_total for e in ev) / 1e3
n = sum(e.count for e in ev if e.self_device_time_total > 0)
print(f"CUDA kernel time {tot:.2f} ms over {n} launches")
print(ev.table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=70), flush=True)
if impl == "kernels" and os.environ.get("WAE_PROFILE_GRAPH", "0") == "1":   # capture after a profiler session fails (legacy-stream dependency); run separately
    opt2 = TS.FlatAdam(tm)
    gs = TS.GraphedTrainStep(tm, opt2, ti, tmf, tg)
    for _ in range(2):
        gs(ti, tmf, tg)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        loss = gs(ti, tmf, tg)
    e1.record(); torch.cuda.synchronize()
    print(f"CUDA-graph replay: {e0.elapsed_time(e1) / 10:.2f} ms per step, loss {float(loss):.4f}")
