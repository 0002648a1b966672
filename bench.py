"""bench.py -- WaveNet-decoder samples/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of the workload BASELINE.json quotes the metric on
(configs[1]): VQ-WAE (hps/vqwae.json) encoder -> VQ -> WaveNet decoder teacher-forced forward, 16 utterances x 1 s
of synthetic 16 kHz audio per GPU, with synthetic MFCC conditioning and random-init weights.

  value     device-timed samples/s of the reference-shaped call (one-hot input resident in HBM, eager launches, logits
            written), tcgen05 bf16 kernels (weak scaling: 16 utt per GPU)
  e2e       the same metric through the public module API from pinned HOST buffers: H2D of the mu-law class indices, MFCCs
            and speaker ids, VQVAE.forward_nll (fused encoder+VQ kernel, front-end kernel, 20 layer kernels, head kernel
            with the NLL from its accumulator), D2H of the loss -- all inside the timing, replayed as one CUDA graph
  roofline  the residual-layer kernel (layer_bf16_v4_kernel): algorithmic FLOPs / CUDA-event time vs the measured BURST bf16 peak
  cpu_baseline   the UNMODIFIED reference modules (baseline/_ref, staged by __graft_entry__.build()) on the host cores, bounded
            sample (one utterance per step); the torch port of the same ATen calls if baseline/_ref is absent
  extras    fp32-faithful stack, autoregressive synthesis (config 4 per-GPU shard: 32 x 48000), IN-WAE forward, VQ search
            throughput, the data-parallel training step, the reference on the same GPU (eager cuDNN / cuBLAS)

--impl reference times the reference's CPU path alone (its modules are pure Python/PyTorch; /root/reference does not exist
on the GPU box, baseline/_ref travels with the snapshot; see DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"] or "--impl=reference" in sys.argv:
    # the reference arm times the reference's CPU path: with no visible device its modules take their CPU branches unmodified
    # (vector_quantization.py:33 moves the one-hot matrix to CUDA whenever torch.cuda.is_available())
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "WaveNet-decoder samples/s (teacher-forced forward)"
UNIT = "samples/s"
B_PER_GPU, T_SAMPLES, FRAMES = 16, 16000, 100          # configs[1]: 16 x 1 s; 100 MFCC frames -> 25 latents -> 16000
FLOP_PER_SAMPLE_LAYER = 557056                         # 2kRG + 2CG + 2HR + 2HS at hps/vqwae.json
FLOP_PER_SAMPLE_SKIP = 2 * 128 * 256                   # the layer's skip 1x1 (2HS), executed inside the head kernel's K = L*H GEMM
FLOP_PER_SAMPLE_TOTAL = 11403264                       # 20 layers + head (first conv on one-hot input = gather)


def LAYER_KERNEL_NAME(L):
    try:
        return L.wae_layer_kernel_name().decode()
    except Exception:
        return "layer_bf16_v2_kernel"


def workload_config(world):
    return {"workload": "VQ-WAE hps/vqwae.json encoder->VQ->WaveNet decoder teacher-forced forward (BASELINE configs[1])",
            "batch_per_gpu": B_PER_GPU, "samples_per_utt": T_SAMPLES, "global_batch": world * B_PER_GPU,
            "parallelism": f"utterance-sharded x{world}, no data-path collective",
            "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush",
            "weights": "synthetic seeded (testing.synth_state_dict)",
            "reference_arm": "--impl reference / cpu_baseline: the unmodified reference modules (baseline/_ref) on the host cores, the "
                             "same VQVAE.forward + teacher-forced NLL on ONE utterance per step (1/16 of the per-GPU batch; "
                             "decoder-only port of the same ATen calls if baseline/_ref is absent) -- see cpu_baseline.sample",
            "value_input": "reference-shaped call, eager launches: one-hot (B,256,T) fp32 + mfcc + speaker ids resident in HBM; "
                           "e2e goes through the class-index input and GraphedForward (see e2e.what)"}


def _peaks():
    """(burst bf16 TFLOP/s, sustained bf16 TFLOP/s, HBM GB/s, source).  The layer kernel is timed in windows of tens of ms at
    full clocks, so its roofline denominator is the BURST figure (B200_PROFILING.md); the sustained one is reported beside it."""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops"]), float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst bf16)"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md: 1.59 PF burst, ~1.4 PF sustained, 6.65 TB/s)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, f"/tmp/wae_clocks_{os.getpid()}.csv"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s in sm if s > 0]
            out = {"sm_mhz": float(np.median(busy[len(busy) // 4:] or busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}
        return out


def build_vqvae(device):
    from wavenet_autoencoders_b200 import testing as T
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(0)
    wn = WaveNet(**T.VQWAE)
    m = VQVAE(c_in=39, hid=64, K=256, wavenet=wn, encoder_hid=256).eval()     # vqwae_train.py:946 with hps/vqwae.json
    m.load_state_dict(T.synth_state_dict(m, 1))
    return m.to(device)


def synth_batch(B, seed):
    rs = np.random.RandomState(seed)
    idx = torch.tensor(rs.randint(0, 256, size=(B, T_SAMPLES)), dtype=torch.long)
    mfcc = torch.tensor(rs.normal(size=(B, 39, FRAMES)), dtype=torch.float32)
    g = torch.tensor(rs.randint(0, 153, size=(B, 1)), dtype=torch.long)
    return idx, mfcc, g


def load_reference_modules():
    """The UNMODIFIED reference modules staged in git-ignored baseline/_ref/ by __graft_entry__.build() (they travel to the GPU
    box with the snapshot).  Returns (VQVAE, WaveNet) or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "wavenet_vocoder")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import warnings
    warnings.filterwarnings("ignore")
    try:
        import vqvae_model as ref_vqvae
        from wavenet_vocoder import WaveNet as RefWaveNet
    except Exception:
        return None
    return ref_vqvae.VQVAE, RefWaveNet


def build_reference_vqvae(device):
    """The reference's own VQVAE (vqvae_model.py:54-71 around wavenet_vocoder.WaveNet) with the weights of build_vqvae."""
    mods = load_reference_modules()
    if mods is None:
        return None
    from wavenet_autoencoders_b200 import testing as T
    RefVQVAE, RefWaveNet = mods
    torch.manual_seed(0)
    m = RefVQVAE(c_in=39, hid=64, K=256, wavenet=RefWaveNet(**T.VQWAE), encoder_hid=256).eval()
    m.load_state_dict(T.synth_state_dict(m, 1))
    return m.to(device)


def cpu_reference_setup():
    """One step of the reference arm: the workload of the GPU arm's e2e (VQVAE.forward -> teacher-forced NLL,
    vqwae_train.py:760-766) on ONE utterance (1/16 of the per-GPU batch), on the host cores.  kind "reference": the real
    modules from baseline/_ref; kind "port": oracle/torch_port.py, decoder only (when baseline/_ref is absent)."""
    m = build_reference_vqvae("cpu")
    if m is None:
        return cpu_port_setup(), "port", ("decoder teacher-forced forward only (no encoder / VQ / loss), B=1 x T=%d per step "
                                          "(1/16 of the per-GPU batch), oracle/torch_port.py, fp32" % T_SAMPLES)
    idx, mfcc, g = synth_batch(1, 100)
    x = torch.nn.functional.one_hot(idx, 256).float().transpose(1, 2).contiguous()

    def step():
        with torch.no_grad():
            y, vq_loss, perp = m(x, mfcc, g)
            return float(torch.nn.functional.cross_entropy(y[:, :, :-1], idx[:, 1:]))
    return step, "reference", ("unmodified reference modules (baseline/_ref): VQVAE.forward (encoder -> VQ -> WaveNet decoder, "
                               "one-hot input) + cross_entropy, B=1 x T=%d per step (1/16 of the per-GPU batch), fp32, torch %s CPU"
                               % (T_SAMPLES, torch.__version__))


def cpu_port_setup():
    """Decoder of the same model on the host cores through oracle/torch_port.py (kind "port")."""
    from oracle import torch_port
    from oracle import wavenet_oracle as wo
    from wavenet_autoencoders_b200 import testing as T
    m = build_vqvae("cpu")
    sd = {k[len("wavenet."):]: v.numpy() for k, v in m.state_dict().items() if k.startswith("wavenet.")}
    p = wo.extract_params(sd, T.VQWAE["layers"], T.VQWAE["stacks"])
    tp = torch_port.params_from_numpy(p)
    idx, mfcc, g = synth_batch(1, 100)
    x = torch.nn.functional.one_hot(idx, 256).float().transpose(1, 2).contiguous()
    with torch.no_grad():
        quant = torch.randn(1, 64, FRAMES // 4)
        c_up = torch_port.upsample(tp, quant)
        gv = torch.tensor(wo.speaker_vectors(p, g.numpy()))

    def step():
        with torch.no_grad():
            return torch_port.stack_forward(tp, x, c_up, gv)
    return step


def library_baseline(dev, x, idx, mfcc, g):
    """The unmodified reference modules (baseline/_ref) moved to the GPU: eager PyTorch, cuDNN convolutions + cuBLAS, fp32 with
    TF32 off and on.  Same weights and the same tensors as the headline step (16 x 16000, one-hot input) + cross_entropy."""
    m = build_reference_vqvae(dev)
    if m is None:
        return {"unavailable": "baseline/_ref not staged (run __graft_entry__.build() where /root/reference exists)"}
    out = {"what": "reference VQVAE.forward(one-hot x, mfcc, g) + F.cross_entropy on the B200, eager launches, B=16 x T=16000",
           "torch": torch.__version__}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32, benchmark=True), torch.no_grad():
            for _ in range(2):
                y = m(x, mfcc, g)[0]
                loss = torch.nn.functional.cross_entropy(y[:, :, :-1], idx[:, 1:])
            torch.cuda.synchronize()
            n = 5
            e0.record()
            for _ in range(n):
                y = m(x, mfcc, g)[0]
                loss = torch.nn.functional.cross_entropy(y[:, :, :-1], idx[:, 1:])
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        key = "tf32" if tf32 else "fp32"
        out[f"{key}_ms_per_step"] = ms
        out[f"{key}_samples_per_s"] = x.shape[0] * x.shape[2] / (ms * 1e-3)
        out[f"{key}_loss"] = float(loss)
        del y
    torch.backends.cuda.matmul.allow_tf32 = False
    return out


def cpu_baseline_subprocess():
    """The cpu_baseline leg = the reference arm itself, run as a child process with the GPU hidden (the unmodified reference
    picks its device by torch.cuda.is_available()).  About 10-30 s of CPU work."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "12", "--warmup", "1"],
                           capture_output=True, text=True, timeout=900, env={k: v for k, v in os.environ.items()
                                                                             if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, sample = cpu_reference_setup()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * T_SAMPLES / dt
    cfg = workload_config(args.gpus)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ar-steps", type=int, default=48000, help="AR extra: samples per utterance (configs[3]: 48000 = 3 s)")
    ap.add_argument("--ar-fp32-steps", type=int, default=2560, help="AR extra, fp32-faithful SIMT kernel: samples per utterance (truncated)")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # which algorithm / protocol NCCL picks for the gradient all-reduce: its INFO log goes to a file that rank 0 summarises
        # into extras.nccl (the stdout of this process must carry the one JSON line only)
        nccl_log = f"/tmp/wae_nccl_{os.getpid()}.log"
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,TUNING,COLL" if os.environ.get("WAE_NCCL_COLL") else "INIT,TUNING"
        os.environ["NCCL_DEBUG_FILE"] = nccl_log
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry the one JSON line only
        sys.stdout.flush()
        saved_out = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_out, 1)
            os.close(saved_out)

    from wavenet_autoencoders_b200 import _lib
    L = _lib.lib()
    model = build_vqvae(dev)
    model.wavenet.precision = "bf16"
    B = B_PER_GPU                                               # weak scaling: every rank runs its own 16 utterances
    idx_h, mfcc_h, g_h = synth_batch(B, 1000 + rank)
    idx_p, mfcc_p, g_p = idx_h.pin_memory(), mfcc_h.pin_memory(), g_h.pin_memory()
    idx, mfcc, g = idx_p.to(dev), mfcc_p.to(dev), g_p.to(dev)
    x = torch.nn.functional.one_hot(idx, 256).float().transpose(1, 2).contiguous()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_device():
        with torch.no_grad():
            return model(x, mfcc, g)[0]

    from wavenet_autoencoders_b200.losses import teacher_forced_nll

    def step_e2e():
        # the package's own public call on what the data pipeline holds: class indices go in as they are (the first conv
        # gathers rows; no (B,256,T) one-hot is built) and the NLL is one pass over the logits
        with torch.no_grad():
            i_d = idx_p.to(dev, non_blocking=True)
            m_d = mfcc_p.to(dev, non_blocking=True)
            g_d = g_p.to(dev, non_blocking=True)
            loss = model.forward_nll(i_d, m_d, g_d, i_d, 1)[0]                       # vqwae_train.py:760-766, from the head kernel's accumulator
            return float(loss.item())                                                # D2H of the step's result

    def step_e2e_onehot():
        # the same through the reference-shaped call: one-hot float input, torch cross_entropy on the shifted slices
        with torch.no_grad():
            i_d = idx_p.to(dev, non_blocking=True)
            m_d = mfcc_p.to(dev, non_blocking=True)
            g_d = g_p.to(dev, non_blocking=True)
            xo = torch.nn.functional.one_hot(i_d, 256).float().transpose(1, 2).contiguous()
            y = model(xo, m_d, g_d)[0]
            loss = torch.nn.functional.cross_entropy(y[:, :, :-1], i_d[:, 1:])
            return float(loss.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    clocks = ClockSampler(local) if rank == 0 else None
    n0 = _lib.launch_count()
    ms = timed(step_device, args.steps, args.warmup)
    launches = (_lib.launch_count() - n0) * args.steps // (args.steps + args.warmup)
    clk = clocks.stop() if clocks else None
    value = world * B * T_SAMPLES * args.steps / (ms * 1e-3)

    ms_e2e_eager = timed(step_e2e, args.steps, args.warmup)
    # the same call replayed as one CUDA graph (GraphedForward: same kernels, captured once; the eager call's ~75 launches are
    # issued from Python after each step's loss read-back, more slowly than the frame-rate kernels at the front execute)
    graphed, graphed_lg, graphed_err = None, None, None
    try:
        from wavenet_autoencoders_b200.graphed import GraphedForward
        graphed = GraphedForward(model, idx, mfcc, g, with_logits=False)      # loss-only: the logits are never materialised
        graphed_lg = GraphedForward(model, idx, mfcc, g)                      # the variant that also returns the logits
        ref_loss = step_e2e()
        got_loss = float(graphed(idx_p, mfcc_p, g_p)[3].item())
        lg_loss = float(graphed_lg(idx_p, mfcc_p, g_p)[3].item())             # one-pass NLL over the written logits
        if not (abs(got_loss - ref_loss) <= 1e-6 * max(1.0, abs(ref_loss)) and abs(got_loss - lg_loss) <= 2e-6 * max(1.0, abs(lg_loss))):
            raise RuntimeError(f"graph replay loss {got_loss} != eager loss {ref_loss} / loss from the logits {lg_loss}")
    except Exception as e:                                       # reported in the JSON line; the eager number stands then
        graphed, graphed_lg, graphed_err = None, None, f"{type(e).__name__}: {e}"[:300]
        torch.cuda.synchronize()

    if dist is not None:                                         # every rank must take the same branch: timed() holds collectives
        flag = torch.tensor([1 if graphed is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and graphed is not None:
            graphed, graphed_err = None, "graph capture failed on another rank"

    def step_e2e_graphed():
        return float(graphed(idx_p, mfcc_p, g_p)[3].item())      # H2D copies into the captured buffers, replay, D2H of the loss

    ms_e2e = timed(step_e2e_graphed, args.steps, args.warmup) if graphed is not None else ms_e2e_eager
    e2e_value = world * B * T_SAMPLES * args.steps / (ms_e2e * 1e-3)

    # the same step when the caller wants the LOGITS back (262 MB fp32 per step): D2H into pinned host memory inside the timing
    logits_host = torch.empty(B, 256, T_SAMPLES, dtype=torch.float32).pin_memory()

    def step_e2e_logits():
        if graphed is not None:
            out = graphed_lg(idx_p, mfcc_p, g_p)
            logits_host.copy_(out[0], non_blocking=True)
            return float(out[3].item())
        with torch.no_grad():
            i_d = idx_p.to(dev, non_blocking=True)
            y = model(i_d, mfcc_p.to(dev, non_blocking=True), g_p.to(dev, non_blocking=True))[0]
            logits_host.copy_(y, non_blocking=True)
            return float(teacher_forced_nll(y, i_d).item())
    n_lg = max(3, args.steps // 4)
    ms_e2e_logits = timed(step_e2e_logits, n_lg, 2) / n_lg
    ms_e2e_onehot = timed(step_e2e_onehot, args.steps, args.warmup)
    h2d = idx_p.numel() * 8 + mfcc_p.numel() * 4 + g_p.numel() * 8

    # ---- roofline of the dominant kernel, measured live with CUDA events around each launch ----
    import ctypes
    L.wae_profile_enable(1)
    for _ in range(3):
        step_device()
    torch.cuda.synchronize()
    ms_kind = (ctypes.c_float * 4)()
    n_kind = (ctypes.c_int32 * 4)()
    L.wae_profile_read(ms_kind, n_kind, 4)      # discard warm-up
    prof_steps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(prof_steps):
        step_device()
    e1.record()
    torch.cuda.synchronize()
    L.wae_profile_read(ms_kind, n_kind, 4)
    L.wae_profile_enable(0)
    prof_total = e0.elapsed_time(e1)
    layer_launches = max(int(n_kind[1]), 1)
    layer_ms = float(ms_kind[1]) / layer_launches
    n_layers = 20
    # FLOPs one layer launch EXECUTES: the layer's 2kRG + 2CG + 2HR.  Its skip 1x1 (2HS = 65,536 of SURVEY 8(d)'s 557,056 per
    # sample per layer) runs in head_bf16_kernel as part of the K = L*H skip GEMM and is credited there, not here; the last
    # layer has no residual 1x1.
    flops_per_launch = B * T_SAMPLES * (FLOP_PER_SAMPLE_LAYER - FLOP_PER_SAMPLE_SKIP - 2 * 128 * 256 / n_layers)
    head_flops = B * T_SAMPLES * (n_layers * FLOP_PER_SAMPLE_SKIP + 2 * 256 * 256 + 2 * 256 * 256)
    peak_tf, peak_tf_sust, peak_hbm, peak_src = _peaks()
    achieved = flops_per_launch / (layer_ms * 1e-3) / 1e12
    roofline = {"kernel": LAYER_KERNEL_NAME(L), "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": None, "peak_source": peak_src,
                "frac_of_sustained_peak": achieved / peak_tf_sust, "sustained_peak": peak_tf_sust,
                "avg_launch_ms": layer_ms, "share_of_step": float(ms_kind[1]) / prof_total,
                "head_share_of_step": float(ms_kind[2]) / prof_total, "prep_share_of_step": float(ms_kind[0]) / prof_total,
                "step_frac_of_peak": value / world * FLOP_PER_SAMPLE_TOTAL / 1e12 / peak_tf,
                "flops_counted": "per launch: B*T*(2kRG + 2CG + 2HR) = 491,520/sample; the layer's skip 1x1 (65,536/sample) is "
                                 "executed by and credited to the head kernel",
                "head_frac": head_flops / (float(ms_kind[2]) / max(int(n_kind[2]), 1) * 1e-3) / 1e12 / peak_tf,
                "stack_frac": B * T_SAMPLES * FLOP_PER_SAMPLE_TOTAL * prof_steps
                              / ((float(ms_kind[0]) + float(ms_kind[1]) + float(ms_kind[2])) * 1e-3) / 1e12 / peak_tf}
    # the same fractions on the e2e path (class-index input: first conv gathered by the front-end kernel; NLL from the head
    # kernel's accumulator, no logits written) -- the decoder stack as the e2e number runs it
    try:
        i_dev, m_dev, g_dev = idx_p.to(dev), mfcc_p.to(dev), g_p.to(dev)
        L.wae_profile_enable(1)
        with torch.no_grad():
            for _ in range(2):
                model.forward_nll(i_dev, m_dev, g_dev, i_dev, 1)
            torch.cuda.synchronize()
            L.wae_profile_read(ms_kind, n_kind, 4)
            for _ in range(prof_steps):
                model.forward_nll(i_dev, m_dev, g_dev, i_dev, 1)
            torch.cuda.synchronize()
        L.wae_profile_read(ms_kind, n_kind, 4)
        L.wae_profile_enable(0)
        roofline["index_nll_path"] = {
            "stack_frac": B * T_SAMPLES * FLOP_PER_SAMPLE_TOTAL * prof_steps
                          / ((float(ms_kind[0]) + float(ms_kind[1]) + float(ms_kind[2])) * 1e-3) / 1e12 / peak_tf,
            "head_frac": head_flops / (float(ms_kind[2]) / max(int(n_kind[2]), 1) * 1e-3) / 1e12 / peak_tf,
            "prep_ms": float(ms_kind[0]) / prof_steps, "layers_ms": float(ms_kind[1]) / prof_steps, "head_ms": float(ms_kind[2]) / prof_steps,
            "what": "prep (gate bias + front-end kernel incl. the first-conv gather) + 20 layer kernels + head kernel in NLL mode, "
                    "every FLOP of SURVEY 8(d)'s 11,403,264 per sample counted once, vs the burst bf16 peak"}
    except Exception as e:      # never let an auxiliary measurement take the line down
        roofline["index_nll_path"] = {"error": str(e)[:200]}
    try:
        roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "layer_kernel_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass

    extras = {}
    if not args.no_extras:
        # fp32-faithful stack (the parity mode of configs[1])
        model.wavenet.precision = "fp32"
        ms32 = timed(step_device, max(2, args.steps // 5), 1)
        extras["fp32_stack_samples_per_s"] = world * B * T_SAMPLES * max(2, args.steps // 5) / (ms32 * 1e-3)
        # autoregressive synthesis, configs[3] AS STATED: 32 utterances per GPU (256 over 8 GPUs) x 48000 samples (3 s),
        # categorical sampling on supplied uniforms, bf16 tensor-core cluster kernel; the fp32-faithful SIMT kernel is ~8x
        # slower and timed on a truncated run (stated)
        wn = model.wavenet
        gar = torch.randint(0, 153, (32, 1), device=dev)
        for prec, Tar in (("bf16", (args.ar_steps // 640) * 640), ("fp32", (args.ar_fp32_steps // 640) * 640)):
            if Tar <= 0:
                continue
            wn.precision = prec
            lat = torch.randn(32, 64, Tar // 640, device=dev)
            u = torch.rand(Tar, 32, device=dev)
            wn.incremental_forward(c=lat[:, :, :1], g=gar, T=640, uniforms=u[:640], return_indices=True)   # warm-up / packing
            barrier()
            e0.record()
            out_ar = wn.incremental_forward(c=lat, g=gar, T=Tar, uniforms=u, return_indices=True)
            e1.record()
            barrier()
            t_ar = max_over_ranks(e0.elapsed_time(e1)) * 1e-3
            extras[f"ar_{prec}_samples_per_s"] = world * 32 * Tar / t_ar
            extras[f"ar_{prec}_realtime_factor_per_utt"] = Tar / t_ar / 16000.0
            extras[f"ar_{prec}_us_per_step"] = 1e6 * t_ar / Tar
            extras[f"ar_{prec}_samples_per_utt"] = Tar
            extras[f"ar_{prec}_kernel"] = "/".join(str(v) for v in wn.last_ar_variant)       # weight type / cluster size / utterances per cluster
            if prec == "bf16":
                # "roofline" of a latency-bound kernel: every step streams the cluster's whole weight set (11.53 MB bf16) L2 -> SM once,
                # and walks 2 dependent DSMEM exchanges per layer + 3 in the head
                wt, cs, upc = wn.last_ar_variant
                n_clusters = -(-32 // upc)
                w_bytes = 20 * (256 * 832 + 512 * 128) * 2 + (256 * 256 + 256 * 256) * 2
                extras["ar_roofline"] = {"bound": "latency (2L+3 dependent cluster exchanges per step) / L2->SM weight stream",
                                         "us_per_step": 1e6 * t_ar / Tar, "realtime_us_per_step": 62.5,
                                         "clusters": n_clusters, "ctas": n_clusters * cs,
                                         "l2_to_sm_gbs": n_clusters * w_bytes / (t_ar / Tar) / 1e9,
                                         "l2_cap_gbs_B300_MICROARCH": 6300 * 1.965,
                                         "dependent_exchanges_per_step": 2 * 20 + 3,
                                         "ns_per_exchange_if_latency_bound": 1e9 * t_ar / Tar / (2 * 20 + 3)}
                assert out_ar.shape == (32, Tar)
            del lat, u, out_ar
        extras["ar_config"] = (f"32 utt/GPU x {(args.ar_steps // 640) * 640} samples bf16 (configs[3]: 256 utterances x 48000 over 8 GPUs); "
                               f"fp32 SIMT kernel truncated to {(args.ar_fp32_steps // 640) * 640} samples; categorical sampling, cluster kernels")
        # VQ search throughput, N sweep (HBM roofline: 4D read + 4D write + 8 B index per vector)
        from wavenet_autoencoders_b200.vector_quantization import VectorQuantize
        vq = VectorQuantize(256, 64).to(dev)
        for N in (400, 1 << 20):
            lat_v = torch.randn(N // 25 if N == 400 else 64, 64, 25 if N == 400 else N // 64, device=dev) * 0.05
            with torch.no_grad():
                for _ in range(3):
                    vq(lat_v)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    vq(lat_v)
                e1.record()
                torch.cuda.synchronize()
            sec = e0.elapsed_time(e1) * 1e-3 / 10
            extras[f"vq_vectors_per_s_N{N}"] = N / sec
            extras[f"vq_hbm_frac_N{N}"] = N * (64 * 4 * 2 + 8) / sec / 1e9 / peak_hbm
        model.wavenet.precision = "bf16"
        # configs[4]: IN-WAE decoder (G = 368) teacher-forced forward, 64 x 2 s per GPU, bf16 tensor-core stack
        # (full B x T sweep: tools/inwae_sweep.py -> profiles/)
        from wavenet_autoencoders_b200 import testing as T
        from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet as _WN
        for name in ("inwae", "inwae_30x3"):
            cfg = T.CONFIGS[name]
            torch.manual_seed(0)
            wn5 = _WN(**cfg).eval()
            wn5.load_state_dict(T.synth_state_dict(wn5, 1))
            wn5 = wn5.to(dev)
            wn5.precision = "bf16"
            hop5 = T.hop(cfg)
            B5, T5 = 64, 32000 // hop5 * hop5
            i5 = torch.randint(0, 256, (B5, T5), device=dev)
            x5 = torch.zeros(B5, 256, T5, device=dev).scatter_(1, i5[:, None, :], 1.0)
            c5 = torch.randn(B5, 64, T5 // hop5, device=dev)
            g5 = torch.randint(0, 153, (B5, 1), device=dev)
            with torch.no_grad():
                for _ in range(2):
                    y5 = wn5(x5, c5, g5)
                    del y5
                barrier()
                e0.record()
                for _ in range(3):
                    y5 = wn5(x5, c5, g5)
                    del y5
                e1.record()
                barrier()
            sec = max_over_ranks(e0.elapsed_time(e1)) * 1e-3 / 3
            H5, L5 = cfg["gate_channels"] // 2, cfg["layers"]
            fl5 = L5 * (2 * 3 * 256 * 2 * H5 + 2 * 64 * 2 * H5 + 2 * H5 * 256 + 2 * H5 * 256) + 2 * 256 * 256 + 2 * 256 * 256
            extras[f"{name}_fwd_samples_per_s"] = world * B5 * T5 / sec
            extras[f"{name}_fwd_frac_of_bf16_peak"] = B5 * T5 * fl5 / sec / 1e12 / peak_tf
            del wn5, x5, c5, i5
            torch.cuda.empty_cache()
        # configs[2]: data-parallel VQ-WAE training step, 8 utterances x 7680-sample windows per GPU (8000 is not a multiple
        # of the encoder's stride x hop), bf16 tcgen05 forward + GEMM backward, fp32 master weights, Adam
        from wavenet_autoencoders_b200 import train_step as TS
        torch.backends.cudnn.benchmark = True            # the (out-of-scope, cuDNN) encoder's default wgrad algorithm is 3 x 1.8 ms
        tm = build_vqvae(dev).train()
        tm.wavenet.precision = "bf16"
        opt = TS.FlatAdam(tm, lr=4e-4, clip=100.0)       # clip + Adam on one flat buffer (f4); the all-reduce works on it directly
        if world > 1:                                    # decoder bucket all-reduced under the upsampler / VQ / encoder backward
            opt.enable_overlap(tm)
        rs = np.random.RandomState(7 + rank)
        Bt, Tt = 8, 7680
        ti = torch.tensor(rs.randint(0, 256, size=(Bt, Tt)), dtype=torch.long, device=dev)
        tmf = torch.tensor(rs.normal(size=(Bt, 39, Tt // 160)), dtype=torch.float32, device=dev)
        tg = torch.tensor(rs.randint(0, 153, size=(Bt, 1)), dtype=torch.long, device=dev)
        nt = 5
        # the whole step (all-reduce included) captured into one CUDA graph -- captured BEFORE any eager training step of this
        # model: autograd binds the parameters' gradient accumulators to the stream of their first use, and the legacy
        # default stream cannot take part in a capture
        gstep = TS.GraphedTrainStep(tm, opt, ti, tmf, tg, world=world)
        for _ in range(2):
            gstep(ti, tmf, tg)
        barrier()
        e0.record()
        for _ in range(nt):
            tloss = gstep(ti, tmf, tg)
        e1.record()
        barrier()
        t_ms = max_over_ranks(e0.elapsed_time(e1)) / nt
        extras["train_samples_per_s"] = world * Bt * Tt / (t_ms * 1e-3)
        extras["train_ms_per_step"] = t_ms
        extras["train_loss_after_steps"] = float(tloss)
        # eager launches of the same step (about 3000 small kernels: host-bound)
        for _ in range(2):
            TS.train_step(tm, opt, ti, tmf, tg, world=world)
        barrier()
        e0.record()
        for _ in range(nt):
            TS.train_step(tm, opt, ti, tmf, tg, world=world)
        e1.record()
        barrier()
        extras["train_eager_ms_per_step"] = max_over_ranks(e0.elapsed_time(e1)) / nt
        if world > 1:            # replicas must hold bit-identical parameters after the steps (same all-reduced gradients everywhere)
            import torch.distributed as _d
            chk = opt.flat_p.double().sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            _d.all_reduce(lo, op=_d.ReduceOp.MIN)
            _d.all_reduce(hi, op=_d.ReduceOp.MAX)
            extras["train_replicas_in_sync"] = bool(float(lo) == float(hi))
        if world > 1:            # all-reduce alone, same flat buffer size (7.6 M fp32 gradients), device-timed
            import torch.distributed as _d
            flat = torch.zeros(sum(p.numel() for p in tm.parameters()), device=dev)
            for _ in range(3):
                _d.all_reduce(flat)
            barrier()
            e0.record()
            for _ in range(10):
                _d.all_reduce(flat)
            e1.record()
            barrier()
            extras["train_allreduce_ms"] = max_over_ranks(e0.elapsed_time(e1)) / 10
            extras["train_allreduce_bytes"] = flat.numel() * 4
            br = getattr(opt, "bucketed", None)
            if br is not None:
                extras["train_allreduce_buckets"] = [int(sl.numel()) * 4 for sl in br.slices]
                extras["train_allreduce_buckets_started_inside_backward"] = int(br.overlapped)
            try:                                         # NCCL's own account of what it runs (INIT / TUNING lines of its INFO log)
                lines = open(os.environ.get("NCCL_DEBUG_FILE", "")).read().splitlines()
                keep = [ln.split("NCCL INFO ", 1)[-1][:160] for ln in lines
                        if any(k in ln for k in ("NVLS", "Connected all", "AllReduce", "Algo", "nRanks", "channels"))]
                extras["nccl"] = keep[:6] + (["..."] if len(keep) > 12 else []) + keep[-6:]
            except OSError as e:
                extras["nccl"] = [f"no NCCL log: {e}"]
        del gstep
        extras["train_config"] = ("VQ-WAE step: 8 utt/GPU x 7680 samples, class-index input (first conv = row gather), tcgen05 bf16 forward "
                                  "keeping tanh / sigmoid of the gates, tcgen05 backward (wae_stack_backward_bf16_2s: dgrad GEMMs with the "
                                  "dilated taps as TMA boxes and fused gate / residual / ReLU epilogues on the main stream, MN-major split-K "
                                  "wgrads and the bias column sums on side streams), loss + softmax gradient fused (forward_nll), encoder and "
                                  "conditioning upsampler forward + backward on this library's fp32 kernels (no cuDNN / cuBLAS convolution in "
                                  "the step), decoder weight preparation forked at the start of the step over 9 side streams, fp32 master "
                                  "weights, Adam, grad-clip 100; N > 1: flat gradient all-reduced in two buckets, the decoder's (80 % of the "
                                  "bytes) under the rest of the backward; the step is replayed as one CUDA graph")
        del tm, opt
        torch.backends.cudnn.benchmark = False
        torch.cuda.empty_cache()
        extras["inwae_config"] = "decoder only, 64 utt/GPU x 32000 samples, bf16 tcgen05 stack (gate rows in two accumulator passes)"

    # ---- the reference's own eager PyTorch path on this GPU (cuDNN / cuBLAS), BASELINE.md section 4: same weights, same tensors ----
    if not args.no_extras and rank == 0:
        try:
            extras["library_baseline"] = library_baseline(dev, x, idx, mfcc, g)
        except Exception as e:
            extras["library_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1:
        cpu = cpu_baseline_subprocess()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "with_logits_to_host": {"value": world * B * T_SAMPLES / (ms_e2e_logits * 1e-3), "ms_per_step": ms_e2e_logits,
                                            "d2h_bytes_per_step": 4 + logits_host.numel() * 4,
                                            "what": "same step, plus the (16,256,16000) fp32 logits copied to pinned host memory"},
                    "what": "pinned host class indices/mfcc/speaker -> H2D -> VQVAE.forward_nll(indices): encoder -> VQ -> decoder with the teacher-forced NLL taken from the head kernel's accumulator (logits never written) -> D2H loss"
                            + (", replayed as one CUDA graph (GraphedForward)" if graphed is not None else ", eager launches"),
                    "eager_api_value": world * B * T_SAMPLES * args.steps / (ms_e2e_eager * 1e-3),
                    "graph_error": graphed_err,
                    "onehot_api_value": world * B * T_SAMPLES * args.steps / (ms_e2e_onehot * 1e-3),
                    "onehot_api_what": "same, reference-shaped call: one-hot (B,256,T) fp32 built on the device + torch cross_entropy"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
        }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
