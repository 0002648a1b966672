"""A very short AR run of the tensor-core kernel at the tiny test shape (target of compute-sanitizer). GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_autoencoders_b200 import testing as T
from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
cfg = T.CONFIGS["tiny"]
torch.manual_seed(0)
m = WaveNet(**cfg).eval(); m.load_state_dict(T.synth_state_dict(m, 1)); m = m.cuda()
m.precision, m.ar_cluster, m.ar_impl = "bf16", int(sys.argv[1]) if len(sys.argv) > 1 else 4, "mma"
B, Tn = 3, 32
lat = torch.randn(B, 16, Tn // T.hop(cfg), device="cuda"); g = torch.randint(0, 5, (B, 1), device="cuda")
init = torch.zeros(B, cfg["out_channels"], 1, device="cuda"); init[:, 3] = 1.0
idx = m.incremental_forward(initial_input=init, c=lat, g=g, T=Tn, uniforms=torch.rand(Tn, B, device="cuda"), return_indices=True)
torch.cuda.synchronize(); print("ok", idx[0, :8].tolist())
