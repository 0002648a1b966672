"""One short AR run at the vqwae shape (ncu target). GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_autoencoders_b200 import testing as T
from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
cfg = T.CONFIGS["vqwae"]
torch.manual_seed(0)
m = WaveNet(**cfg).eval(); m.load_state_dict(T.synth_state_dict(m, 1)); m = m.cuda()
m.precision, m.ar_cluster, m.ar_utts_per_cluster = "bf16", 8, 8
B, Tn = 8, 640
lat = torch.randn(B, 64, 1, device="cuda"); g = torch.randint(0, 153, (B, 1), device="cuda")
for _ in range(2):
    m.incremental_forward(c=lat, g=g, T=Tn, uniforms=torch.rand(Tn, B, device="cuda"), return_indices=True)
torch.cuda.synchronize(); print("ok")
