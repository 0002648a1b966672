"""Deterministic model configurations and synthetic weights shared by tools/make_golden.py, tests/ and bench.py.

Weights are a pure function of (state_dict key order, shapes, seed) via numpy, so the real reference model
(in the container that has /root/reference) and this package's model (anywhere) can be loaded with identical
parameters without shipping checkpoints: ``model.load_state_dict(synth_state_dict(model, seed))``.
"""
from __future__ import annotations

import numpy as np
import torch

# WaveNet(**kwargs) per preset.  vqwae: hps/vqwae.json:24-46 as passed by vqwae_train.py:926-944;
# inwae: hps/inae_hp.json:24-45 (file-true 20x2) and the 30x3 variant BASELINE.json names.
VQWAE = dict(out_channels=256, layers=20, stacks=2, residual_channels=256, gate_channels=256,
             skip_out_channels=256, kernel_size=3, dropout=0.0, cin_channels=64, gin_channels=32, n_speakers=153,
             upsample_conditional_features=True, upsample_net="ConvInUpsampleNetwork",
             upsample_params={"upsample_scales": [4, 4, 8, 5], "cin_channels": 64, "cin_pad": 0},
             scalar_input=False, use_speaker_embedding=True, output_distribution="Logistic", cin_pad=0)
INWAE = dict(VQWAE, gate_channels=368, gin_channels=64,
             upsample_params={"upsample_scales": [4, 4, 4, 5], "cin_channels": 64, "cin_pad": 0})
INWAE_30x3 = dict(INWAE, layers=30, stacks=3)
TINY = dict(out_channels=32, layers=4, stacks=2, residual_channels=64, gate_channels=64, skip_out_channels=64,
            kernel_size=3, dropout=0.0, cin_channels=16, gin_channels=8, n_speakers=5,
            upsample_conditional_features=True, upsample_net="ConvInUpsampleNetwork",
            upsample_params={"upsample_scales": [4, 4], "cin_channels": 16, "cin_pad": 0},
            scalar_input=False, use_speaker_embedding=True, output_distribution="Logistic", cin_pad=0)
TINY_K2 = dict(TINY, kernel_size=2, layers=6, stacks=3)
TINY_MOL = dict(TINY, out_channels=6, scalar_input=True)
TINY_GAUSS = dict(TINY, out_channels=6, scalar_input=True, output_distribution="Normal")
CONFIGS = {"vqwae": VQWAE, "inwae": INWAE, "inwae_30x3": INWAE_30x3, "tiny": TINY, "tiny_k2": TINY_K2,
           "tiny_mol": TINY_MOL, "tiny_gauss": TINY_GAUSS}


def hop(cfg) -> int:
    return int(np.prod(cfg["upsample_params"]["upsample_scales"]))


def synth_state_dict(model: torch.nn.Module, seed: int) -> dict:
    """Random but well-scaled parameters for every entry of model.state_dict(), reproducible from `seed`."""
    out = {}
    for i, (name, ref) in enumerate(model.state_dict().items()):
        rs = np.random.RandomState((seed * 1000003 + i * 7919) % (2 ** 31 - 1))
        shape = tuple(ref.shape)
        if not ref.dtype.is_floating_point:
            out[name] = ref.clone()
            continue
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        if name.endswith("weight_g"):
            v = rs.uniform(0.7, 1.3, size=shape)
        elif name.endswith("bias"):
            v = rs.normal(0.0, 0.1, size=shape)
        elif "embed_speakers" in name:
            v = rs.normal(0.0, 0.3, size=shape)
        elif "embedding" in name:                       # VQ codebooks
            v = rs.normal(0.0, 0.5, size=shape)
        elif "up_layers" in name and name.endswith("weight_v"):
            v = rs.uniform(0.5, 1.5, size=shape) / fan_in
        elif name.startswith("ema_") or ".ema_" in name:
            v = np.zeros(shape)
        else:
            v = rs.normal(0.0, 1.0, size=shape) / np.sqrt(fan_in)
        out[name] = torch.tensor(v, dtype=ref.dtype)
    return out


def synth_inputs(cfg: dict, B: int, T: int, seed: int):
    """(x one-hot/scalar (B,Oin,T), idx (B,T) or None, latent c (B,C,T/hop), speaker ids g (B,1))."""
    rs = np.random.RandomState(seed)
    assert T % hop(cfg) == 0
    if cfg["scalar_input"]:
        idx = None
        x = torch.tensor(rs.uniform(-1, 1, size=(B, 1, T)), dtype=torch.float32)
    else:
        idx = torch.tensor(rs.randint(0, cfg["out_channels"], size=(B, T)), dtype=torch.long)
        x = torch.nn.functional.one_hot(idx, cfg["out_channels"]).float().transpose(1, 2).contiguous()
    c = torch.tensor(rs.normal(size=(B, cfg["cin_channels"], T // hop(cfg))), dtype=torch.float32)
    g = torch.tensor(rs.randint(0, cfg["n_speakers"], size=(B, 1)), dtype=torch.long)
    return x, idx, c, g
