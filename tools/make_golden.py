"""Generate tests/golden/*.npz from the REAL reference modules (run in the build container only:
needs /root/reference; the GPU box never runs this).

    python tools/make_golden.py

Every fixture holds the inputs' recipe (config name, seeds, sizes) plus the reference's outputs, so a test can
rebuild the identical model with wavenet_autoencoders_b200.testing.synth_state_dict and compare.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WAE_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")

sys.path.insert(0, ROOT)
from wavenet_autoencoders_b200 import testing as T  # noqa: E402  (pure torch/numpy helpers; no CUDA needed)

sys.path.insert(0, REF)  # the reference's own packages: wavenet_vocoder, vector_quantization, vqvae_model
import vector_quantization as ref_vq  # noqa: E402
import vqvae_model as ref_vqvae  # noqa: E402
from wavenet_vocoder import WaveNet as RefWaveNet  # noqa: E402
from wavenet_vocoder import mixture as ref_mixture  # noqa: E402

assert RefWaveNet.__module__.startswith("wavenet_vocoder") and REF in sys.modules["wavenet_vocoder"].__file__


def build_ref(cfg_name, seed):
    torch.manual_seed(0)
    m = RefWaveNet(**T.CONFIGS[cfg_name]).eval()
    m.load_state_dict(T.synth_state_dict(m, seed))
    return m


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def wavenet_case(name, cfg_name, seed, B, Tn, in_seed, stride=1, incremental=False):
    cfg = T.CONFIGS[cfg_name]
    m = build_ref(cfg_name, seed)
    x, idx, c, g = T.synth_inputs(cfg, B, Tn, in_seed)
    with torch.no_grad():
        y = m(x, c, g)
        c_up = m.upsample_net(c)
        arrays = dict(cfg=cfg_name, seed=seed, in_seed=in_seed, B=B, T=Tn, stride=stride,
                      logits=y[:, :, ::stride], logits_abs_sum=y.abs().sum().item(), logits_sum=y.double().sum().item(),
                      c_up=c_up[:, :, ::stride])
        if incremental:
            # teacher-forced incremental pass == forward (the reference's own consistency relation)
            yi = m.incremental_forward(initial_input=x[:, :, :1], c=c, g=g, T=Tn, test_inputs=x, softmax=False, quantize=False)
            arrays["inc_logits"] = yi[:, :, ::stride]
            # free-running, no sampling: softmax probabilities are fed back (softmax=True, quantize=False)
            Tfree = 3 * T.hop(cfg)
            yf = m.incremental_forward(initial_input=x[:, :, :1], c=c[:, :, :3], g=g, T=Tfree,
                                       test_inputs=x[:, :, :1], softmax=True, quantize=False)
            arrays["free_probs"] = yf
            arrays["Tfree"] = Tfree
    save(name, **arrays)


class UniformFeeder:
    """Make Tensor.uniform_(a, b) consume a known stream u in [0,1): value = a + (b-a)*u (what ATen computes)."""

    def __init__(self, stream):
        self.stream, self.pos = stream, 0

    def __enter__(self):
        self.orig = torch.Tensor.uniform_
        feeder = self

        def fake(t, a=0.0, b=1.0):
            n = t.numel()
            u = feeder.stream[feeder.pos:feeder.pos + n].reshape(t.shape)
            feeder.pos += n
            return t.copy_(torch.tensor(a + (b - a) * u, dtype=t.dtype))
        torch.Tensor.uniform_ = fake
        return self

    def __exit__(self, *a):
        torch.Tensor.uniform_ = self.orig


def sampler_case():
    rs = np.random.RandomState(11)
    nmix, N = 5, 64
    y = torch.tensor(rs.normal(size=(N, 3 * nmix, 1)) * np.array([1.0] * nmix + [0.3] * nmix + [1.0] * nmix)[None, :, None]
                     - np.array([0.0] * (2 * nmix) + [3.0] * nmix)[None, :, None], dtype=torch.float32)
    u = rs.uniform(size=(N, nmix + 1)).astype(np.float32)
    # the reference draws nmix uniforms for all rows first, then one per row
    stream = np.concatenate([u[:, :nmix].reshape(-1), u[:, nmix]]).astype(np.float64)
    with UniformFeeder(stream):
        x = ref_mixture.sample_from_discretized_mix_logistic(y)
    save("sampler_mol", y=y[:, :, 0], u=u, x=x.reshape(-1))


def vq_case(name, kind, K, D, B, Tn, seed, codebook):
    rs = np.random.RandomState(seed)
    x = torch.tensor(rs.normal(size=(B, D, Tn)) * 0.5, dtype=torch.float32)
    torch.manual_seed(seed)
    m = getattr(ref_vq, kind)(K, D)
    if codebook == "trained":
        for p in m.parameters():
            p.data = torch.tensor(rs.normal(size=tuple(p.shape)) * 0.5, dtype=torch.float32)
    with torch.no_grad():
        quant, loss, perp = m(x)
    arrays = dict(kind=kind, K=K, D=D, x=x, quant=quant.contiguous(), vq_loss=loss.item(), perp=perp.item())
    for n, p in m.named_parameters():
        arrays["param_" + n.replace(".", "__")] = p.detach()
    save(name, **arrays)


def vqvae_case():
    cfg = T.CONFIGS["tiny"]
    torch.manual_seed(0)
    wn = RefWaveNet(**cfg)
    m = ref_vqvae.VQVAE(c_in=39, hid=cfg["cin_channels"], K=32, wavenet=wn, encoder_hid=48).eval()
    m.load_state_dict(T.synth_state_dict(m, 5))
    rs = np.random.RandomState(6)
    frames = 20                     # -> 5 latents -> T = 80
    mfcc = torch.tensor(rs.normal(size=(2, 39, frames)), dtype=torch.float32)
    Tn = 5 * T.hop(cfg)
    idx = torch.tensor(rs.randint(0, cfg["out_channels"], size=(2, Tn)))
    x = torch.nn.functional.one_hot(idx, cfg["out_channels"]).float().transpose(1, 2).contiguous()
    g = torch.tensor(rs.randint(0, cfg["n_speakers"], size=(2, 1)))
    with torch.no_grad():
        y, vq_loss, perp = m(x, mfcc, g)
        lat = m.encoder(mfcc)
        quant = m.encode(mfcc)
    save("vqvae_tiny", mfcc=mfcc, idx=idx, g=g, logits=y, vq_loss=vq_loss.item(), perp=perp.item(), latents=lat, quant=quant)


if __name__ == "__main__":
    wavenet_case("wavenet_tiny", "tiny", seed=1, B=2, Tn=320, in_seed=2, incremental=True)
    wavenet_case("wavenet_tiny_k2", "tiny_k2", seed=3, B=3, Tn=160, in_seed=4, incremental=True)
    wavenet_case("wavenet_tiny_mol", "tiny_mol", seed=7, B=2, Tn=160, in_seed=8)
    wavenet_case("wavenet_vqwae", "vqwae", seed=1, B=1, Tn=1280, in_seed=2, stride=16)
    wavenet_case("wavenet_inwae", "inwae", seed=2, B=1, Tn=640, in_seed=3, stride=8)
    sampler_case()
    vq_case("vq_plain_default", "VectorQuantize", 256, 64, 4, 25, 21, "default")
    vq_case("vq_plain_trained", "VectorQuantize", 256, 64, 16, 25, 22, "trained")
    vq_case("vq_sliced_default", "SlicedVectorQuantize", 256, 64, 4, 25, 23, "default")
    vq_case("vq_sliced_trained", "SlicedVectorQuantize", 256, 64, 16, 25, 24, "trained")
    vqvae_case()
