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


class RecordingF:
    """Stands in for `torch.nn.functional` inside the reference's wavenet module: remembers the logits of the last softmax
    call so that the patched OneHotCategorical below can draw from (logits, supplied uniform) with the oracle's inverse-CDF
    sampler -- the RNG contract of SURVEY.md 8c (L3).  Harness-side only; the reference files are untouched."""

    def __init__(self, real):
        self._real, self.last_logits = real, None

    def __getattr__(self, k):
        return getattr(self._real, k)

    def softmax(self, x, dim=None, **kw):
        self.last_logits = x.detach()
        return self._real.softmax(x, dim=dim, **kw)


def sampled_free_run(m, x, c, g, Tn, u):
    """Reference incremental_forward(softmax=True, quantize=True) with OneHotCategorical(probs).sample() replaced by the
    oracle's inverse-CDF draw on the supplied uniforms u (T,B).  Returns the sampled classes (B,T)."""
    import wavenet_vocoder.wavenet as ref_wn_mod
    from oracle import sampling
    rec = RecordingF(ref_wn_mod.F)
    step = [0]
    O = m.out_channels

    class FakeOneHotCategorical:
        def __init__(self, probs):
            self.B = probs.shape[0]

        def sample(self):
            t = step[0]
            step[0] += 1
            lg = rec.last_logits.numpy()
            k = [sampling.categorical_from_uniform(lg[b], float(u[t, b])) for b in range(self.B)]
            return torch.nn.functional.one_hot(torch.tensor(k), O).float()
    real_f, real_d = ref_wn_mod.F, torch.distributions.OneHotCategorical
    ref_wn_mod.F, torch.distributions.OneHotCategorical = rec, FakeOneHotCategorical
    try:
        y = m.incremental_forward(initial_input=x[:, :, :1], c=c, g=g, T=Tn, test_inputs=x[:, :, :1], softmax=True, quantize=True)
    finally:
        ref_wn_mod.F, torch.distributions.OneHotCategorical = real_f, real_d
    assert step[0] == Tn
    return y.argmax(1)


def wavenet_case(name, cfg_name, seed, B, Tn, in_seed, stride=1, incremental=False, free=True, sampled_seed=None):
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
            if free:
                # free-running, no sampling: softmax probabilities are fed back (softmax=True, quantize=False)
                Tfree = 3 * T.hop(cfg)
                yf = m.incremental_forward(initial_input=x[:, :, :1], c=c[:, :, :3], g=g, T=Tfree,
                                           test_inputs=x[:, :, :1], softmax=True, quantize=False)
                arrays["free_probs"] = yf
                arrays["Tfree"] = Tfree
        if sampled_seed is not None:
            # L3: free-running categorical synthesis on a supplied uniform stream (T,B)
            u = torch.rand(Tn, B, generator=torch.Generator().manual_seed(sampled_seed))
            arrays["sampled_seed"] = sampled_seed
            arrays["sampled"] = sampled_free_run(m, x, c, g, Tn, u.numpy()).to(torch.int16)
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



class NormalFeeder:
    """Make torch.distributions.Normal.sample() consume a known N(0,1) stream: value = loc + scale * eps (what
    torch.normal(mean, std) computes: eps*std, then +mean)."""

    def __init__(self, stream):
        self.stream, self.pos = stream, 0

    def __enter__(self):
        self.orig = torch.distributions.Normal.sample
        feeder = self

        def fake(dist, sample_shape=torch.Size()):
            shape = dist._extended_shape(sample_shape)
            n = int(np.prod(shape)) if len(shape) else 1
            eps = torch.tensor(feeder.stream[feeder.pos:feeder.pos + n], dtype=dist.loc.dtype).reshape(shape)
            feeder.pos += n
            return eps * dist.scale.expand(shape) + dist.loc.expand(shape)
        torch.distributions.Normal.sample = fake
        return self

    def __exit__(self, *a):
        torch.distributions.Normal.sample = self.orig


def sampler_gauss_case():
    """mixture.py:221-270 on supplied draws: nmix uniforms (mixture indicator, Gumbel-max) + one N(0,1) value per row; also the
    single-Gaussian layouts C = 2 and C = 3."""
    rs = np.random.RandomState(12)
    arrays = {}
    for tag, C in (("mix", 15), ("c2", 2), ("c3", 3)):
        nmix = 1 if C == 2 else C // 3
        N = 64
        y = rs.normal(size=(N, C, 1))
        y[:, -nmix:, 0] = y[:, -nmix:, 0] * 0.7 - 2.0           # log-scales
        if C >= 3:
            y[:, nmix:2 * nmix, 0] *= 0.4                        # means
        else:
            y[:, 0, 0] *= 0.4
        y = torch.tensor(y, dtype=torch.float32)
        u = np.concatenate([rs.uniform(size=(N, nmix)), rs.normal(size=(N, 1))], axis=1).astype(np.float32)
        with UniformFeeder(u[:, :nmix].reshape(-1).astype(np.float64)), NormalFeeder(u[:, nmix].astype(np.float64)):
            x = ref_mixture.sample_from_mix_gaussian(y)
        arrays.update({f"y_{tag}": y[:, :, 0], f"u_{tag}": u, f"x_{tag}": x.reshape(-1)})
    save("sampler_gauss", **arrays)


def scalar_ar_case(name, cfg_name, seed, B, Tn, in_seed, u_seed):
    """Free-running synthesis of a scalar-input model (MoL or mixture of Gaussians) through the reference's
    incremental_forward with its random draws supplied: per step the reference draws B*nmix uniforms (indicator), then B
    values (logistic uniform / normal eps) -- mixture.py:138,151 / :251,266."""
    cfg = T.CONFIGS[cfg_name]
    m = build_ref(cfg_name, seed)
    _, _, c, g = T.synth_inputs(cfg, B, Tn, in_seed)
    nmix = cfg["out_channels"] // 3
    rs = np.random.RandomState(u_seed)
    gauss = cfg["output_distribution"] == "Normal"
    u = rs.uniform(size=(Tn, B, nmix + 1)).astype(np.float32)
    if gauss:
        u[:, :, nmix] = rs.normal(size=(Tn, B)).astype(np.float32)
    init = torch.zeros(B, 1, 1)
    with torch.no_grad():
        if gauss:
            with UniformFeeder(u[:, :, :nmix].reshape(-1).astype(np.float64)), NormalFeeder(u[:, :, nmix].reshape(-1).astype(np.float64)):
                y = m.incremental_forward(initial_input=init, c=c, g=g, T=Tn, test_inputs=init, log_scale_min=-7.0)
        else:
            stream = np.concatenate([np.concatenate([u[t, :, :nmix].reshape(-1), u[t, :, nmix]]) for t in range(Tn)]).astype(np.float64)
            with UniformFeeder(stream):
                y = m.incremental_forward(initial_input=init, c=c, g=g, T=Tn, test_inputs=init, log_scale_min=-7.0)
    save(name, cfg=cfg_name, seed=seed, in_seed=in_seed, B=B, T=Tn, u=u, samples=y[:, 0, :])


def vq_ema_case(name, kind, K, D, B, Tn, seed, steps=3):
    """The EMA variants only run where `.cuda()` works (vector_quantization.py:183,277 test the function object, not its
    result).  Harness-side shim: Tensor.cuda -> identity for the duration of the call, so the reference's own forward runs
    on the CPU.  `steps` training-mode forwards (each rewrites the codebook before its gather), then one eval forward."""
    rs = np.random.RandomState(seed)
    torch.manual_seed(seed)
    m = getattr(ref_vq, kind)(K, D).train()
    for p in m.parameters():
        p.data = torch.tensor(rs.normal(size=tuple(p.shape)) * 0.5, dtype=torch.float32)
    arrays = dict(kind=kind, K=K, D=D, steps=steps)
    for n, p in m.named_parameters():
        arrays["param_" + n.replace(".", "__")] = p.detach().clone()
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for s in range(steps + 1):
            if s == steps:
                m.eval()
            x = torch.tensor(rs.normal(size=(B, D, Tn)) * 0.5, dtype=torch.float32)
            with torch.no_grad():
                quant, loss, perp = m(x)
            arrays.update({f"x{s}": x, f"quant{s}": quant.contiguous(), f"vq_loss{s}": loss.item(), f"perp{s}": perp.item()})
            for n, p in list(m.named_parameters()) + list(m.named_buffers()):
                arrays[f"after{s}_" + n.replace(".", "__")] = p.detach().clone()
    finally:
        torch.Tensor.cuda = real_cuda
    save(name, **arrays)


def vqvae_vqwae_case():
    """BASELINE configs[1] at its real shape: VQVAE(WaveNet(**vqwae), c_in=39, hid=64, K=256 (default), encoder_hid=256);
    100 MFCC frames -> 25 latents -> T = 16000, B = 2.  Codes are read off the reference's own argmin call."""
    cfg = T.CONFIGS["vqwae"]
    torch.manual_seed(0)
    m = ref_vqvae.VQVAE(c_in=39, hid=cfg["cin_channels"], K=256, wavenet=RefWaveNet(**cfg), encoder_hid=256).eval()
    m.load_state_dict(T.synth_state_dict(m, 5))
    rs = np.random.RandomState(16)
    B, frames, Tn, stride = 2, 100, 16000, 64
    mfcc = torch.tensor(rs.normal(size=(B, 39, frames)), dtype=torch.float32)
    idx = torch.tensor(rs.randint(0, cfg["out_channels"], size=(B, Tn)))
    x = torch.nn.functional.one_hot(idx, cfg["out_channels"]).float().transpose(1, 2).contiguous()
    g = torch.tensor(rs.randint(0, cfg["n_speakers"], size=(B, 1)))
    seen = []
    real_argmin = torch.argmin

    def rec_argmin(*a, **k):
        r = real_argmin(*a, **k)
        seen.append(r.clone())
        return r
    torch.argmin = rec_argmin
    try:
        with torch.no_grad():
            y, vq_loss, perp = m(x, mfcc, g)
    finally:
        torch.argmin = real_argmin
    with torch.no_grad():
        lat = m.encoder(mfcc)
        quant = m.encode(mfcc)
    assert len(seen) == 1
    save("vqvae_vqwae", seed=5, B=B, T=Tn, stride=stride, mfcc=mfcc, idx=idx.to(torch.int16), g=g, logits=y[:, :, ::stride],
         logits_sum=y.double().sum().item(), vq_loss=vq_loss.item(), perp=perp.item(), latents=lat, quant=quant,
         codes=seen[0].view(B, -1))


CASES = {
    "wavenet_tiny": lambda: wavenet_case("wavenet_tiny", "tiny", seed=1, B=2, Tn=320, in_seed=2, incremental=True),
    "wavenet_tiny_k2": lambda: wavenet_case("wavenet_tiny_k2", "tiny_k2", seed=3, B=3, Tn=160, in_seed=4, incremental=True),
    "wavenet_tiny_mol": lambda: wavenet_case("wavenet_tiny_mol", "tiny_mol", seed=7, B=2, Tn=160, in_seed=8),
    "wavenet_vqwae": lambda: wavenet_case("wavenet_vqwae", "vqwae", seed=1, B=1, Tn=1280, in_seed=2, stride=16),
    "wavenet_inwae": lambda: wavenet_case("wavenet_inwae", "inwae", seed=2, B=1, Tn=640, in_seed=3, stride=8),
    # the benchmarked shapes with the d = 512 history (1025 rows) wrapping and the t-1024 taps live, two utterances
    "wavenet_vqwae_b2": lambda: wavenet_case("wavenet_vqwae_b2", "vqwae", seed=1, B=2, Tn=2560, in_seed=12, stride=8,
                                             incremental=True, free=False, sampled_seed=31),
    "wavenet_inwae_b2": lambda: wavenet_case("wavenet_inwae_b2", "inwae", seed=2, B=2, Tn=2560, in_seed=13, stride=8,
                                             incremental=True, free=False),
    "sampler_mol": sampler_case,
    "sampler_gauss": sampler_gauss_case,
    "wavenet_tiny_mol_ar": lambda: scalar_ar_case("wavenet_tiny_mol_ar", "tiny_mol", seed=7, B=3, Tn=96, in_seed=8, u_seed=41),
    "wavenet_tiny_gauss_ar": lambda: scalar_ar_case("wavenet_tiny_gauss_ar", "tiny_gauss", seed=9, B=3, Tn=96, in_seed=10, u_seed=42),
    "vq_plain_default": lambda: vq_case("vq_plain_default", "VectorQuantize", 256, 64, 4, 25, 21, "default"),
    "vq_plain_trained": lambda: vq_case("vq_plain_trained", "VectorQuantize", 256, 64, 16, 25, 22, "trained"),
    "vq_sliced_default": lambda: vq_case("vq_sliced_default", "SlicedVectorQuantize", 256, 64, 4, 25, 23, "default"),
    "vq_sliced_trained": lambda: vq_case("vq_sliced_trained", "SlicedVectorQuantize", 256, 64, 16, 25, 24, "trained"),
    "vq_ema_plain": lambda: vq_ema_case("vq_ema_plain", "VectorQuantizeEMA", 64, 32, 4, 50, 25),
    "vq_ema_sliced": lambda: vq_ema_case("vq_ema_sliced", "SlicedVectorQuantizeEMA", 64, 32, 4, 50, 26),
    "vqvae_tiny": vqvae_case,
    "vqvae_vqwae": vqvae_vqwae_case,
}

if __name__ == "__main__":
    for case in (sys.argv[1:] or list(CASES)):
        CASES[case]()
