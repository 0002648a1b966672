"""Pins the oracle (oracle/) against golden vectors produced by the REAL reference (tools/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import build_model, load_golden, rel_err
from oracle import sampling, torch_port, vq_oracle
from oracle import wavenet_oracle as wo
from wavenet_autoencoders_b200 import testing as T

WAVENET_CASES = ["wavenet_tiny", "wavenet_tiny_k2", "wavenet_tiny_mol", "wavenet_vqwae", "wavenet_inwae"]


def _setup(case):
    g = load_golden(case)
    cfg_name = str(g["cfg"])
    cfg = T.CONFIGS[cfg_name]
    m = build_model(cfg_name, int(g["seed"]))
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    p = wo.extract_params(sd, cfg["layers"], cfg["stacks"])
    x, idx, c, spk = T.synth_inputs(cfg, int(g["B"]), int(g["T"]), int(g["in_seed"]))
    return g, cfg, p, x.numpy(), c.numpy(), spk.numpy()


@pytest.mark.parametrize("case", WAVENET_CASES)
def test_numpy_oracle_forward_matches_reference(case):
    g, cfg, p, x, c, spk = _setup(case)
    s = int(g["stride"])
    c_up = wo.upsample_conditioning(p, c)
    assert rel_err(c_up[:, :, ::s], g["c_up"]) < 2e-6
    y = wo.forward(p, x, c, spk)
    assert y.shape == (int(g["B"]), cfg["out_channels"], int(g["T"]))
    assert rel_err(y[:, :, ::s], g["logits"]) < 2e-5
    assert abs(float(np.abs(y).sum()) - float(g["logits_abs_sum"])) < 1e-4 * float(g["logits_abs_sum"])


@pytest.mark.parametrize("case", WAVENET_CASES)
def test_torch_port_forward_matches_reference(case):
    g, cfg, p, x, c, spk = _setup(case)
    tp = torch_port.params_from_numpy(p)
    with torch.no_grad():
        c_up = torch_port.upsample(tp, torch.tensor(c))
        gv = torch.tensor(wo.speaker_vectors(p, spk))
        y = torch_port.stack_forward(tp, torch.tensor(x), c_up, gv).numpy()
    s = int(g["stride"])
    assert rel_err(y[:, :, ::s], g["logits"]) < 2e-5


@pytest.mark.parametrize("case", ["wavenet_tiny", "wavenet_tiny_k2"])
def test_oracle_incremental_matches_reference(case):
    g, cfg, p, x, c, spk = _setup(case)
    Tn = int(g["T"])
    forced = np.ascontiguousarray(x.transpose(0, 2, 1))
    y = wo.incremental_forward(p, Tn, c=c, g=spk, initial_input=forced[:, 0], test_inputs=forced)   # (B,T,O)
    assert rel_err(y.transpose(0, 2, 1), g["inc_logits"]) < 2e-5
    # free-running with softmax feedback (softmax=True, quantize=False): probabilities fed back as the next input
    Tf = int(g["Tfree"])
    probs = wo.incremental_forward(
        p, Tf, c=c[:, :, :3], g=spk, initial_input=forced[:, 0], test_inputs=forced[:, :1],
        sampler=lambda t, lg: np.stack([sampling.softmax_probs(r) for r in lg]))
    assert rel_err(probs.transpose(0, 2, 1), g["free_probs"]) < 5e-5


def test_torch_port_ar_matches_reference():
    g, cfg, p, x, c, spk = _setup("wavenet_tiny")
    tp = torch_port.params_from_numpy(p)
    forced = torch.tensor(x).transpose(1, 2).contiguous()
    with torch.no_grad():
        c_btc = torch_port.upsample(tp, torch.tensor(c)).transpose(1, 2).contiguous()
        gv = torch.tensor(wo.speaker_vectors(p, spk))
        y = torch_port.ar_generate(tp, 64, c_btc, gv, forced[:, 0], test_inputs=forced).numpy()
    assert rel_err(y.transpose(0, 2, 1), g["inc_logits"][:, :, :64]) < 2e-5


def test_mol_sampler_matches_reference():
    g = load_golden("sampler_mol")
    got = np.array([sampling.mol_from_uniform(y, u) for y, u in zip(g["y"], g["u"])], np.float32)
    np.testing.assert_allclose(got, g["x"], rtol=2e-5, atol=2e-6)


def test_categorical_sampler_is_inverse_cdf():
    rs = np.random.RandomState(0)
    for _ in range(50):
        lg = rs.normal(size=256).astype(np.float32) * 3
        u = float(rs.uniform())
        k = sampling.categorical_from_uniform(lg, u)
        p = np.exp(lg.astype(np.float64) - lg.max())
        cdf = np.cumsum(p / p.sum())
        k64 = int(np.searchsorted(cdf, u, side="right"))
        assert abs(k - min(k64, 255)) <= 1 and (k == min(k64, 255) or abs(cdf[min(k, k64)] - u) < 1e-5)
    lg = np.zeros(32, np.float32)
    assert sampling.categorical_from_uniform(lg, 0.0) == 0
    assert sampling.categorical_from_uniform(lg, 0.999999) == 31
    assert abs(sampling.softmax_probs(lg).sum() - 1) < 1e-6


def _codes_from_quant(quant_bdt, codebook):
    """The reference does not return indices: recover them as the codebook row nearest to its output."""
    q = quant_bdt.transpose(0, 2, 1).reshape(-1, codebook.shape[1]).astype(np.float64)
    d = ((q[:, None, :] - codebook[None].astype(np.float64)) ** 2).sum(-1)
    return d.argmin(1)


@pytest.mark.parametrize("case", ["vq_plain_default", "vq_plain_trained"])
def test_vq_oracle_matches_reference(case):
    g = load_golden(case)
    cb = g["param_embedding__weight"]
    quant, loss, perp, idx = vq_oracle.vq_forward(g["x"], cb)
    ref_idx = _codes_from_quant(g["quant"], cb)
    _, best, second = vq_oracle.search(g["x"], cb)
    mism = np.flatnonzero(idx.reshape(-1) != ref_idx)
    # any disagreement must be a rounding-level near-tie of the reference's own fp32 distances
    scale = np.abs(best) + 1e-30
    assert np.all((second - best)[mism] <= 4 * np.spacing(scale[mism].astype(np.float32)) + 1e-12), (case, len(mism))
    if case.endswith("trained"):
        assert len(mism) == 0
        np.testing.assert_array_equal(quant, g["quant"])           # bit-exact straight-through values
    assert abs(float(loss) - float(g["vq_loss"])) <= 1e-5 * abs(float(g["vq_loss"])) + 1e-9
    if len(mism) == 0:
        assert abs(float(perp) - float(g["perp"])) <= 1e-5 * float(g["perp"])


@pytest.mark.parametrize("case", ["vq_sliced_default", "vq_sliced_trained"])
def test_sliced_vq_oracle_matches_reference(case):
    g = load_golden(case)
    cb1, cb2 = g["param_embedding1__weight"], g["param_embedding2__weight"]
    quant, loss, perp, idx = vq_oracle.sliced_vq_forward(g["x"], cb1, cb2)
    sd = cb1.shape[1]
    r1 = _codes_from_quant(g["quant"][:, :sd], cb1)
    r2 = _codes_from_quant(g["quant"][:, sd:], cb2)
    n_mism = int((idx[..., 0].reshape(-1) != r1).sum() + (idx[..., 1].reshape(-1) != r2).sum())
    if case.endswith("trained"):
        assert n_mism == 0
        np.testing.assert_array_equal(quant, g["quant"])
        assert abs(float(perp) - float(g["perp"])) <= 1e-5 * float(g["perp"])
    assert abs(float(loss) - float(g["vq_loss"])) <= 1e-5 * abs(float(g["vq_loss"])) + 1e-9


def test_torch_port_vq_matches_reference():
    g = load_golden("vq_plain_trained")
    q, loss, perp, idx = torch_port.vq_forward(torch.tensor(g["x"]), torch.tensor(g["param_embedding__weight"]))
    np.testing.assert_array_equal(q.numpy(), g["quant"])
    assert abs(loss.item() - float(g["vq_loss"])) < 1e-6


def test_vq_edge_cases():
    cb = np.array([[0.0, 0.0], [0.0, 0.0], [1.0, 1.0]], np.float32)          # duplicate codes -> first index wins
    x = np.zeros((1, 2, 3), np.float32)
    idx, _, _ = vq_oracle.search(x, cb)
    assert idx.tolist() == [0, 0, 0]
    idx, _, _ = vq_oracle.search(np.zeros((0, 2, 3), np.float32), cb)          # empty batch
    assert idx.shape == (0,)


def test_oracle_vqvae_composition():
    """encoder (torch, out of scope) -> oracle VQ -> oracle decoder reproduces the reference VQVAE.forward."""
    g = load_golden("vqvae_tiny")
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    cfg = T.CONFIGS["tiny"]
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=cfg["cin_channels"], K=32, wavenet=WaveNet(**cfg), encoder_hid=48).eval()
    m.load_state_dict(T.synth_state_dict(m, 5))
    with torch.no_grad():
        lat = m.encoder(torch.tensor(g["mfcc"])).numpy()
    assert rel_err(lat, g["latents"]) < 1e-6
    quant, loss, perp, _ = vq_oracle.vq_forward(g["latents"], m.vq.embedding.weight.detach().numpy())
    np.testing.assert_array_equal(quant, g["quant"])
    sd = {k[len("wavenet."):]: v.numpy() for k, v in m.state_dict().items() if k.startswith("wavenet.")}
    p = wo.extract_params(sd, cfg["layers"], cfg["stacks"])
    x = np.eye(cfg["out_channels"], dtype=np.float32)[g["idx"]].transpose(0, 2, 1)
    y = wo.forward(p, x, quant, g["g"])
    assert rel_err(y, g["logits"]) < 2e-5


def test_postprocess_oracle_known_values():
    """The restated mu-law inverse (nnmnkwii's published formula; parity unpinned) at values derivable by hand, its consistency
    with the forward companding of the same library, and the IIR against a direct recurrence."""
    from oracle import postprocess_oracle as po
    mu = 256
    assert po.inv_mulaw(0.0, mu) == 0.0
    assert abs(po.inv_mulaw(1.0, mu) - 1.0) < 1e-12 and abs(po.inv_mulaw(-1.0, mu) + 1.0) < 1e-12
    assert abs(po.inv_mulaw_quantize(mu, mu) - 1.0) < 1e-12 and abs(po.inv_mulaw_quantize(0, mu) + 1.0) < 1e-12
    assert po.inv_mulaw_quantize(128, mu) == 0.0                        # 2 * 128 / 256 - 1 = 0
    x = np.linspace(-1, 1, 1001)
    y = np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)            # mulaw(x, mu)
    np.testing.assert_allclose(po.inv_mulaw(y, mu), x, atol=1e-12)
    # silence maps to class 127 under mulaw_quantize = int((y + 1) / 2 * mu) (wavenet.py:288, audio.py:96) and decodes to ~0
    assert int((0.0 + 1) / 2 * mu * (1 - 1e-9)) == 127 and abs(po.inv_mulaw_quantize(127, mu)) < 2e-4
    rs = np.random.RandomState(0)
    v = rs.normal(size=(2, 300))
    w = np.zeros_like(v)
    for t in range(300):
        w[:, t] = v[:, t] + 0.85 * (w[:, t - 1] if t else 0.0)
    np.testing.assert_allclose(po.inv_preemphasis(v, 0.85), w, atol=1e-12)
