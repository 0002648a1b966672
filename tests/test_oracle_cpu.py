"""Pins the oracle (oracle/) against golden vectors produced by the REAL reference (tools/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import build_model, load_golden, rel_err
from oracle import sampling, torch_port, vq_oracle
from oracle import wavenet_oracle as wo
from wavenet_autoencoders_b200 import testing as T

WAVENET_CASES = ["wavenet_tiny", "wavenet_tiny_k2", "wavenet_tiny_mol", "wavenet_vqwae", "wavenet_inwae", "wavenet_vqwae_b2",
                 "wavenet_inwae_b2"]


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


@pytest.mark.parametrize("case", ["wavenet_tiny", "wavenet_tiny_k2", "wavenet_vqwae_b2", "wavenet_inwae_b2"])
def test_oracle_incremental_matches_reference(case):
    """*_b2: the benchmarked 20-layer shapes, two utterances, T = 2560 -- the d = 512 history (1025 rows) wraps and the
    taps at t-1024 are live (conv.py:34-46)."""
    g, cfg, p, x, c, spk = _setup(case)
    Tn = int(g["T"])
    s = int(g["stride"])
    forced = np.ascontiguousarray(x.transpose(0, 2, 1))
    y = wo.incremental_forward(p, Tn, c=c, g=spk, initial_input=forced[:, 0], test_inputs=forced)   # (B,T,O)
    assert rel_err(y.transpose(0, 2, 1)[:, :, ::s], g["inc_logits"]) < 2e-5
    if "sampled" in g:
        # L3: free-running categorical synthesis on the golden's uniform stream; the reference was driven by the same
        # inverse-CDF sampler, so the classes must agree for the whole run (first divergence-free window = T)
        u = torch.rand(Tn, int(g["B"]), generator=torch.Generator().manual_seed(int(g["sampled_seed"]))).numpy()
        O = cfg["out_channels"]
        picks = []

        def sampler(t, logits):
            k = np.array([sampling.categorical_from_uniform(logits[b], float(u[t, b])) for b in range(logits.shape[0])])
            picks.append(k)
            return np.eye(O, dtype=np.float32)[k]
        wo.incremental_forward(p, Tn, c=c, g=spk, initial_input=forced[:, 0], test_inputs=forced[:, :1], sampler=sampler)
        got = np.stack(picks, 1)
        first_div = int(np.argmax((got != g["sampled"]).any(0))) if (got != g["sampled"]).any() else Tn
        assert first_div == Tn, f"first divergence at step {first_div}"
    if "free_probs" not in g:
        return
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


@pytest.mark.parametrize("case,hid,K,seed", [("vqvae_tiny", 48, 32, 5), ("vqvae_vqwae", 256, 256, None)])
def test_encoder_oracle_matches_reference_latents(case, hid, K, seed):
    """oracle/encoder_oracle.py (numpy restatement of vqvae_model.py:9-51) against the latents the REAL reference's encoder
    produced (goldens), and the reference's codes from the C oracle's search on those latents (row f3's checker)."""
    from oracle import encoder_oracle as eo
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    g = load_golden(case)
    cfg = T.CONFIGS["tiny" if case == "vqvae_tiny" else "vqwae"]
    torch.manual_seed(0)
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    m = VQVAE(c_in=39, hid=cfg["cin_channels"], K=K, wavenet=WaveNet(**cfg), encoder_hid=hid).eval()
    m.load_state_dict(T.synth_state_dict(m, int(g["seed"]) if seed is None else seed))
    sd = {k: v.numpy() for k, v in m.state_dict().items() if k.startswith("encoder.")}
    lat = eo.encoder_forward(sd, g["mfcc"])
    assert lat.shape == g["latents"].shape
    assert rel_err(lat, g["latents"]) < 2e-6
    if "codes" in g:
        _, _, _, idx = vq_oracle.vq_forward(lat, m.vq.embedding.weight.detach().numpy())
        np.testing.assert_array_equal(idx, g["codes"])


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


# ------------------------------------------------------------------ round-2 pins: Gaussian sampler, scalar-input AR, EMA VQ
def test_gauss_sampler_matches_reference():
    """oracle/sampling.gauss_from_draws against mixture.py:221-270 run with its uniform_/Normal.sample draws supplied."""
    g = load_golden("sampler_gauss")
    for tag in ("mix", "c2", "c3"):
        got = np.array([sampling.gauss_from_draws(y, u) for y, u in zip(g[f"y_{tag}"], g[f"u_{tag}"])], np.float32)
        np.testing.assert_allclose(got, g[f"x_{tag}"], rtol=2e-6, atol=2e-7, err_msg=tag)
        assert np.abs(g[f"x_{tag}"]).max() <= 1.0 and (np.abs(g[f"x_{tag}"]) < 1.0).any()


@pytest.mark.parametrize("case,fn", [("wavenet_tiny_mol_ar", "mol_from_uniform"), ("wavenet_tiny_gauss_ar", "gauss_from_draws")])
def test_oracle_scalar_ar_matches_reference(case, fn):
    """Free-running synthesis of scalar-input models: the oracle loop + oracle sampler against the reference's
    incremental_forward fed the same draws (wavenet.py:324-331)."""
    g = load_golden(case)
    cfg = T.CONFIGS[str(g["cfg"])]
    m = build_model(str(g["cfg"]), int(g["seed"]))
    p = wo.extract_params({k: v.numpy() for k, v in m.state_dict().items()}, cfg["layers"], cfg["stacks"])
    B, Tn = int(g["B"]), int(g["T"])
    _, _, c, spk = T.synth_inputs(cfg, B, Tn, int(g["in_seed"]))
    u = g["u"]
    draw = getattr(sampling, fn)
    y = wo.incremental_forward(p, Tn, c=c.numpy(), g=spk.numpy(), initial_input=np.zeros((B, 1), np.float32),
                               test_inputs=np.zeros((B, 1, 1), np.float32),
                               sampler=lambda t, lg: np.array([[draw(lg[b], u[t, b])] for b in range(B)], np.float32))
    np.testing.assert_allclose(y[:, :, 0], g["samples"], atol=5e-5)


@pytest.mark.parametrize("case", ["vq_ema_plain", "vq_ema_sliced"])
def test_vq_ema_oracle_matches_reference(case):
    """vq_oracle.ema_update (+ search) against the reference's EMA classes run in training mode for several steps
    (vector_quantization.py:156-235, :257-306; run on the CPU through a harness-side `.cuda()` shim, tools/make_golden.py):
    cluster sizes, per-code sums, the codebook overwritten BEFORE the gather, commitment-only loss."""
    g = load_golden(case)
    K, D, steps = int(g["K"]), int(g["D"]), int(g["steps"])
    sliced = str(g["kind"]).startswith("Sliced")
    names = ["1", "2"] if sliced else [""]
    sd = D // len(names)
    cbs = [g[f"param_embedding{n}__weight"].copy() for n in names]
    sizes = [np.zeros(K, np.float32) for _ in names]
    ws = [np.zeros((K, sd), np.float32) for _ in names]
    for s in range(steps + 1):
        x = g[f"x{s}"]
        B, _, Tn = x.shape
        flat = x.transpose(0, 2, 1).reshape(-1, D)
        q = np.empty((B * Tn, D), np.float32)
        perp = 0.0
        for i, n in enumerate(names):
            idx, _, _ = vq_oracle.search(x, cbs[i], i * sd, sd)
            if s < steps:                                    # training-mode forward
                sizes[i], ws[i], cbs[i] = vq_oracle.ema_update(flat[:, i * sd:(i + 1) * sd], idx, K, sizes[i], ws[i], 0.99)
            q[:, i * sd:(i + 1) * sd] = cbs[i][idx]
            perp += float(vq_oracle.perplexity(idx, K))
            np.testing.assert_allclose(cbs[i], g[f"after{s}_embedding{n}__weight"], rtol=2e-5, atol=1e-7)
            np.testing.assert_allclose(sizes[i], g[f"after{s}_ema_cluster_size{n}"], rtol=2e-5, atol=1e-9)
            np.testing.assert_allclose(ws[i], g[f"after{s}_ema_w{n}"], rtol=2e-5, atol=1e-7)
        quant = (flat + (q - flat)).reshape(B, Tn, D).transpose(0, 2, 1)
        np.testing.assert_allclose(quant, g[f"quant{s}"], rtol=2e-5, atol=1e-6)
        loss = 0.25 * np.mean((q.astype(np.float64) - flat) ** 2)
        assert abs(loss - float(g[f"vq_loss{s}"])) <= 1e-5 * float(g[f"vq_loss{s}"])
        assert abs(perp - float(g[f"perp{s}"])) <= 1e-5 * float(g[f"perp{s}"])
