"""CPU tests of the host side: the C-ABI library loads and exports what include/wae_b200.h declares, the product
path fails loudly without a GPU (no fallback), state_dict layout matches the reference, weight packing is right."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, build_model, load_golden, rel_err
from oracle import wavenet_oracle as wo
from wavenet_autoencoders_b200 import _lib, packing, testing as T
from wavenet_autoencoders_b200 import vector_quantization as vqm


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "wae_b200.h")).read()
    declared = set(re.findall(r"\b(wae_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes parsed"
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wae_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.wae_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.lib()
    assert lib.wae_device_check(0) == -2 and b"no CPU fallback" in lib.wae_last_error()
    m = build_model("tiny", 1)
    x, _, c, g = T.synth_inputs(T.CONFIGS["tiny"], 1, 32, 0)
    with pytest.raises(_lib.WaeError):
        with torch.no_grad():
            m(x, c, g)
    with pytest.raises(_lib.WaeError):
        m.incremental_forward(initial_input=x[:, :, :1], c=c, g=g, T=32)
    with pytest.raises(_lib.WaeError):
        vqm.VectorQuantize(8, 4)(torch.zeros(1, 4, 3))
    with pytest.raises(_lib.WaeError):
        vqm.SlicedVectorQuantize(8, 4).encode_indices(torch.zeros(1, 4, 3))
    # the additive entry points added later obey the same rule: CUDA tensors or a loud error, never a CPU path
    from wavenet_autoencoders_b200.graphed import GraphedForward
    from wavenet_autoencoders_b200.losses import teacher_forced_nll
    from wavenet_autoencoders_b200.postprocess import waveform_from_synthesis
    with pytest.raises(_lib.WaeError):
        waveform_from_synthesis(torch.zeros(1, 8, dtype=torch.long))
    with pytest.raises(_lib.WaeError):
        teacher_forced_nll(torch.zeros(1, 4, 8), torch.zeros(1, 8, dtype=torch.long))
    with pytest.raises(_lib.WaeError):
        GraphedForward(m, torch.zeros(1, 32, dtype=torch.long), torch.zeros(1, 39, 4), torch.zeros(1, 1, dtype=torch.long))
    # ... and no GPU is touched before the arguments are validated
    with pytest.raises(ValueError):
        waveform_from_synthesis(torch.zeros(1, 8, dtype=torch.long), input_type="alaw")
    assert lib.wae_synth_postprocess(None, 0, 1, 8, 256, 0.0, 0.0, None, None) != 0           # rejected (no sm_100 device / null)
    assert lib.wae_vq_set_variant(2) != 0 and b"wae_vq_set_variant" in lib.wae_last_error()
    assert lib.wae_vq_set_variant(0) == 0


def test_state_dict_layout_matches_reference():
    """Key names/shapes of SURVEY.md 3.4 (302 keys, 7,555,218 parameters at hps/vqwae.json)."""
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet, receptive_field_size
    wn = WaveNet(**T.VQWAE)
    m = VQVAE(c_in=39, hid=64, wavenet=wn, encoder_hid=256)
    sd = m.state_dict()
    assert len(sd) == 302
    assert sum(p.numel() for p in m.parameters()) == 7555218
    assert sum(p.numel() for p in wn.parameters()) == 5982546
    for k, shape in {"wavenet.conv_layers.3.conv.weight_v": (256, 256, 3), "wavenet.conv_layers.3.conv.weight_g": (256, 1, 1),
                     "wavenet.conv_layers.19.conv1x1c.weight_v": (256, 64, 1), "wavenet.conv_layers.0.conv1x1g.weight_g": (256, 1, 1),
                     "wavenet.conv_layers.7.conv1x1_out.bias": (256,), "wavenet.conv_layers.7.conv1x1_skip.weight_v": (256, 128, 1),
                     "wavenet.first_conv.weight_v": (256, 256, 1), "wavenet.last_conv_layers.3.bias": (256,),
                     "wavenet.embed_speakers.weight": (153, 32), "wavenet.upsample_net.conv_in.weight": (64, 64, 1),
                     "wavenet.upsample_net.upsample.up_layers.5.weight_v": (1, 1, 1, 17),
                     "encoder.net.2.conv.weight": (256, 256, 5), "encoder.lin.weight": (64, 256), "vq.embedding.weight": (256, 64)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert wn.receptive_field == 4093 == receptive_field_size(20, 2, 3)
    assert [f.conv.dilation[0] for f in wn.conv_layers] == [2 ** (i % 10) for i in range(20)]
    assert wn.has_speaker_embedding() and wn.local_conditioning_enabled()
    wn.make_generation_fast_()
    assert "conv_layers.0.conv.weight" in wn.state_dict() and "conv_layers.0.conv.weight_g" not in wn.state_dict()
    with pytest.raises(AssertionError):
        WaveNet(layers=5, stacks=2)


def test_api_errors_match_reference():
    m = build_model("tiny", 1)
    x, _, c, g = T.synth_inputs(T.CONFIGS["tiny"], 1, 32, 0)
    with pytest.raises(Exception):                      # upsampled c length != T (wavenet.py:196-200)
        with torch.no_grad():
            m(x[:, :, :31], c, g)
    m.train()
    with pytest.raises(RuntimeError):                   # conv.py:19-20
        m.incremental_forward(c=c, g=g, T=32)


def test_training_path_is_differentiable_and_matches_oracle():
    g = load_golden("wavenet_tiny")
    cfg = T.CONFIGS["tiny"]
    m = build_model("tiny", int(g["seed"])).train()
    x, _, c, spk = T.synth_inputs(cfg, int(g["B"]), int(g["T"]), int(g["in_seed"]))
    # default train_impl="kernels": a gradient request that the CUDA kernels cannot serve (CPU tensors, fp32) fails loudly, in
    # train and in eval mode -- no silent torch-op fallback; the composite is an explicit opt-in (test infrastructure)
    for mode in (m.train, m.eval):
        mode()
        with pytest.raises(_lib.WaeError, match="no silent fallback"):
            m(x, c, spk)
    m.train()
    m.train_impl = "autograd"
    y = m(x, c, spk)
    assert rel_err(y.detach().numpy(), g["logits"]) < 2e-5
    y.square().mean().backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    # the last layer's residual 1x1 never receives a gradient
    assert grads["conv_layers.3.conv1x1_out.weight_v"] is None or float(grads["conv_layers.3.conv1x1_out.weight_v"].abs().sum()) == 0
    assert float(grads["conv_layers.0.conv.weight_v"].abs().sum()) > 0


@pytest.mark.parametrize("cfg_name", ["tiny", "tiny_k2", "vqwae", "inwae"])
def test_packing_layouts(cfg_name):
    cfg = T.CONFIGS[cfg_name]
    m = build_model(cfg_name, 3)
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    p = wo.extract_params(sd, cfg["layers"], cfg["stacks"])
    sh = packing.stack_shape(m)
    assert (sh.layers, sh.kernel_size, sh.R, sh.G, sh.S, sh.C, sh.Gi, sh.O) == (
        cfg["layers"], cfg["kernel_size"], cfg["residual_channels"], cfg["gate_channels"], cfg["skip_out_channels"],
        cfg["cin_channels"], cfg["gin_channels"], cfg["out_channels"])
    H, L, kw, R = sh.H, sh.layers, sh.kernel_size, sh.R
    lay = p["layers"][L - 1]
    # fp32 pack: [K1][G] with pair-permuted columns
    pf = packing.pack_f32(m)
    w1 = pf.t["w1"][L - 1].numpy()
    assert w1.shape == (kw * R + sh.C, sh.G)
    q, i = 5, 2
    np.testing.assert_allclose(w1[1 * R + 7, 8 * q + i], lay["w"][4 * q + i, 7, 1], rtol=1e-6)           # tanh ch, tap 1
    np.testing.assert_allclose(w1[kw * R + 3, 8 * q + 4 + i], lay["wc"][H + 4 * q + i, 3], rtol=1e-6)   # sigmoid ch, cond row
    np.testing.assert_allclose(pf.t["w2"][0].numpy()[9, R + 11], p["layers"][0]["ws"][11, 9], rtol=1e-6)
    # bf16 pack: K-major [2*Hh][K1p]; gate halves padded to Hh = round16(H), split into two passes beyond 128 channels
    pb = packing.pack_bf16(m)
    w1b = pb.t["w1"][L - 1].float().numpy()
    Hh = -(-H // 16) * 16
    Ha = min(Hh, 128)
    assert w1b.shape[1] % 64 == 0 and w1b.shape[0] == 2 * Hh

    def gate_row(half, ch):          # row of tanh (0) / sigmoid (1) channel ch
        return half * Ha + ch if ch < Ha else 2 * Ha + half * (Hh - Ha) + (ch - Ha)
    for ch in {3, H - 1}:
        for half in (0, 1):
            np.testing.assert_allclose(w1b[gate_row(half, ch), (kw - 1) * R + 5], lay["w"][half * H + ch, 5, kw - 1], rtol=1e-2)
    if Hh > H:
        assert float(np.abs(w1b[gate_row(0, H):gate_row(0, Hh - 1) + 1]).sum()) == 0                       # padded channels are zero
    assert float(np.abs(w1b[:, kw * R + sh.C:]).sum()) == 0                                                # K padding is zero
    np.testing.assert_allclose(pb.t["bs_sum"].numpy(), sum(l["bs"] for l in p["layers"]), rtol=1e-5, atol=1e-6)
    # AR pack: per-(stage, rank) row slices
    for cs, wt in ((8, "fp32"), (16, "bf16")):
        pa = packing.pack_ar(m, cluster=cs, wtype=wt)
        offs = pa.t["layer_off"].numpy().reshape(2 * L + 2, cs)
        assert (offs % 16 == 0).all() and (np.diff(offs.reshape(-1)) >= 0).all()
        dt = np.float32 if wt == "fp32" else None
        r = cs - 1
        p0, p1 = packing.part(H, r, cs), packing.part(H, r + 1, cs)
        K1p = kw * R + (-(-sh.C // 64) * 64 if sh.C else 0)
        blob = pa.t["blob"]
        n = 2 * (p1 - p0) * K1p
        raw = blob[offs[2 * (L - 1), r]: offs[2 * (L - 1), r] + n * (4 if wt == "fp32" else 2)]
        rows = (raw.view(torch.float32) if wt == "fp32" else raw.view(torch.bfloat16).float()).numpy()
        rows = rows.reshape(K1p // 32, 2 * (p1 - p0), 32).transpose(1, 0, 2).reshape(-1, K1p)   # undo [K/32][rows][32]
        np.testing.assert_allclose(rows[0, :R], lay["w"][p0, :, 0], rtol=1e-2)            # tanh row of first pair, oldest tap
        np.testing.assert_allclose(rows[1, :R], lay["w"][H + p0, :, 0], rtol=1e-2)        # its sigmoid partner
    # cache invalidation: an in-place parameter update changes the fingerprint
    fp = packing.params_fingerprint(m)
    with torch.no_grad():
        next(m.parameters()).add_(1.0)
    assert packing.params_fingerprint(m) != fp


def test_vq_module_api():
    for cls, names in ((vqm.VectorQuantize, {"embedding.weight"}), (vqm.SlicedVectorQuantize, {"embedding1.weight", "embedding2.weight"}),
                       (vqm.VectorQuantizeEMA, {"embedding.weight", "ema_cluster_size", "ema_w"}),
                       (vqm.SlicedVectorQuantizeEMA, {"embedding1.weight", "embedding2.weight", "ema_cluster_size1", "ema_w1",
                                                      "ema_cluster_size2", "ema_w2"})):
        m = cls(16, 8)
        assert set(m.state_dict()) == names
    m = vqm.SlicedVectorQuantize(16, 8, K1=4)
    assert m.embedding2.weight.shape == (4, 4) and float(m.embedding1.weight.abs().max()) <= 1 / 16


@pytest.mark.parametrize("cfg_name", ["tiny", "tiny_k2", "vqwae"])
def test_hand_derived_stack_backward_matches_autograd_float64(cfg_name):
    """training.stack_backward (what the GPU training path runs in bf16 on the activations its tcgen05 forward saves) in
    float64 against torch autograd over the same layer equations: every weight, bias, the conditioning and the speaker vector."""
    import math
    from wavenet_autoencoders_b200 import packing, training
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    cfg = T.CONFIGS[cfg_name]
    torch.manual_seed(0)
    m = WaveNet(**cfg).train()
    m.load_state_dict(T.synth_state_dict(m, 4))
    sh = packing.stack_shape(m)
    L, R, H, C, kw = sh.layers, sh.R, sh.H, sh.C, sh.kernel_size
    B, Tn = (2, 96) if cfg_name != "vqwae" else (1, 1280)      # vqwae: all 20 layers, dilations up to 512 against T = 1280
    x, idx, c, spk = T.synth_inputs(cfg, B, Tn, 5)
    D = torch.float64
    with torch.no_grad():
        c_up0 = m.upsample_net(c).to(D)
        g0 = m._speaker_vectors(spk, B).to(D)
        y_mod = m._forward_autograd(x, c_up0.float(), g0.float(), False)
    ws = [None if w is None else w.detach().to(D).requires_grad_(True) for w in training.live_weights(m)]
    c_up, gv = c_up0.clone().requires_grad_(True), g0.clone().requires_grad_(True)
    P = training.PER_LAYER
    Hp, Cp = -(-H // 64) * 64, -(-C // 64) * 64
    base = L * P
    xs, hs = [], []
    cur = (ws[base][:, :, 0] @ x.to(D)).transpose(1, 2) + ws[base + 1]                 # (B,T,R) channels-last
    ccl = c_up.transpose(1, 2)
    skips = 0
    for l in range(L):
        W1, b1, Wc, Wg, Wo, bo, Ws, bs = ws[l * P: l * P + P]
        xs.append(cur)
        z = sum(training._shift(cur, (kw - 1 - j) * sh.dilations[l]) @ W1[:, :, j].t() for j in range(kw)) + b1
        z = z + ccl @ Wc[:, :, 0].t() + (gv @ Wg[:, :, 0].t())[:, None, :]
        h = torch.tanh(z[..., :H]) * torch.sigmoid(z[..., H:])
        hs.append(h)
        skips = skips + h @ Ws[:, :, 0].t() + bs
        cur = (h @ Wo[:, :, 0].t() + bo + cur) * math.sqrt(0.5)
    s = torch.relu(skips * math.sqrt(1.0 / L))
    s = torch.relu(s @ ws[base + 2][:, :, 0].t() + ws[base + 3])
    logits = (s @ ws[base + 4][:, :, 0].t() + ws[base + 5]).transpose(1, 2)            # (B,O,T)
    assert rel_err(logits.detach().float().numpy(), y_mod.numpy()) < 1e-5                # the same function as the module
    cot = torch.randn_like(logits)
    ins = [w for w in ws if w is not None] + [c_up, gv]
    ref = torch.autograd.grad((logits * cot).sum(), ins, allow_unused=True)
    x_all = torch.stack([t.detach() for t in xs])
    h_all = torch.nn.functional.pad(torch.stack([t.detach() for t in hs]), (0, Hp - H))
    c_cl = torch.nn.functional.pad(ccl.detach(), (0, Cp - C))
    _, dc, dg, grads = training.stack_backward(sh, list(sh.dilations), x.to(D), gv.detach(), x_all, h_all, c_cl,
                                               [None if w is None else w.detach() for w in ws], cot, cdt=D, adt=D)
    got = [g for g, w in zip(grads, ws) if w is not None] + [dc, dg]
    names = [i for i, w in enumerate(ws) if w is not None] + ["c", "g"]
    for n, a, b in zip(names, ref, got):
        if a is None:                                   # the last layer's residual 1x1: no gradient under autograd either
            assert b is None or float(b.abs().max()) == 0, n
            continue
        assert b is not None, n
        assert rel_err(b.numpy(), a.numpy()) < 1e-9, (n, rel_err(b.numpy(), a.numpy()))


def test_library_sass_contains_tcgen05_tma_and_cluster_instructions():
    """The built library really is sm_100a tensor-core / TMA code: cuobjdump's SASS of libwae_b200.so holds the tcgen05 MMA
    (UTCHMMA, also .2CTA), TMEM loads (LDTM), tcgen05.commit (UTCBAR), TMA tensor loads / stores (UTMALDG / UTMASTG), bulk
    copies (UBLKCP, the AR kernel's weight stream) and the AR kernel's mma.sync (HMMA.16816.F32.BF16) -- the mnemonics
    /opt/skills/guides/B200_PROFILING.md names as proof.  Skipped where cuobjdump is not installed."""
    import re
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or ("/usr/local/cuda/bin/cuobjdump" if os.path.exists("/usr/local/cuda/bin/cuobjdump") else None)
    if exe is None:
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "wavenet_autoencoders_b200", "libwae_b200.so")
    r = subprocess.run([exe, "-sass", so], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert "sm_100a" in r.stdout
    counts = {m: len(re.findall(r"\b" + re.escape(m), r.stdout))
              for m in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA.16816.F32.BF16")}
    missing = [m for m, n in counts.items() if n == 0]
    assert not missing, f"SASS lacks {missing}: {counts}"


def test_representation_text_dump_matches_numpy_savetxt(tmp_path):
    """wae_dump_text (host-only entry point of the C ABI) == np.savetxt(fmt='%.6f'), the ABX feature dump of
    inference_2019.py:262: byte-identical incl. negative zero, tiny and large values."""
    import numpy as np
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    rs = np.random.RandomState(0)
    a = (rs.randn(211, 64) * np.exp(rs.randn(211, 64) * 4)).astype(np.float32)
    a[0, :6] = [0.0, -0.0, 1e-7, -5e-7, 123456.789, 0.4999995]
    VQVAE.dump_representation(tmp_path / "a.txt", a)
    np.savetxt(tmp_path / "b.txt", a, fmt="%.6f")
    assert (tmp_path / "a.txt").read_bytes() == (tmp_path / "b.txt").read_bytes()
    VQVAE.dump_representation(tmp_path / "c.txt", a[:3, :1], decimals=3)
    np.savetxt(tmp_path / "d.txt", a[:3, :1], fmt="%.3f")
    assert (tmp_path / "c.txt").read_bytes() == (tmp_path / "d.txt").read_bytes()


def test_frontend_and_encoder_structs_for_the_fused_kernels():
    """Host logic of the round-2 fusions (no GPU): packing.pack_frontend expresses ConvInUpsampleNetwork / UpsampleNetwork as
    the plain-array struct of wae_stack_forward_bf16_lat (or declines), Encoder.fused_struct mirrors vqvae_model.py:25-51 for
    wae_encoder_vq_forward and wae_encoder_vq_supported agrees with the documented limits, frame arithmetic per SURVEY 9."""
    import torch
    from wavenet_autoencoders_b200 import _lib, packing, testing as T
    from wavenet_autoencoders_b200.vqvae_model import Encoder
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    wn = WaveNet(**T.VQWAE).eval()
    fe = packing.pack_frontend(wn)
    assert fe is not None and fe.total_scale == 640 and fe.struct.n_stages == 4
    assert [fe.struct.scale[i] for i in range(4)] == [4, 4, 8, 5]
    assert fe.struct.conv_in_w_t and all(fe.struct.filter[i] for i in range(4))
    wt = fe.keep[0]
    assert torch.equal(wt, wn.upsample_net.conv_in.weight.detach()[:, :, 0].t())          # [in][out]
    plain = WaveNet(**dict(T.VQWAE, upsample_net="UpsampleNetwork", upsample_params={"upsample_scales": [16, 40], "cin_channels": 64}))
    fp = packing.pack_frontend(plain)
    assert fp is not None and fp.struct.conv_in_w_t is None and fp.total_scale == 640
    padded = WaveNet(**dict(T.VQWAE, upsample_params={"upsample_scales": [4, 4, 8, 5], "cin_channels": 64, "cin_pad": 2}, cin_pad=2))
    assert packing.pack_frontend(padded) is None                                          # conv_in with context: staged path
    fine = WaveNet(**dict(T.VQWAE, upsample_params={"upsample_scales": [2, 2], "cin_channels": 64, "cin_pad": 0}))
    assert packing.pack_frontend(fine) is None                                            # 32 latent frames per 128-sample tile: staged path
    assert packing.pack_frontend(WaveNet(**dict(T.VQWAE, upsample_conditional_features=False))) is None

    enc = Encoder(hid=256, c_in=39, c_out=64)
    st = enc.fused_struct()
    assert st is not None and st.n_layers == 10 and st.hid == 256 and st.D == 64
    spec = [(st.layer[i].cin, st.layer[i].cout, st.layer[i].k, st.layer[i].stride, st.layer[i].relu, st.layer[i].residual) for i in range(10)]
    assert spec[0] == (39, 256, 3, 1, 1, 0) and spec[1] == (256, 256, 3, 1, 1, 1) and spec[2] == (256, 256, 5, 2, 1, 0)
    assert spec[3] == (256, 256, 5, 2, 1, 0) and spec[4][2:] == (3, 1, 1, 1) and all(s_[2:] == (1, 1, 1, 1) for s_ in spec[6:])
    assert Encoder(hid=768).fused_struct() is None                                        # wider than the kernel's 256 channels
    assert Encoder(hid=64, c_in=13, c_out=16).fused_struct() is not None
    assert [enc.out_frames(f) for f in (1, 2, 100, 300, 50)] == [1, 1, 25, 75, 13]        # 50 frames -> 13 latents
    L = _lib.lib()
    assert L.wae_encoder_vq_workspace(16, 100) == 4096 + 16 * 2 * 256 * 136 * 4
    assert L.wae_encoder_vq_workspace(2, 1000) == 4096 + 2 * 11 * 2 * 256 * 136 * 4      # 250 latents -> 11 blocks of 24 per utterance
    lanes = packing._NoLanes()
    with lanes.lane(3):
        pass
    lanes.join()
