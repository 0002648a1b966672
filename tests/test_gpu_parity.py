"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle and the reference's
golden vectors on identical seeded inputs.  Nothing here reads /root/reference."""
import math

import numpy as np
import pytest
import torch

from conftest import build_model, load_golden, rel_err
from oracle import sampling, vq_oracle
from oracle import wavenet_oracle as wo
from wavenet_autoencoders_b200 import _lib, testing as T
from wavenet_autoencoders_b200 import vector_quantization as vqm

pytestmark = pytest.mark.gpu

# stated tolerances (max|diff| / max|ref|)
TOL_FP32 = 1e-3     # BASELINE north_star: "decoder logits within 1e-3 relative in fp32" (we see ~1e-5)
TOL_BF16 = 3.5e-2   # bf16 operands + bf16 residual stream through 20 layers, fp32 accumulate: the "stated looser bound" = 2x the
                    # largest error measured on any golden (1.6e-2, tools/measure_tolerances.py); per golden below
TOL_BF16_CASE = {"wavenet_tiny": 1.3e-2, "wavenet_tiny_k2": 1.3e-2, "wavenet_vqwae": 2.6e-2, "wavenet_inwae": 3.3e-2}   # 2x measured


def _inputs(case):
    g = load_golden(case)
    cfg_name = str(g["cfg"])
    cfg = T.CONFIGS[cfg_name]
    m = build_model(cfg_name, int(g["seed"]), "cuda")
    x, idx, c, spk = T.synth_inputs(cfg, int(g["B"]), int(g["T"]), int(g["in_seed"]))
    return g, cfg, m, x.cuda(), c.cuda(), spk.cuda()


def test_library_reports_sm100():
    _lib.check(_lib.lib().wae_device_check(0), "wae_device_check")


# ------------------------------------------------------------------ tcgen05 building block
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 256, 832), (128 * 5 + 7, 32, 128), (4096, 256, 2560)])
def test_tcgen05_gemm_matches_fp32(M, N, K):
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(_lib.lib().wae_gemm_bf16_tn(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), M, N, K, _lib.stream_ptr()), "gemm")
    ref = a.float() @ b.float().t()
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5


# ------------------------------------------------------------------ teacher-forced stack
@pytest.mark.parametrize("case", ["wavenet_tiny", "wavenet_tiny_k2", "wavenet_tiny_mol", "wavenet_vqwae", "wavenet_inwae"])
def test_forward_fp32_matches_reference_golden(case):
    g, cfg, m, x, c, spk = _inputs(case)
    m.precision = "fp32"
    with torch.no_grad():
        y = m(x, c, spk)
        c_up = m.upsample_net(c)
    s = int(g["stride"])
    assert rel_err(c_up[:, :, ::s].cpu().numpy(), g["c_up"]) < 1e-5
    assert y.shape == (int(g["B"]), cfg["out_channels"], int(g["T"]))
    assert rel_err(y[:, :, ::s].cpu().numpy(), g["logits"]) < TOL_FP32
    assert rel_err(y[:, :, ::s].cpu().numpy(), g["logits"]) < 1e-4   # what the fp32 kernels actually achieve
    with torch.no_grad():
        ys = m(x, c, spk, softmax=True)
    with pytest.raises(_lib.WaeError, match="no silent fallback"):     # grad mode + fp32 kernels: fails loudly (no torch-op fallback)
        m(x, c, spk)
    assert torch.allclose(ys.sum(1), torch.ones_like(ys.sum(1)), atol=1e-4)


@pytest.mark.parametrize("case", ["wavenet_tiny", "wavenet_tiny_k2", "wavenet_vqwae", "wavenet_inwae"])
def test_forward_bf16_matches_reference_golden(case):
    g, cfg, m, x, c, spk = _inputs(case)
    m.precision = "bf16"
    with torch.no_grad():
        y = m(x, c, spk)
    s = int(g["stride"])
    err = rel_err(y[:, :, ::s].cpu().numpy(), g["logits"])
    assert err < TOL_BF16_CASE[case], err


def test_programmatic_launch_chain_is_bit_identical(monkeypatch):
    """The version-4 layer kernels and the head are launched programmatically behind each other (griddepcontrol: the next
    kernel's prologue runs under the previous kernel's tail, its body waits for the previous kernel to complete).  The logits
    must be bit-identical to the fully serialised chain (WAE_PDL=0) -- eagerly, back to back without a synchronisation in
    between (the case in which a missing wait would show), and inside a replayed CUDA graph."""
    g, cfg, m, x, c, spk = _inputs("wavenet_vqwae_b2")
    m.precision = "bf16"
    with torch.no_grad():
        monkeypatch.setenv("WAE_PDL", "0")
        y_serial = m(x, c, spk).clone()
        monkeypatch.setenv("WAE_PDL", "1")
        ys = [m(x, c, spk).clone() for _ in range(4)]
        torch.cuda.synchronize()
        for y in ys:
            assert torch.equal(y, y_serial)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(x, c, spk)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            y_graph = m(x, c, spk)
        for _ in range(3):
            y_graph.zero_()
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(y_graph, y_serial)


@pytest.mark.parametrize("case", ["wavenet_tiny", "wavenet_vqwae"])
def test_forward_bf16_fused_last_upsample_stage(case):
    """wae_stack_forward_bf16_up (last upsampler stage fused into the stack) against the unfused call on the
    materialised (B,C,T) conditioning, and the mixed one-hot / dense first-conv path inside one block."""
    g, cfg, m, x, c, spk = _inputs(case)
    m.precision = "bf16"
    m.fuse_frontend = False        # the whole-front-end kernel (next test) would take over otherwise
    with torch.no_grad():
        assert m.upsample_net(c, defer_last=True) is not None          # the fused path is the one forward() takes
        y_fused = m(x, c, spk)
        c_up = m.upsample_net(c)
        y_plain = m.stack_forward(x, c_up, m._speaker_vectors(spk, x.shape[0]))
        assert rel_err(y_fused.cpu().numpy(), y_plain.cpu().numpy()) < 5e-3
        x2 = x.clone()
        x2[:, :, 5::7] = torch.softmax(torch.randn_like(x2[:, :, 5::7]), dim=1)   # every 7th sample dense
        m.precision = "fp32"
        y32 = m(x2, c, spk)
        m.precision = "bf16"
        y16 = m(x2, c, spk)
    assert rel_err(y16.cpu().numpy(), y32.cpu().numpy()) < TOL_BF16


@pytest.mark.parametrize("case", ["wavenet_tiny", "wavenet_vqwae", "wavenet_inwae"])
def test_forward_bf16_fused_conditioning_frontend(case):
    """wae_stack_forward_bf16_lat: conv_in + every upsampler stage evaluated per 128-sample block inside the
    stack's conditioning kernel, from the latent frames.  (a) without conv_in the stage pyramid repeats the staged kernels'
    arithmetic: logits BIT-identical to the staged path; (b) with conv_in (an fp32 dot product in another summation order than
    the library matmul) the bf16 conditioning may differ in its last bit: logits within 2e-3; (c) reference golden; (d) the
    class-index input and the fused NLL go through the same front-end."""
    g, cfg, m, x, c, spk = _inputs(case)
    m.precision = "bf16"
    B, T = x.shape[0], x.shape[-1]
    idx = x.argmax(1)
    with torch.no_grad():
        assert m._pack("fe") is not None
        n0 = _lib.launch_count()
        y_fe = m(x, c, spk)
        n_fe = _lib.launch_count() - n0
        m.fuse_frontend = False
        n0 = _lib.launch_count()
        y_st = m(x, c, spk)
        n_st = _lib.launch_count() - n0
        assert n_fe == n_st - (len(m.upsample_net.upsample.scales) - 1), (n_fe, n_st)      # the stage launches are gone
        assert rel_err(y_fe.cpu().numpy(), y_st.cpu().numpy()) < 2e-3
        s = int(g["stride"])
        assert rel_err(y_fe[:, :, ::s].cpu().numpy(), g["logits"]) < TOL_BF16
        # (a) no conv_in: feed the staged path the conv_in output, the fused one an identity conv_in
        c_in = torch.matmul(m.upsample_net.conv_in.weight[:, :, 0], c)
        y_a = m.stack_forward(x, m.upsample_net.upsample(c_in), m._speaker_vectors(spk, B))
        m.fuse_frontend = True
        w_keep = m.upsample_net.conv_in.weight.detach().clone()
        m.upsample_net.conv_in.weight.copy_(torch.eye(w_keep.shape[0], device=w_keep.device).unsqueeze(-1))   # in place: bumps _version
        y_b = m(x, c_in, spk)
        m.upsample_net.conv_in.weight.copy_(w_keep)
        assert torch.equal(y_a, y_b)
        # (d) class indices + NLL from the head kernel
        y_idx = m(idx, c, spk)
        assert torch.equal(y_idx, m(x, c, spk))
        nll = m.forward_nll(idx, c, spk, idx, 1)
        ref = torch.nn.functional.cross_entropy(y_idx[:, :, :-1].double(), idx[:, 1:])
        assert abs(float(nll) - float(ref)) < 2e-5 * max(1.0, abs(float(ref)))
        with pytest.raises(Exception):
            m(x[:, :, :-cfg["upsample_params"]["upsample_scales"][-1]], c, spk)       # c does not cover T: same failure as the reference


@pytest.mark.parametrize("G", [40, 296, 368, 512])
def test_forward_bf16_gate_padding_and_two_pass(G):
    """Gate widths the tcgen05 layer kernel pads (H % 16 != 0) or splits into two accumulator passes (G > 256; IN-WAE has
    G = 368), against the numpy oracle on a small stack."""
    cfg = dict(T.CONFIGS["tiny"], residual_channels=64, gate_channels=G, skip_out_channels=64, layers=4, stacks=2,
               upsample_conditional_features=False)
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(0)
    m = WaveNet(**cfg).eval()
    m.load_state_dict(T.synth_state_dict(m, 5))
    p = wo.extract_params({k: v.numpy() for k, v in m.state_dict().items()}, cfg["layers"], cfg["stacks"])
    m = m.cuda()
    rs = np.random.RandomState(G)
    B, Tn = 3, 300
    idx = rs.randint(0, cfg["out_channels"], size=(B, Tn))
    x = np.eye(cfg["out_channels"], dtype=np.float32)[idx].transpose(0, 2, 1).copy()
    c = rs.normal(size=(B, cfg["cin_channels"], Tn)).astype(np.float32)
    spk = rs.randint(0, cfg["n_speakers"], size=(B, 1))
    ref = wo.forward(p, x, c, spk)
    m.precision = "bf16"
    with torch.no_grad():
        y = m(torch.tensor(x).cuda(), torch.tensor(c).cuda(), torch.tensor(spk).cuda())
    assert rel_err(y.cpu().numpy(), ref) < TOL_BF16, G


def test_forward_bf16_from_class_indices_and_fused_nll():
    """Additive fast paths: (B,T) integer classes instead of the one-hot tensor (identical logits: the same row gather), and
    the one-pass teacher-forced NLL against torch's cross_entropy on the shifted slices (vqwae_train.py:760-766)."""
    from wavenet_autoencoders_b200.losses import teacher_forced_nll
    g, cfg, m, x, c, spk = _inputs("wavenet_tiny")
    m.precision = "bf16"
    idx = x.argmax(1)
    with torch.no_grad():
        y_onehot = m(x, c, spk)
        y_idx = m(idx, c, spk)
    assert torch.equal(y_onehot, y_idx)
    ref = torch.nn.functional.cross_entropy(y_idx[:, :, :-1], idx[:, 1:])
    got = teacher_forced_nll(y_idx, idx)
    assert abs(float(got) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    # NLL from the head kernel's accumulator, logits never written (wae_stack_nll_bf16_idx): row f2
    with torch.no_grad():
        fused = m.forward_nll(idx, c, spk, idx, 1)
    assert abs(float(fused) - float(ref)) < 1e-6 * max(1.0, abs(float(ref))), (float(fused), float(ref))
    m.precision = "fp32"                       # no index kernel there: expanded to one-hot on the host side
    with torch.no_grad():
        assert torch.equal(m(idx, c, spk), m(x, c, spk))


def test_head_fused_nll_at_benchmark_shape():
    """wae_stack_nll_bf16_idx at the vqwae shape (ragged T, several utterances): loss from the TMEM accumulator == F.cross_entropy on
    the logits the same kernels write; with logits_out the written logits are bit-identical to the plain forward's."""
    cfg = T.CONFIGS["vqwae"]
    m = build_model("vqwae", 1, "cuda")
    m.precision = "bf16"
    x, idx, c, spk = T.synth_inputs(cfg, 3, 1920, 5)
    idx, c, spk = idx.cuda(), c.cuda(), spk.cuda()
    with torch.no_grad():
        y = m(idx, c, spk)
        ref = torch.nn.functional.cross_entropy(y[:, :, :-1].double(), idx[:, 1:])
        got = m.forward_nll(idx, c, spk, idx, 1)
        lg = torch.empty_like(y)
        got2 = m._nll_from_indices(idx, c, spk, idx, 1, logits_out=lg)
    assert abs(float(got) - float(ref)) < 1e-6 * max(1.0, abs(float(ref))), (float(got), float(ref))
    assert float(got) == float(got2) or abs(float(got) - float(got2)) < 1e-6
    assert torch.equal(lg, y)


def test_forward_ragged_tail_and_dense_input_fp32_bf16():
    """T not a multiple of any tile size; dense (non one-hot) x exercises the first-conv GEMV path."""
    cfg = dict(T.CONFIGS["tiny"], upsample_conditional_features=False)
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    torch.manual_seed(0)
    m = WaveNet(**cfg).eval()
    m.load_state_dict(T.synth_state_dict(m, 9))
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    p = wo.extract_params(sd, cfg["layers"], cfg["stacks"])
    m = m.cuda()
    rs = np.random.RandomState(3)
    for B, Tn in [(1, 1), (2, 77), (3, 333)]:
        x = rs.uniform(size=(B, cfg["out_channels"], Tn)).astype(np.float32)
        x /= x.sum(1, keepdims=True)
        c = rs.normal(size=(B, cfg["cin_channels"], Tn)).astype(np.float32)
        spk = rs.randint(0, cfg["n_speakers"], size=(B, 1))
        ref = wo.forward(p, x, c, spk)
        for prec, tol in (("fp32", 1e-4), ("bf16", TOL_BF16)):
            m.precision = prec
            with torch.no_grad():
                y = m(torch.tensor(x).cuda(), torch.tensor(c).cuda(), torch.tensor(spk).cuda())
            assert rel_err(y.cpu().numpy(), ref) < tol, (B, Tn, prec)


def test_forward_full_size_properties_bf16():
    """BASELINE config 2 size (16 x 16000): causality and batch independence -- size-independent properties."""
    cfg = T.CONFIGS["vqwae"]
    m = build_model("vqwae", 1, "cuda")
    m.precision = "bf16"
    x, idx, c, spk = T.synth_inputs(cfg, 16, 16000, 5)
    x, c, spk = x.cuda(), c.cuda(), spk.cuda()
    with torch.no_grad():
        y = m(x, c, spk)
        assert torch.isfinite(y).all()
        # batch independence: utterance 3 alone gives the same logits
        y3 = m(x[3:4], c[3:4], spk[3:4])
        assert rel_err(y3.cpu().numpy(), y[3:4].cpu().numpy()) < 1e-6
        # causality: changing the input at t >= 9000 leaves logits before 9000 untouched
        x2 = x.clone()
        x2[:, :, 9000:] = x2[:, :, 9000:].roll(1, dims=1)
        y2 = m(x2, c, spk)
        assert torch.equal(y2[:, :, :9000], y[:, :, :9000])
        assert not torch.equal(y2[:, :, 9000:], y[:, :, 9000:])
    # fp32 kernels on a slice of the same batch agree with the bf16 ones within the stated bf16 bound (the last
    # ~850 samples of the slice see a different upsampled conditioning: the smoothing filters reach over the cut)
    m.precision = "fp32"
    with torch.no_grad():
        yf = m(x[:2, :, :6400], c[:2, :, :10], spk[:2])
    err = rel_err(y[:2, :, :5000].cpu().numpy(), yf[:, :, :5000].cpu().numpy())
    assert err < TOL_BF16, err


# ------------------------------------------------------------------ autoregressive synthesis
@pytest.mark.parametrize("case,prec,cluster,tol", [
    ("wavenet_tiny", "fp32", 16, 1e-4), ("wavenet_tiny", "fp32", 8, 1e-4), ("wavenet_tiny", "bf16", 8, 3e-2),
    ("wavenet_tiny_k2", "fp32", 8, 1e-4), ("wavenet_tiny_k2", "bf16", 16, 3e-2),
    ("wavenet_tiny", "bf16-simt", 8, 3e-2), ("wavenet_tiny_k2", "bf16-simt", 8, 3e-2)])
def test_incremental_teacher_forced_logits(case, prec, cluster, tol):
    """L1 of the RNG contract: identical test_inputs for all T, softmax=False, quantize=False -> per-step logits.
    "bf16" runs the tensor-core (mma.sync) AR kernel, "bf16-simt" the CUDA-core kernel on bf16 weights."""
    g, cfg, m, x, c, spk = _inputs(case)
    if prec.endswith("-simt"):
        prec, m.ar_impl = "bf16", "simt"
    m.precision, m.ar_cluster = prec, cluster
    y = m.incremental_forward(initial_input=x[:, :, :1], c=c, g=spk, T=int(g["T"]), test_inputs=x, softmax=False, quantize=False)
    assert y.shape == tuple(g["inc_logits"].shape)
    assert rel_err(y.cpu().numpy(), g["inc_logits"]) < tol
    if prec == "fp32":   # free-running with probabilities fed back (dense first-conv path inside the kernel)
        Tf = int(g["Tfree"])
        yf = m.incremental_forward(initial_input=x[:, :, :1], c=c[:, :, :3], g=spk, T=Tf, test_inputs=x[:, :, :1],
                                   softmax=True, quantize=False)
        assert rel_err(yf.cpu().numpy(), g["free_probs"]) < 1e-3


def test_incremental_sampling_matches_oracle_free_running():
    """L2 + L3: the fused categorical sampler consumes the supplied uniforms; free-running classes are compared with
    the oracle driven by the same uniforms (first divergence-free window must cover the whole run in fp32)."""
    g, cfg, m, x, c, spk = _inputs("wavenet_tiny")
    m.precision, m.ar_cluster = "fp32", 8
    B, Tn = 2, 160
    u = torch.rand(Tn, B, generator=torch.Generator().manual_seed(3))
    init = x[:, :, :1]
    out = m.incremental_forward(initial_input=init, c=c[:, :, :Tn // T.hop(cfg)], g=spk, T=Tn, softmax=True, quantize=True,
                                uniforms=u.cuda())
    assert out.shape == (B, cfg["out_channels"], Tn) and torch.equal(out.sum(1), torch.ones(B, Tn, device="cuda"))
    got = m.last_sampled_indices.cpu().numpy()
    sd = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    p = wo.extract_params(sd, cfg["layers"], cfg["stacks"])
    O = cfg["out_channels"]
    picks = []

    def sampler(t, logits):
        k = np.array([sampling.categorical_from_uniform(logits[b], float(u[t, b])) for b in range(B)])
        picks.append(k)
        return np.eye(O, dtype=np.float32)[k]
    wo.incremental_forward(p, Tn, c=c[:, :, :Tn // T.hop(cfg)].cpu().numpy(), g=spk.cpu().numpy(),
                           initial_input=init[:, :, 0].cpu().numpy(), sampler=sampler)
    want = np.stack(picks, 1)
    first_div = int(np.argmax((got != want).any(0))) if (got != want).any() else Tn
    assert first_div == Tn, f"first divergence at step {first_div}"


def test_incremental_mol_and_padding():
    """Scalar-input model (mixture of logistics), B not a multiple of utterances-per-cluster, ranks with no logit rows."""
    cfg = T.CONFIGS["tiny_mol"]
    m = build_model("tiny_mol", 7, "cuda")
    m.precision, m.ar_cluster = "fp32", 8
    B, Tn = 3, 64
    x, _, c, spk = T.synth_inputs(cfg, B, Tn, 8)
    nmix = cfg["out_channels"] // 3
    u = torch.rand(Tn, B, nmix + 1, generator=torch.Generator().manual_seed(5))
    y = m.incremental_forward(initial_input=None, c=c.cuda(), g=spk.cuda(), T=Tn, uniforms=u.cuda())
    assert y.shape == (B, 1, Tn) and float(y.abs().max()) <= 1.0
    sd = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    p = wo.extract_params(sd, cfg["layers"], cfg["stacks"])
    ref = wo.incremental_forward(
        p, Tn, c=c.numpy(), g=spk.numpy(), initial_input=np.zeros((B, 1), np.float32),
        sampler=lambda t, lg: np.array([[sampling.mol_from_uniform(lg[b], u[t, b].numpy())] for b in range(B)], np.float32))
    np.testing.assert_allclose(y[:, 0].cpu().numpy(), ref[:, :, 0], atol=2e-4)


# ------------------------------------------------------------------ VQ
@pytest.mark.parametrize("case", ["vq_plain_default", "vq_plain_trained", "vq_sliced_default", "vq_sliced_trained"])
def test_vq_bit_exact_vs_oracle_and_golden(case):
    g = load_golden(case)
    kind, K, D = str(g["kind"]), int(g["K"]), int(g["D"])
    mod = getattr(vqm, kind)(K, D).cuda().eval()
    with torch.no_grad():
        for n, p_ in mod.named_parameters():
            p_.copy_(torch.tensor(g["param_" + n.replace(".", "__")]))
        quant, loss, perp = mod(torch.tensor(g["x"]).cuda())
    if kind == "VectorQuantize":
        oq, ol, op, oi = vq_oracle.vq_forward(g["x"], g["param_embedding__weight"])
    else:
        oq, ol, op, oi = vq_oracle.sliced_vq_forward(g["x"], g["param_embedding1__weight"], g["param_embedding2__weight"])
    np.testing.assert_array_equal(mod.last_codes.cpu().numpy(), oi)            # bit-exact indices vs oracle
    np.testing.assert_array_equal(quant.cpu().numpy(), oq)                    # bit-exact quantised output
    assert abs(loss.item() - float(ol)) <= 1e-6 * abs(float(ol)) + 1e-10
    assert abs(perp.item() - float(op)) <= 1e-5 * float(op)
    if case.endswith("trained"):                                              # ... and vs the reference itself
        np.testing.assert_array_equal(quant.cpu().numpy(), g["quant"])
        assert abs(loss.item() - float(g["vq_loss"])) <= 1e-5 * float(g["vq_loss"])
        assert abs(perp.item() - float(g["perp"])) <= 1e-5 * float(g["perp"])


def test_vq_large_and_edge_cases():
    rs = np.random.RandomState(0)
    # ragged sizes: N not a multiple of the 64-vector tile, K not a multiple of the 128-code chunk, odd D
    for B, D, Tn, K in [(1, 64, 1, 256), (3, 24, 37, 100), (2, 64, 1000, 300), (5, 7, 13, 3)]:
        x = (rs.normal(size=(B, D, Tn)) * 0.5).astype(np.float32)
        cb = (rs.normal(size=(K, D)) * 0.5).astype(np.float32)
        mod = vqm.VectorQuantize(K, D).cuda()
        with torch.no_grad():
            mod.embedding.weight.copy_(torch.tensor(cb))
            q, loss, perp = mod(torch.tensor(x).cuda())
        oq, ol, op, oi = vq_oracle.vq_forward(x, cb)
        np.testing.assert_array_equal(mod.last_codes.cpu().numpy(), oi)
        np.testing.assert_array_equal(q.cpu().numpy(), oq)
    # duplicate codewords: first index wins (argmin semantics)
    mod = vqm.VectorQuantize(4, 2).cuda()
    with torch.no_grad():
        mod.embedding.weight.copy_(torch.tensor([[1.0, 1.0], [0.0, 0.0], [0.0, 0.0], [1.0, 1.0]]))
        mod(torch.zeros(1, 2, 5).cuda())
    assert mod.last_codes.cpu().tolist() == [[1] * 5]
    # idempotence at scale (size-independent property): quantising the quantised output returns the same codes
    x = torch.randn(64, 64, 4096, device="cuda") * 0.5
    mod = vqm.VectorQuantize(256, 64).cuda()
    with torch.no_grad():
        mod.embedding.weight.normal_(0, 0.5)
        mod(x)
        codes = mod.last_codes.clone()
        q = mod.embedding.weight[codes].permute(0, 2, 1).contiguous()
        mod(q)
    assert torch.equal(mod.last_codes, codes)
    assert int(torch.bincount(codes.flatten(), minlength=256).sum()) == 64 * 4096


def test_vq_autograd_and_ema():
    x = (torch.randn(4, 64, 25, device="cuda") * 0.5).requires_grad_(True)
    mod = vqm.VectorQuantize(64, 64).cuda()
    mod.embedding.weight.data.normal_(0, 0.5)
    q, loss, perp = mod(x)
    (q.sum() + loss).backward()
    assert x.grad is not None and mod.embedding.weight.grad is not None
    # straight-through: d(quant)/dx = identity
    assert torch.allclose(x.grad, torch.ones_like(x) + torch.autograd.grad(mod(x)[1], x)[0], atol=1e-6)
    # EMA variant: one training step against the oracle restatement
    ema = vqm.VectorQuantizeEMA(32, 16).cuda().train()
    ema.embedding.weight.data.normal_(0, 0.5)
    cb0 = ema.embedding.weight.detach().cpu().numpy().copy()
    xe = torch.randn(3, 16, 50, device="cuda") * 0.5
    q, loss, perp = ema(xe)
    idx, _, _ = vq_oracle.search(xe.cpu().numpy(), cb0)
    flat = xe.permute(0, 2, 1).reshape(-1, 16).cpu().numpy()
    size, w, cb1 = vq_oracle.ema_update(flat, idx, 32, np.zeros(32, np.float32), np.zeros((32, 16), np.float32), 0.99)
    np.testing.assert_array_equal(ema.last_codes.cpu().numpy().reshape(-1), idx)
    np.testing.assert_allclose(ema.embedding.weight.detach().cpu().numpy(), cb1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(q.detach().cpu().numpy(), cb1[idx].reshape(3, 50, 16).transpose(0, 2, 1), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ composition
def test_vqvae_forward_matches_reference_golden():
    g = load_golden("vqvae_tiny")
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    cfg = T.CONFIGS["tiny"]
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=cfg["cin_channels"], K=32, wavenet=WaveNet(**cfg), encoder_hid=48).eval()
    m.load_state_dict(T.synth_state_dict(m, 5))
    m = m.cuda()
    x = torch.nn.functional.one_hot(torch.tensor(g["idx"]), cfg["out_channels"]).float().transpose(1, 2).contiguous().cuda()
    with torch.no_grad():
        y, vq_loss, perp = m(x, torch.tensor(g["mfcc"]).cuda(), torch.tensor(g["g"]).cuda())
        quant = m.encode(torch.tensor(g["mfcc"]).cuda())
    assert rel_err(quant.cpu().numpy(), g["quant"]) < 1e-5       # encoder runs in cuDNN; codes must still agree
    assert rel_err(y.cpu().numpy(), g["logits"]) < TOL_FP32
    assert abs(vq_loss.item() - float(g["vq_loss"])) < 1e-4 * float(g["vq_loss"])
    assert abs(perp.item() - float(g["perp"])) < 1e-4 * float(g["perp"])


# ------------------------------------------------------------------ training: tcgen05 forward + GEMM backward
def _flat_stats(ref, got):
    fa = torch.cat([t.double().flatten() for t in ref])
    fb = torch.cat([t.double().flatten() for t in got])
    return float((fa * fb).sum() / (fa.norm() * fb.norm())), float((fb - fa).norm() / fa.norm())


def _per_tensor_stats(ref, got):
    """(lowest cosine, highest relative L2 error) over the tensors, each compared on its own: a wrong gradient of ONE small
    tensor (a bias, a weight_g) cannot hide inside the norm of the whole gradient vector."""
    cos, l2 = 1.0, 0.0
    for a, b in zip(ref, got):
        a, b = a.double().flatten(), b.double().flatten()
        if float(a.norm()) == 0.0:
            assert float(b.norm()) == 0.0
            continue
        cos = min(cos, float((a @ b) / (a.norm() * b.norm())))
        l2 = max(l2, float((a - b).norm() / a.norm()))
    return cos, l2


@pytest.mark.parametrize("cfg_name", ["tiny", "tiny_k2", "vqwae"])
def test_training_backward_matches_autograd(cfg_name):
    """Gradients of the teacher-forced NLL through training.StackTrainFunction (bf16 kernels forward, hand-derived backward).
    The derivation itself is checked in float64 against autograd on the CPU (tests/test_host_cpu.py, < 1e-9, also at this
    20-layer shape).  Here, on the GPU:
      (a) the bf16 backward on the tensor-core kernels (wae_stack_backward_bf16, csrc/wn_bwd.cu) against the SAME derivation
          evaluated as fp32 torch expressions on the SAME saved activations -- isolates the kernels' bf16 arithmetic;
      (b) end to end against torch autograd over the fp32 composite (true fp32: no TF32) -- additionally contains the bf16
          forward, whose ReLU masks / gate saturations differ slightly from the fp32 forward's.
    Gradients of bias-like quantities are heavily cancelling sums over time, so bf16 rounding noise is visible on them;
    the criteria are cosine similarity and relative L2 error of the whole gradient vector."""
    from wavenet_autoencoders_b200 import training
    cfg = T.CONFIGS[cfg_name]
    m = build_model(cfg_name, 3, "cuda").train()
    B, Tn = (3, 320) if cfg_name != "vqwae" else (2, 1280)      # vqwae: the real 20-layer shape, two latent frames
    x, idx, c, spk = T.synth_inputs(cfg, B, Tn, 11)
    x, idx, spk = x.cuda(), idx.cuda(), spk.cuda()
    m.precision = "bf16"
    out = {}
    for impl in ("autograd", "kernels"):
        m.train_impl = impl
        m.zero_grad(set_to_none=True)
        cc = c.cuda().clone().requires_grad_(True)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            y = m(x, cc, spk)
            loss = torch.nn.functional.cross_entropy(y[:, :, :-1], idx[:, 1:])
            loss.backward()
        out[impl] = (float(loss), cc.grad.clone(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    l0, dc0, g0 = out["autograd"]
    l1, dc1, g1 = out["kernels"]
    assert abs(l0 - l1) < 2e-2 * max(1.0, abs(l0))
    assert set(g0) == set(g1), set(g0) ^ set(g1)
    names = sorted(g0)
    cos_b, l2_b = _flat_stats([g0[n] for n in names] + [dc0], [g1[n] for n in names] + [dc1])
    pcos_b, pl2_b = _per_tensor_stats([g0[n] for n in names] + [dc0], [g1[n] for n in names] + [dc1])

    # (a): same saved activations, bf16 kernels vs fp32 torch expressions
    with torch.no_grad():
        c_up, gv = m.upsample_net(c.cuda()), m._speaker_vectors(spk, B)

    class Ctx:
        def save_for_backward(self, *ts):
            self.saved = ts
    ctx = Ctx()
    logits = training.StackTrainFunction.forward(ctx, m, x, c_up, gv, *training.live_weights(m))
    lg = logits.clone().requires_grad_(True)
    torch.nn.functional.cross_entropy(lg[:, :, :-1], idx[:, 1:]).backward()
    xf, gf, x_all, h_all, c_cl, r1, r2, *wts = ctx.saved
    f32 = lambda t: None if t is None else t.float()
    assert training.tc_backward_supported(ctx.sh)                  # (a) exercises wae_stack_backward_bf16, not the library composite
    with torch.no_grad():
        ra = training.stack_backward(ctx.sh, ctx.dil, xf, gf, x_all, h_all, c_cl, wts, lg.grad, r1=r1, r2=r2, pk=ctx.pk,
                                     gate=getattr(ctx, "gate", None))
        if getattr(ctx, "gate", None) is not None:
            # the same backward with the gate GEMM recomputed (no kept factors): the two differ by the bf16 rounding of the kept
            # tanh / sigmoid only
            rc_ = training.stack_backward(ctx.sh, ctx.dil, xf, gf, x_all, h_all, c_cl, wts, lg.grad, r1=r1, r2=r2, pk=ctx.pk)
            keep_ = [i for i, w in enumerate(wts) if w is not None and ra[3][i] is not None]
            cs_, l2_ = _flat_stats([rc_[3][i] for i in keep_] + [rc_[1], rc_[2]], [ra[3][i] for i in keep_] + [ra[1], ra[2]])
            print(f"{cfg_name}: kept gate factors vs recomputed gate GEMM: cosine {cs_:.6f} rel L2 {l2_:.3e}")
            assert cs_ > 0.9995 and l2_ < 3e-2, (cs_, l2_)
        rb = training.stack_backward(ctx.sh, ctx.dil, xf, gf, f32(x_all), f32(h_all), f32(c_cl), wts, lg.grad, cdt=torch.float32)
    keep = [i for i, w in enumerate(wts) if w is not None and ra[3][i] is not None]
    cos_a, l2_a = _flat_stats([rb[3][i] for i in keep] + [rb[1], rb[2]], [ra[3][i] for i in keep] + [ra[1], ra[2]])
    pcos_a, pl2_a = _per_tensor_stats([rb[3][i] for i in keep] + [rb[1], rb[2]], [ra[3][i] for i in keep] + [ra[1], ra[2]])
    print(f"{cfg_name}: (a) bf16 vs fp32 backward on the same activations: cosine {cos_a:.5f} rel L2 {l2_a:.3e}, worst tensor "
          f"{pcos_a:.5f} / {pl2_a:.3e};  (b) vs fp32 autograd end to end: cosine {cos_b:.5f} rel L2 {l2_b:.3e}, worst tensor "
          f"{pcos_b:.5f} / {pl2_b:.3e}")
    # bounds = 2x the measured deviations (tools/measure_tolerances.py; whole vector and worst single tensor)
    bound = {"tiny": dict(a=(0.9990, 0.05), pa=(0.995, 0.10), b=(0.994, 0.11), pb=(0.988, 0.23)),
             "tiny_k2": dict(a=(0.9990, 0.05), pa=(0.995, 0.10), b=(0.994, 0.11), pb=(0.988, 0.23)),
             "vqwae": dict(a=(0.994, 0.16), pa=(0.97, 0.30), b=(0.978, 0.30), pb=(0.94, 0.50))}[cfg_name]
    for (c_, l_), key in (((cos_a, l2_a), "a"), ((pcos_a, pl2_a), "pa"), ((cos_b, l2_b), "b"), ((pcos_b, pl2_b), "pb")):
        assert c_ > bound[key][0] and l_ < bound[key][1], (key, c_, l_, bound[key])


def test_training_kernels_match_torch():
    """The gather / element-wise kernels of the backward pass (csrc/wn_train.cu) against their torch expressions."""
    from wavenet_autoencoders_b200 import training
    L = _lib.lib()
    st = _lib.stream_ptr()
    B, Tn, R, C, Cp, kw, H = 3, 200, 64, 16, 64, 3, 184        # H = 184: the IN-WAE gate half (23 chunks of 8)
    bf = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, Tn, R, device="cuda", generator=g).to(bf)
    c = torch.nn.functional.pad(torch.randn(B, Tn, C, device="cuda", generator=g), (0, Cp - C)).to(bf)
    for d in (1, 7, 64, 512):                                  # dilation beyond T: every tap but the newest is zero padding
        out = torch.empty(B, Tn, kw * R + Cp, dtype=bf, device="cuda")
        _lib.check(L.wae_train_im2col(_lib.ptr(x), _lib.ptr(c), B, Tn, R, Cp, kw, d, _lib.ptr(out), st), "im2col")
        ref = torch.cat([training._shift(x, (kw - 1 - j) * d) for j in range(kw)] + [c], dim=-1)
        assert torch.equal(out, ref), d
        dxcat = torch.randn(B, Tn, kw * R + Cp, device="cuda", generator=g).to(bf)
        dxo = torch.randn(B, Tn, R, device="cuda", generator=g).to(bf)
        dC = torch.ones(B, Tn, C, device="cuda")
        dx = torch.empty(B, Tn, R, dtype=bf, device="cuda")
        _lib.check(L.wae_train_dx_accum(_lib.ptr(dxcat), _lib.ptr(dxo), B, Tn, R, C, Cp, kw, d, 0.5, _lib.ptr(dx), _lib.ptr(dC), st),
                   "dx_accum")
        want = dxo.float()
        for j in range(kw):
            want = want + training._unshift(dxcat[..., j * R:(j + 1) * R], (kw - 1 - j) * d).float()
        assert rel_err(dx.float().cpu().numpy(), (want * 0.5).cpu().numpy()) < 1e-2
        assert torch.allclose(dC, 1.0 + dxcat[..., kw * R: kw * R + C].float())
    z = torch.randn(B, Tn, 2 * H, device="cuda", generator=g).to(bf)
    gb = torch.randn(B, 2 * H, device="cuda", generator=g)
    wide = torch.randn(B, Tn, 3 * 192, device="cuda", generator=g).to(bf)          # dh_a is a strided view, like dHskip[..., l*Hp:]
    dh_a = wide[..., 192: 192 + H]
    dh_b = torch.randn(B, Tn, H, device="cuda", generator=g).to(bf)
    dz = torch.empty(B, Tn, 2 * H, dtype=bf, device="cuda")
    dgb = torch.zeros(B, 2 * H, device="cuda")
    _lib.check(L.wae_train_gate_bwd(_lib.ptr(z), _lib.ptr(gb), dh_a.data_ptr(), 3 * 192, _lib.ptr(dh_b), B, Tn, H, _lib.ptr(dz),
                                    _lib.ptr(dgb), st), "gate_bwd")
    zz = z.float() + gb[:, None, :]
    th, sg = torch.tanh(zz[..., :H]), torch.sigmoid(zz[..., H:])
    dh = dh_a.float() + dh_b.float()
    want = torch.cat([dh * sg * (1 - th * th), dh * th * sg * (1 - sg)], dim=-1)
    assert rel_err(dz.float().cpu().numpy(), want.cpu().numpy()) < 1e-2
    assert rel_err(dgb.cpu().numpy(), want.sum(1).cpu().numpy()) < 1e-3


@pytest.mark.parametrize("frames", [100, 37, 1])
def test_encoder_kernels_match_torch_convs(frames):
    """wae_conv1d_relu_res (one launch per ConvReLURes block + the final Linear) against the torch/cuDNN fp32 path of the same
    module (vqvae_model.py:9-51): strides 1 and 2, kernel sizes 1/3/5, residual and non-residual blocks, ragged lengths."""
    from wavenet_autoencoders_b200.vqvae_model import Encoder
    torch.manual_seed(0)
    enc = Encoder(hid=256, c_in=39, c_out=64).cuda().eval()
    x = torch.randn(3, 39, frames, device="cuda")
    with torch.no_grad():
        got = enc(x)                                        # kernels (inference, CUDA)
        assert enc._kernels_ok(x)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ref = enc.lin(enc.net(x).permute(0, 2, 1)).permute(0, 2, 1)
    assert got.shape == ref.shape
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    w0 = enc.net[0].conv.weight
    with torch.no_grad():
        w0.mul_(2.0)                                        # in-place parameter update: the transposed-weight cache must notice
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ref2 = enc.lin(enc.net(x).permute(0, 2, 1)).permute(0, 2, 1)
        assert rel_err(enc(x).cpu().numpy(), ref2.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("frames,B,sliced", [(100, 16, False), (37, 3, True), (1, 2, False), (128, 2, False), (129, 2, False), (300, 3, True), (1000, 1, False)])
def test_encoder_vq_fused_kernel(frames, B, sliced):
    """wae_encoder_vq_forward (encoder + Linear + VQ search in one launch, cluster of 8 CTAs per item) against
    (a) the torch/cuDNN fp32 encoder (latents <= 1e-5 relative; another summation order), (b) the standalone search kernel
    wae_vq_search run on the fused kernel's own latents: codes, quantised values, loss and perplexity BIT-identical -- the
    search arithmetic is the same; (c) utterances longer than 128 frames (24-latent blocks with a recomputed halo)."""
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=64, K=256, wavenet=None, encoder_hid=256).eval()
    if sliced:
        m.vq = vqm.SlicedVectorQuantize(256, 64)
    m.load_state_dict(T.synth_state_dict(m, 3))
    m = m.cuda()
    x = torch.randn(B, 39, frames, device="cuda")
    with torch.no_grad():
        enc = m.encoder.fused_struct()
        assert enc is not None and m.vq._fusable(x)
        n0 = _lib.launch_count()
        quant, loss, perp = m._encode_quantize(x)
        assert _lib.launch_count() - n0 == 2                 # the fused kernel + the statistics finaliser
        codes = m.vq.last_codes.clone()
        F4 = m.encoder.out_frames(frames)
        q2, l2, p2 = m.vq._forward_fused(enc, x, F4, want_latents=True)
        lat = m.vq.last_latents
        assert torch.equal(q2, quant) and torch.equal(m.vq.last_codes, codes)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ref = m.encoder.lin(m.encoder.net(x).permute(0, 2, 1)).permute(0, 2, 1)
        assert lat.shape == ref.shape
        assert rel_err(lat.cpu().numpy(), ref.cpu().numpy()) < 1e-5
        if frames <= 300:                                    # the numpy oracle (pinned to the reference's latents, test_oracle_cpu.py)
            from oracle import encoder_oracle as eo
            olat = eo.encoder_forward({k: v.cpu().numpy() for k, v in m.state_dict().items() if k.startswith("encoder.")}, x.cpu().numpy())
            assert rel_err(lat.cpu().numpy(), olat) < 1e-5
        qs, ls, ps = m.vq(lat)                               # standalone search kernel on the same latents
        assert torch.equal(m.vq.last_codes, codes)
        assert torch.equal(qs.contiguous(), quant)
        assert abs(float(ls) - float(loss)) <= 1e-6 * abs(float(ls)) and abs(float(ps) - float(perp)) <= 1e-6 * abs(float(ps))
        m.fuse_encoder = False                               # module by module: same codes unless a latent sits on a near-tie
        qu, lu, pu = m._encode_quantize(x)
        agree = float((m.vq.last_codes == codes).float().mean())
        assert agree > 0.995, agree


def test_encode_batch_ragged_matches_per_utterance_encode(tmp_path):
    """VQVAE.encode_batch (inference_2019.py:225-262 batched): utterances of different lengths in ONE launch (per-utterance
    lengths inside the fused kernel) == encode() utterance by utterance, bit for bit; the text dump equals np.savetxt('%.6f')."""
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=64, K=256, wavenet=None, encoder_hid=256).eval()
    m.load_state_dict(T.synth_state_dict(m, 3))
    m = m.cuda()
    rs = np.random.RandomState(5)
    lens = [100, 37, 1, 128, 64, 99, 5]
    feats = [rs.normal(size=(n, 39)).astype(np.float32) for n in lens]
    n0 = _lib.launch_count()
    reps = m.encode_batch(feats)
    assert _lib.launch_count() - n0 == 2
    for f, rep in zip(feats, reps):
        one = m.encode(torch.tensor(f.T[None]).cuda())[0].t().cpu().numpy()          # (T', D), the reference's per-utterance call
        assert rep.shape == one.shape
        np.testing.assert_array_equal(rep, one)
    long_feats = [rs.normal(size=(n, 39)).astype(np.float32) for n in (300, 131, 700)]    # > 128 frames: tiled items, ragged
    for f, rep in zip(long_feats, m.encode_batch(long_feats)):
        np.testing.assert_array_equal(rep, m.encode(torch.tensor(f.T[None]).cuda())[0].t().cpu().numpy())
    m.dump_representation(tmp_path / "rep.txt", reps[0])
    np.savetxt(tmp_path / "ref.txt", reps[0], fmt="%.6f")
    assert (tmp_path / "rep.txt").read_bytes() == (tmp_path / "ref.txt").read_bytes()


def test_flat_adam_matches_torch_adam_with_clipping():
    """train_step.FlatAdam (wae_sumsq + wae_adam_step on one flat buffer) against clip_grad_norm_ + torch.optim.Adam."""
    from wavenet_autoencoders_b200.train_step import FlatAdam
    torch.manual_seed(0)

    def make():
        torch.manual_seed(1)
        return torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Tanh(), torch.nn.Linear(53, 11)).cuda()
    ma, mb = make(), make()
    oa = torch.optim.Adam(ma.parameters(), lr=4e-4, betas=(0.9, 0.999), eps=1e-8)
    ob = FlatAdam(mb, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, clip=0.05, ema_decay=0.9)
    shadow = [p.detach().clone() for p in ma.parameters()]
    for step in range(5):
        x = torch.randn(64, 37, device="cuda") * (10.0 if step % 2 else 0.1)      # steps that clip and steps that do not
        for m_, o_ in ((ma, oa), (mb, ob)):
            o_.zero_grad()
            m_(x).square().mean().backward()
        torch.nn.utils.clip_grad_norm_(ma.parameters(), 0.05)
        oa.step()
        ob.step()
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=1e-7), step
        for sh, pa in zip(shadow, ma.parameters()):                   # vqwae_train.py:346-350
            sh -= (1.0 - 0.9) * (sh - pa.detach())
        for sh, (pb, eb) in zip(shadow, ob.ema_state().items()):
            assert torch.allclose(sh, eb, rtol=2e-5, atol=1e-7), step
    assert float(ob.step_a) == 5.0
    # the torch.optim surface the reference's loop uses: a scheduled rate written into param_groups every step
    # (vqwae_train.py:730-735), model.zero_grad() detaching .grad, state_dict round trip for checkpoints
    sd = ob.state_dict()
    mc = make()
    with torch.no_grad():
        for pc, pb in zip(mc.parameters(), mb.parameters()):
            pc.copy_(pb)
    oc = FlatAdam(mc, lr=1.0, betas=(0.9, 0.999), eps=1e-8, clip=0.05, ema_decay=0.9)
    oc.load_state_dict(sd)
    for step in range(5, 8):
        lr = 4e-4 * 0.5 ** (step - 4)
        for g_ in oa.param_groups + ob.param_groups + oc.param_groups:
            g_["lr"] = lr
        x = torch.randn(64, 37, device="cuda")
        for m_ in (ma, mb, mc):
            m_.zero_grad()                                           # set_to_none: FlatAdam must pick the fresh .grad tensors up
            m_(x).square().mean().backward()
        torch.nn.utils.clip_grad_norm_(ma.parameters(), 0.05)
        oa.step(); ob.step(); oc.step()
        for pa, pb, pc in zip(ma.parameters(), mb.parameters(), mc.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=1e-7), step
            assert torch.equal(pb, pc), step                         # resumed optimiser == the one that kept running
            assert pb.grad.data_ptr() >= ob.flat_g.data_ptr()        # views restored


def test_vq_codebook_view_at_unaligned_offset():
    """Codebooks that are views into a flat parameter buffer (train_step.FlatAdam) may start at any 4-byte offset."""
    torch.manual_seed(0)
    K, D, B, Tn = 64, 32, 2, 37
    flat = torch.randn(K * D + 3, device="cuda")
    x = torch.randn(B, D, Tn, device="cuda")
    outs = []
    for off in (0, 1, 3):
        cb = flat[off:off + K * D].view(K, D)
        cb.copy_(flat[:K * D].view(K, D).clone() if off else cb)
        idx = torch.empty(B * Tn, dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().wae_vq_search(_lib.ptr(x), B, D, Tn, 0, D, cb.data_ptr(), K, _lib.ptr(idx), None, None, None,
                                            _lib.stream_ptr()), "wae_vq_search")
        ref = torch.cdist(x.permute(0, 2, 1).reshape(-1, D), cb).argmin(1)
        assert (idx == ref).float().mean() > 0.99                      # cdist's arithmetic differs in the last ulp on near-ties
        outs.append(idx)


def test_vq_resident_kernel_bit_exact_vs_oracle_and_chunked():
    """Many vectors: wae_vq_search switches to the codebook-resident persistent kernel.  Codes and quantised values must equal
    the oracle's AND the chunked kernel's bit for bit, for ragged tiles, K off the 128-code chunk, odd / wide slices (sub_d
    not a multiple of 4, sub_d > 64: no register prefetch) and a codebook too large to stay resident (falls back)."""
    rs = np.random.RandomState(3)
    L = _lib.lib()
    try:
        for B, D, Tn, K in [(3, 64, 2001, 300), (2, 24, 3000, 100), (2, 7, 3000, 3), (2, 132, 2500, 64), (1, 64, 5000, 1024)]:
            x = (rs.normal(size=(B, D, Tn)) * 0.5).astype(np.float32)
            cb = (rs.normal(size=(K, D)) * 0.5).astype(np.float32)
            oq, ol, op, oi = vq_oracle.vq_forward(x, cb)
            mod = vqm.VectorQuantize(K, D).cuda()
            outs = []
            for variant in (0, 1):
                _lib.check(L.wae_vq_set_variant(variant), "wae_vq_set_variant")
                with torch.no_grad():
                    mod.embedding.weight.copy_(torch.tensor(cb))
                    q, loss, perp = mod(torch.tensor(x).cuda())
                np.testing.assert_array_equal(mod.last_codes.cpu().numpy(), oi)
                np.testing.assert_array_equal(q.cpu().numpy(), oq)
                assert abs(float(loss) - float(ol)) <= 1e-5 * abs(float(ol))
                assert abs(float(perp) - float(op)) <= 1e-4 * abs(float(op))
                outs.append((mod.last_codes.clone(), q.clone()))
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        # sliced module (two 32-wide halves of a 64-channel latent), many vectors
        x = (rs.normal(size=(4, 64, 2500)) * 0.5).astype(np.float32)
        cb1 = (rs.normal(size=(256, 32)) * 0.5).astype(np.float32)
        cb2 = (rs.normal(size=(256, 32)) * 0.5).astype(np.float32)
        _lib.check(L.wae_vq_set_variant(0), "wae_vq_set_variant")
        mod = vqm.SlicedVectorQuantize(256, 64).cuda()
        with torch.no_grad():
            mod.embedding1.weight.copy_(torch.tensor(cb1))
            mod.embedding2.weight.copy_(torch.tensor(cb2))
            q, loss, perp = mod(torch.tensor(x).cuda())
        i1, _, _ = vq_oracle.search(x, cb1, 0, 32)
        i2, _, _ = vq_oracle.search(x, cb2, 32, 32)
        codes = mod.last_codes.cpu().numpy().reshape(-1, 2)
        np.testing.assert_array_equal(codes[:, 0], np.asarray(i1).reshape(-1))
        np.testing.assert_array_equal(codes[:, 1], np.asarray(i2).reshape(-1))
    finally:
        L.wae_vq_set_variant(0)


def test_graphed_forward_matches_eager():
    """GraphedForward (the inference forward + NLL captured as one CUDA graph) returns what the eager call returns, also after
    new inputs are copied into the captured buffers (pinned host tensors included)."""
    from wavenet_autoencoders_b200.graphed import GraphedForward
    from wavenet_autoencoders_b200.losses import teacher_forced_nll
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    cfg = T.CONFIGS["tiny"]
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=cfg["cin_channels"], K=32, wavenet=WaveNet(**cfg), encoder_hid=48).eval()
    m.load_state_dict(T.synth_state_dict(m, 5))
    m = m.cuda()
    m.wavenet.precision = "bf16"
    g0 = load_golden("vqvae_tiny")
    idx = torch.tensor(g0["idx"]).cuda()
    mfcc = torch.tensor(g0["mfcc"]).cuda()
    spk = torch.tensor(g0["g"]).cuda()
    gf = GraphedForward(m, idx, mfcc, spk)
    assert gf.launches and gf.launches > 0
    gen = torch.Generator().manual_seed(7)
    for trial in range(3):
        if trial:
            idx_h = torch.randint(0, cfg["out_channels"], idx.shape, generator=gen).pin_memory()
            mfcc_h = torch.randn(mfcc.shape, generator=gen).pin_memory()
            spk_h = torch.randint(0, cfg["n_speakers"], spk.shape, generator=gen).pin_memory()
        else:
            idx_h, mfcc_h, spk_h = idx, mfcc, spk
        logits, vq_loss, perp, nll = gf(idx_h, mfcc_h, spk_h)
        with torch.no_grad():
            y, vl, pp = m(idx_h.cuda(), mfcc_h.cuda(), spk_h.cuda())
            ref_nll = teacher_forced_nll(y, idx_h.cuda())
        assert torch.equal(logits, y)
        assert abs(float(nll) - float(ref_nll)) <= 1e-6 * max(1.0, abs(float(ref_nll)))
        assert abs(float(vq_loss) - float(vl)) <= 1e-6 * max(1.0, abs(float(vl)))
        assert abs(float(perp) - float(pp)) <= 1e-5 * max(1.0, abs(float(pp)))


def test_graphed_forward_pipelined_results():
    """GraphedForward.submit / result (losses read one step behind): every ticket returns the scalars the synchronous call gives
    for ITS inputs, with two steps in flight and the slots reused."""
    from wavenet_autoencoders_b200.graphed import GraphedForward
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    cfg = T.CONFIGS["tiny"]
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=cfg["cin_channels"], K=32, wavenet=WaveNet(**cfg), encoder_hid=48).eval()
    m.load_state_dict(T.synth_state_dict(m, 5))
    m = m.cuda()
    m.wavenet.precision = "bf16"
    g0 = load_golden("vqvae_tiny")
    idx, mfcc, spk = torch.tensor(g0["idx"]).cuda(), torch.tensor(g0["mfcc"]).cuda(), torch.tensor(g0["g"]).cuda()
    gf = GraphedForward(m, idx, mfcc, spk, with_logits=False)
    gen = torch.Generator().manual_seed(11)
    batches = [(torch.randint(0, cfg["out_channels"], idx.shape, generator=gen).pin_memory(),
                torch.randn(mfcc.shape, generator=gen).pin_memory(),
                torch.randint(0, cfg["n_speakers"], spk.shape, generator=gen).pin_memory()) for _ in range(5)]
    want = []
    for b in batches:
        _, vq_loss, perp, nll = gf(*b)
        want.append((float(nll), float(vq_loss), float(perp)))
    assert len({w[0] for w in want}) == len(want)          # the batches do give different losses
    got, prev = [], None
    for b in batches:
        t = gf.submit(*b)
        if prev is not None:
            got.append(gf.result(prev))
        prev = t
    got.append(gf.result(prev))
    for gvals, wvals in zip(got, want):                    # the NLL is summed with double atomics: equal up to the summation order
        for a, b in zip(gvals, wvals):
            assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (got, want)


def test_synthesis_postprocess_matches_oracle():
    """wae_synth_postprocess (inverse mu-law + inverse pre-emphasis + gain, synthesis.py:382-394) against the float64 numpy /
    scipy restatement: all three input types, ragged T (shorter than the 256 segments, not a multiple of them), T = 48000."""
    from oracle import postprocess_oracle as po
    from wavenet_autoencoders_b200.postprocess import waveform_from_synthesis
    rs = np.random.RandomState(11)
    for B, Tn in [(1, 1), (3, 100), (2, 1000), (4, 48000)]:
        idx = rs.randint(0, 256, size=(B, Tn))
        for post, coef, gain in [(None, 0.85, 0.0), ("inv_preemphasis", 0.85, 0.0), ("inv_preemphasis", 0.97, 0.55)]:
            got = waveform_from_synthesis(torch.tensor(idx).cuda(), "mulaw-quantize", 256, post, coef, gain).cpu().numpy()
            ref = po.waveform(idx, "mulaw-quantize", 256, post, coef, gain)
            assert got.shape == (B, Tn) and got.dtype == np.float32
            np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5 * max(1.0, float(np.abs(ref).max())))
            if post is None and gain == 0.0:
                np.testing.assert_array_equal(got, ref.astype(np.float32))          # the table is exact to fp32 rounding
        y = rs.uniform(-1, 1, size=(B, Tn)).astype(np.float32)
        for kind in ("mulaw", "raw"):
            got = waveform_from_synthesis(torch.tensor(y).cuda()[:, None, :], kind, 256, "inv_preemphasis", 0.85, 2.0).cpu().numpy()
            ref = po.waveform(y, kind, 256, "inv_preemphasis", 0.85, 2.0)
            np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5 * max(1.0, float(np.abs(ref).max())))
    # one-hot input takes the argmax like the reference (synthesis.py:383)
    idx = rs.randint(0, 256, size=(2, 50))
    onehot = torch.nn.functional.one_hot(torch.tensor(idx), 256).float().transpose(1, 2).cuda()
    a = waveform_from_synthesis(onehot, "mulaw-quantize", 256)
    b = waveform_from_synthesis(torch.tensor(idx).cuda(), "mulaw-quantize", 256)
    assert torch.equal(a, b)
    with pytest.raises(_lib.WaeError):
        waveform_from_synthesis(torch.tensor(idx), "mulaw-quantize", 256)             # CPU tensor: no fallback


@pytest.mark.parametrize("cfg_name,prec", [("tiny", "fp32"), ("tiny", "bf16"), ("tiny_mol", "fp32"), ("tiny_mol", "bf16"), ("vqwae", "bf16")])
def test_incremental_forward_with_fused_postprocess(cfg_name, prec):
    """wae_ar_generate_wave: inverse mu-law, inverse pre-emphasis and gain computed INSIDE the synthesis kernel by
    the thread that emits each sample.  Same sampled stream as the plain call (identical uniforms), and the waveform equals
    (a) the float64 oracle on those samples <= 2e-5 and (b) the separate wae_synth_postprocess launch <= 2e-5 (another
    summation order of the recurrence); without pre-emphasis the table lookup is exact."""
    from oracle import postprocess_oracle as po
    from wavenet_autoencoders_b200.postprocess import waveform_from_synthesis
    cfg = T.CONFIGS[cfg_name]
    m = build_model(cfg_name, 5, "cuda")
    m.precision = prec
    B, Tn = (3, 320) if cfg_name != "vqwae" else (2, 1280)
    x, idx, c, spk = T.synth_inputs(cfg, B, Tn, 21)
    scalar = bool(cfg["scalar_input"])
    nu = (cfg["out_channels"] // 3 + 1) if scalar else None
    gen = torch.Generator(device="cuda").manual_seed(7)
    u = torch.rand((Tn, B, nu) if scalar else (Tn, B), device="cuda", generator=gen)
    kind = "raw" if scalar else "mulaw-quantize"
    for post, coef, gain in ((None, 0.85, 0.0), ("inv_preemphasis", 0.97, 0.55)):
        kw = dict(input_type=kind, quantize_channels=cfg["out_channels"] if not scalar else 256, postprocess=post,
                  preemphasis_coef=coef, global_gain_scale=gain)
        with torch.no_grad():
            y0 = m.incremental_forward(initial_input=x[:, :, :1].cuda(), c=c.cuda(), g=spk.cuda(), T=Tn, uniforms=u,
                                       return_indices=not scalar)
            assert m.last_waveform is None
            y1 = m.incremental_forward(initial_input=x[:, :, :1].cuda(), c=c.cuda(), g=spk.cuda(), T=Tn, uniforms=u,
                                       return_indices=not scalar, wave_postprocess=kw)
        assert torch.equal(y0, y1)                                   # the post-processing does not disturb the synthesis
        wave = m.last_waveform
        assert wave is not None and wave.shape == (B, Tn)
        samples = y1.cpu().numpy() if not scalar else y1[:, 0].cpu().numpy()
        ref = po.waveform(samples, kind, kw["quantize_channels"], post, coef, gain)
        tol = 2e-5 * max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(wave.cpu().numpy(), ref, rtol=0, atol=tol)
        sep = waveform_from_synthesis(y1.long() if not scalar else y1, kind, kw["quantize_channels"], post, coef, gain)
        np.testing.assert_allclose(wave.cpu().numpy(), sep.cpu().numpy(), rtol=0, atol=tol)
        if post is None and not scalar:
            np.testing.assert_array_equal(wave.cpu().numpy(), ref.astype(np.float32))


def test_training_transpose_cast_matches_torch():
    """wae_train_transpose_cast: (B,O,T) fp32 -> (B,T,O) bf16, bit-identical to torch's transposing copy (round to nearest even),
    ragged tiles included."""
    for B, O, Tn in [(1, 2, 1), (3, 30, 77), (2, 256, 1000), (2, 66, 33)]:
        torch.manual_seed(B + O + Tn)
        x = torch.randn(B, O, Tn, device="cuda")
        out = torch.full((B, Tn, O), float("nan"), dtype=torch.bfloat16, device="cuda")
        _lib.check(_lib.lib().wae_train_transpose_cast(_lib.ptr(x), B, O, Tn, _lib.ptr(out), _lib.stream_ptr()), "wae_train_transpose_cast")
        ref = torch.empty(B, Tn, O, dtype=torch.bfloat16, device="cuda")
        ref.copy_(x.transpose(1, 2))
        assert torch.equal(out, ref)


# ------------------------------------------------------------------ round 2: the benchmarked shapes, pinned to the reference
def _b3(t):
    """(2, ...) -> ragged batch of 3: [u0, u1, u0] (utterance 2 must reproduce utterance 0 exactly)."""
    return torch.cat([t, t[:1]], dim=0).contiguous()


# measured on B200 (GPUTEST r2): the bf16 tensor-core AR kernel is within 1.0e-2 (vqwae) / 1.2e-2 (inwae) of the reference's
# fp32 logits over all 2560 steps; asserted at ~2x that
TOL_AR_BF16 = 2.5e-2


@pytest.mark.parametrize("case,prec,cluster,upc,tol", [
    ("wavenet_vqwae_b2", "bf16", 8, 8, TOL_AR_BF16), ("wavenet_vqwae_b2", "bf16", 16, 8, TOL_AR_BF16),
    ("wavenet_vqwae_b2", "bf16", 16, 2, TOL_AR_BF16), ("wavenet_vqwae_b2", "fp32", 16, 2, 1e-4),
    ("wavenet_vqwae_b2", "bf16-simt", 8, 2, TOL_AR_BF16),
    ("wavenet_inwae_b2", "bf16", 8, 8, TOL_AR_BF16), ("wavenet_inwae_b2", "fp32", 16, 2, 1e-4)])
def test_incremental_logits_at_benchmarked_shapes(case, prec, cluster, upc, tol):
    """L1 at the shapes bench.py times (hps/vqwae.json, hps/inae_hp.json; 20 layers, R = 256, dilations to 512), against
    the REFERENCE's incremental_forward output (golden `inc_logits`): T = 2560 so that the 1025-row history of the d = 512
    layers wraps twice and the t-1024 taps are live, B = 3 ragged ([u0, u1, u0]: a partially filled utterance group / a
    second cluster with one utterance).  "bf16": ar_mma_kernel (the 226.5 KB shared layout, cp.async.bulk weight rings,
    history rings in global memory); "fp32" / "bf16-simt": ar_kernel<float/bf16>."""
    g, cfg, m, x, c, spk = _inputs(case)
    if prec.endswith("-simt"):
        prec, m.ar_impl = "bf16", "simt"
    m.precision, m.ar_cluster, m.ar_utts_per_cluster = prec, cluster, upc
    x3, c3, s3 = _b3(x), _b3(c), _b3(spk)
    Tn, s = int(g["T"]), int(g["stride"])
    y = m.incremental_forward(initial_input=x3[:, :, :1], c=c3, g=s3, T=Tn, test_inputs=x3, softmax=False, quantize=False)
    want_kernel = "fp32" if prec == "fp32" else ("bf16" if m.ar_impl == "simt" else "bf16mma")
    if case == "wavenet_inwae_b2" and prec == "bf16":
        want_kernel = "bf16"       # H = 184 over 8 CTAs gives odd slice boundaries: the model falls to ar_kernel<bf16> (wavenet.py)
    assert m.last_ar_variant[0] == want_kernel and m.last_ar_variant[1] == cluster, m.last_ar_variant
    assert y.shape == (3, cfg["out_channels"], Tn)
    assert torch.equal(y[2], y[0])                                    # same utterance in another slot / cluster: bit-identical
    err = rel_err(y[:2, :, ::s].cpu().numpy(), g["inc_logits"])
    late = rel_err(y[:2, :, 2048::s].cpu().numpy(), g["inc_logits"][:, :, 2048 // s:])     # after both ring wraps
    print(f"{case} {prec} cluster {cluster} upc {upc}: rel err {err:.3e} (t >= 2048: {late:.3e})")
    assert err < tol and late < tol, (err, late)


def test_incremental_free_running_vqwae_first_divergence():
    """L3 at the benchmarked shape: free-running categorical synthesis, 2 utterances x 2560 steps on the golden's uniform
    stream.  The reference was driven by the same inverse-CDF sampler (tools/make_golden.py sampled_free_run), so the fp32
    kernel's classes must equal the reference's for the whole run (first divergence-free window = T).  The bf16 tensor-core
    kernel's window is REPORTED (its logits differ at the 1e-2 level, so it leaves the reference's trajectory early; north
    star: "identical sampled samples ... for the first divergence-free window")."""
    g, cfg, m, x, c, spk = _inputs("wavenet_vqwae_b2")
    Tn, B = int(g["T"]), int(g["B"])
    u = torch.rand(Tn, B, generator=torch.Generator().manual_seed(int(g["sampled_seed"]))).cuda()
    want = g["sampled"].astype(np.int64)

    def first_div(got):
        ne = (got != want).any(0)
        return int(np.argmax(ne)) if ne.any() else Tn
    m.precision, m.ar_cluster = "fp32", 16
    out = m.incremental_forward(initial_input=x[:, :, :1], c=c, g=spk, T=Tn, softmax=True, quantize=True, uniforms=u,
                                return_indices=True)
    d32 = first_div(out.cpu().numpy())
    m.precision, m.ar_cluster = "bf16", None
    out = m.incremental_forward(initial_input=x[:, :, :1], c=c, g=spk, T=Tn, softmax=True, quantize=True, uniforms=u,
                                return_indices=True)
    d16 = first_div(out.cpu().numpy())
    print(f"first divergence from the reference's sampled classes: fp32 kernel {d32} / {Tn}, bf16 mma kernel {d16} / {Tn}")
    assert d32 == Tn, f"fp32 AR kernel left the reference trajectory at step {d32}"
    assert d16 >= 1


@pytest.mark.parametrize("case,cfg_name", [("wavenet_tiny_mol_ar", "tiny_mol"), ("wavenet_tiny_gauss_ar", "tiny_gauss")])
def test_incremental_scalar_samplers_match_reference(case, cfg_name):
    """Scalar-input synthesis with the fused mixture-of-logistics (mixture.py:118-156) and mixture-of-Gaussians
    (mixture.py:221-270) samplers against the reference's incremental_forward fed the same draws (uniform_ / Normal.sample
    patched in tools/make_golden.py)."""
    g = load_golden(case)
    cfg = T.CONFIGS[cfg_name]
    m = build_model(cfg_name, int(g["seed"]), "cuda")
    B, Tn = int(g["B"]), int(g["T"])
    _, _, c, spk = T.synth_inputs(cfg, B, Tn, int(g["in_seed"]))
    for prec, atol in (("fp32", 2e-4), ("bf16", None), ("bf16-simt", None)):
        m.ar_impl = "simt" if prec.endswith("-simt") else "auto"
        m.precision, m.ar_cluster = prec.split("-")[0], 8
        y = m.incremental_forward(initial_input=None, c=c.cuda(), g=spk.cuda(), T=Tn, uniforms=torch.tensor(g["u"]).cuda(),
                                  log_scale_min=-7.0)
        # bf16: the tensor-core (mma.sync) kernel now carries the mixture samplers too; "-simt": ar_kernel<bf16>
        assert m.last_ar_variant[0] == {"fp32": "fp32", "bf16": "bf16mma", "bf16-simt": "bf16"}[prec], m.last_ar_variant
        assert y.shape == (B, 1, Tn) and float(y.abs().max()) <= 1.0
        if atol is not None:
            np.testing.assert_allclose(y[:, 0].cpu().numpy(), g["samples"], atol=atol)
        else:       # bf16 weights: same trajectory until rounding flips a mixture indicator; the first steps must agree closely
            np.testing.assert_allclose(y[:, 0, :8].cpu().numpy(), g["samples"][:, :8], atol=5e-2)


@pytest.mark.parametrize("case", ["vq_ema_plain", "vq_ema_sliced"])
def test_vq_ema_modules_match_reference(case):
    """VectorQuantizeEMA / SlicedVectorQuantizeEMA (vector_quantization.py:156-235, :257-306) in training mode for three steps
    and one eval step against the REFERENCE classes' own outputs and state (golden): codes via wae_vq_search, per-code sums
    via wae_vq_ema_stats, the codebook overwritten before the gather."""
    g = load_golden(case)
    K, D, steps = int(g["K"]), int(g["D"]), int(g["steps"])
    mod = getattr(vqm, str(g["kind"]))(K, D).cuda().train()
    with torch.no_grad():
        for n, p_ in mod.named_parameters():
            p_.copy_(torch.tensor(g["param_" + n.replace(".", "__")]))
    for s in range(steps + 1):
        if s == steps:
            mod.eval()
        with torch.no_grad():
            q, loss, perp = mod(torch.tensor(g[f"x{s}"]).cuda())
        np.testing.assert_allclose(q.cpu().numpy(), g[f"quant{s}"], rtol=2e-5, atol=1e-6)
        assert abs(float(loss) - float(g[f"vq_loss{s}"])) <= 1e-5 * float(g[f"vq_loss{s}"])
        assert abs(float(perp) - float(g[f"perp{s}"])) <= 1e-5 * float(g[f"perp{s}"])
        state = dict(mod.named_parameters())
        state.update(dict(mod.named_buffers()))
        for n, t in state.items():
            np.testing.assert_allclose(t.detach().cpu().numpy(), g[f"after{s}_" + n.replace(".", "__")], rtol=2e-5, atol=1e-7,
                                       err_msg=f"{n} after step {s}")


def test_vqvae_forward_at_config2_shape_matches_reference():
    """BASELINE configs[1] at its real shape -- VQVAE(WaveNet(**vqwae), c_in=39, hid=64, K=256, encoder_hid=256), 100 MFCC frames
    -> 25 latents -> T = 16000 -- against the reference's own forward (golden): VQ codes BIT-IDENTICAL to the reference's argmin,
    latents / quantised latents, vq_loss, perplexity, fp32 logits <= 1e-3 (north star), bf16 logits within the stated bound."""
    g = load_golden("vqvae_vqwae")
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    cfg = T.CONFIGS["vqwae"]
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=cfg["cin_channels"], K=256, wavenet=WaveNet(**cfg), encoder_hid=256).eval()
    m.load_state_dict(T.synth_state_dict(m, int(g["seed"])))
    m = m.cuda()
    idx = torch.tensor(g["idx"].astype(np.int64)).cuda()
    x = torch.nn.functional.one_hot(idx, cfg["out_channels"]).float().transpose(1, 2).contiguous()
    mfcc, spk = torch.tensor(g["mfcc"]).cuda(), torch.tensor(g["g"]).cuda()
    s = int(g["stride"])
    with torch.no_grad():
        lat = m.encoder(mfcc)
        assert rel_err(lat.cpu().numpy(), g["latents"]) < 1e-5
        y, vq_loss, perp = m(x, mfcc, spk)
        np.testing.assert_array_equal(m.vq.last_codes.cpu().numpy(), g["codes"])          # bit-identical code indices
        quant = m.encode(mfcc)
        assert rel_err(quant.cpu().numpy(), g["quant"]) < 1e-5
        assert rel_err(y[:, :, ::s].cpu().numpy(), g["logits"]) < 1e-4
        assert abs(float(y.double().sum()) - float(g["logits_sum"])) < 1e-4 * abs(float(g["logits_sum"])) + 1.0
        assert abs(vq_loss.item() - float(g["vq_loss"])) < 1e-5 * float(g["vq_loss"])
        assert abs(perp.item() - float(g["perp"])) < 1e-5 * float(g["perp"])
        m.wavenet.precision = "bf16"
        y16, _, _ = m(idx, mfcc, spk)
        np.testing.assert_array_equal(m.vq.last_codes.cpu().numpy(), g["codes"])
        err = rel_err(y16[:, :, ::s].cpu().numpy(), g["logits"])
        print(f"vqvae_vqwae bf16 rel err {err:.3e}")
        assert err < TOL_BF16, err


def test_packed_weights_follow_raw_pointer_optimizer_updates():
    """FlatAdam / a replayed training graph rewrite parameters through raw pointers (no _version bump, same
    data_ptr); the packed-weight caches of WaveNet.forward / incremental_forward and the encoder's transposed weights must
    still notice.  eval -> optimiser step -> eval must change the logits, and equal a freshly built model's on the new weights."""
    from wavenet_autoencoders_b200.train_step import FlatAdam
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    cfg = T.CONFIGS["tiny"]

    def make():
        torch.manual_seed(0)
        mm = VQVAE(c_in=39, hid=cfg["cin_channels"], K=32, wavenet=WaveNet(**cfg), encoder_hid=48).eval()
        mm.load_state_dict(T.synth_state_dict(mm, 5))
        return mm.cuda()
    m = make()
    g0 = load_golden("vqvae_tiny")
    idx, mfcc, spk = torch.tensor(g0["idx"]).cuda(), torch.tensor(g0["mfcc"]).cuda(), torch.tensor(g0["g"]).cuda()
    x = torch.nn.functional.one_hot(idx, cfg["out_channels"]).float().transpose(1, 2).contiguous()
    opt = FlatAdam(m, lr=1e-2, clip=0.0)
    for prec in ("fp32", "bf16"):
        m.wavenet.precision = prec
        with torch.no_grad():
            y0 = m(x, mfcc, spk)[0].clone()
            lat0 = m.encoder(mfcc).clone()
        opt.flat_g.normal_(0, 1.0)                      # a gradient for every parameter; step() writes through raw pointers
        opt.step()
        with torch.no_grad():
            y1 = m(x, mfcc, spk)[0]
            lat1 = m.encoder(mfcc)
        assert not torch.allclose(lat0, lat1) and not torch.allclose(y0, y1), prec
        fresh = make()
        fresh.load_state_dict(m.state_dict())
        fresh.wavenet.precision = prec
        with torch.no_grad():
            assert torch.equal(fresh(x, mfcc, spk)[0], y1), prec


def test_two_stream_backward_equals_one_stream():
    """wae_stack_backward_bf16_2s (weight-gradient GEMMs on a side stream nothing on the caller's stream waits for; the default
    of the fused-loss training step) against the one-stream call: every parameter gradient of a VQ-WAE step at the 20-layer
    vqwae decoder shape, equal up to the fp32 red.add order of the split-K wgrads (run-to-run spread of the one-stream call
    itself is ~1e-8 of the vector; the bound is per tensor)."""
    import os
    from wavenet_autoencoders_b200 import train_step as TS
    from wavenet_autoencoders_b200.vqvae_model import VQVAE
    from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet
    rs = np.random.RandomState(5)
    B, Tn = 2, 1280
    idx = torch.tensor(rs.randint(0, 256, size=(B, Tn)), dtype=torch.long).cuda()
    mfcc = torch.tensor(rs.normal(size=(B, 39, Tn // 160)), dtype=torch.float32).cuda()
    spk = torch.tensor(rs.randint(0, 153, size=(B, 1)), dtype=torch.long).cuda()
    got = {}
    old = os.environ.get("WAE_BWD_STREAMS")
    try:
        for streams in ("1", "2", "2idx"):
            os.environ["WAE_BWD_STREAMS"] = streams[0]
            torch.manual_seed(0)
            m = VQVAE(c_in=39, hid=64, K=256, wavenet=WaveNet(**T.VQWAE), encoder_hid=256)
            m.load_state_dict(T.synth_state_dict(m, 1))
            m = m.cuda().train()
            m.wavenet.precision, m.wavenet.train_impl = "bf16", "kernels"
            opt = TS.FlatAdam(m)
            opt.step = lambda: None                       # keep the gradients
            n0 = _lib.launch_count()
            loss = TS.train_step(m, opt, idx, mfcc, spk, index_input=streams.endswith("idx"))   # one-hot tensor / the classes themselves
            torch.cuda.synchronize()
            assert _lib.launch_count() > n0
            got[streams] = (float(loss), opt.flat_g.clone(), opt.offsets, [p.numel() for p in opt.params],
                            [n for n, p in m.named_parameters() if p.requires_grad])
    finally:
        if old is None:
            os.environ.pop("WAE_BWD_STREAMS", None)
        else:
            os.environ["WAE_BWD_STREAMS"] = old
    # the same step REPLAYED as a CUDA graph (side streams become parallel branches that really run concurrently, and the caching
    # allocator's reuse of blocks is frozen into the graph: the record_stream notes of _stack_backward_tc are what keeps a lagging
    # weight-gradient GEMM's operands from being handed to the upsampler / encoder backward): learning rate 0, so every replay
    # computes the gradient of the same parameters
    torch.manual_seed(0)
    m = VQVAE(c_in=39, hid=64, K=256, wavenet=WaveNet(**T.VQWAE), encoder_hid=256)
    m.load_state_dict(T.synth_state_dict(m, 1))
    m = m.cuda().train()
    m.wavenet.precision, m.wavenet.train_impl = "bf16", "kernels"
    opt = TS.FlatAdam(m, lr=0.0)
    gs = TS.GraphedTrainStep(m, opt, idx, mfcc, spk)
    for _ in range(3):
        lg = gs(idx, mfcc, spk)
    torch.cuda.synchronize()
    got["graph"] = (float(lg), opt.flat_g.clone(), opt.offsets, [p.numel() for p in opt.params], None)
    (l1, a, offs, sizes, names) = got["1"]
    assert float(a.abs().max()) > 0
    for other in ("2", "2idx", "graph"):
        l2, b = got[other][0], got[other][1]
        assert abs(l1 - l2) <= 1e-6 * abs(l1), other
        for n, o, k in zip(names, offs, sizes):
            x, y = a[o:o + k].double(), b[o:o + k].double()
            if float(x.norm()) == 0.0:
                assert float(y.norm()) == 0.0, (other, n)
                continue
            assert float((x - y).norm() / x.norm()) < 1e-4, (other, n)


@pytest.mark.parametrize("scales,B,C,F", [([4, 4], 3, 16, 7), ([4, 4, 8, 5], 2, 64, 12), ([3, 7], 2, 5, 1)])
def test_upsampler_training_kernels_match_torch_autograd(scales, B, C, F):
    """UpsampleNetwork under autograd on the GPU (upsample.UpsampleStageFunction: wae_upsample_stage forward,
    wae_upsample_stage_backward) against the reference's composite (F.interpolate + weight-normed Conv2d, upsample.py:37-49)
    differentiated by torch in true fp32: output, d input, d weight_g / d weight_v of every stage; and the kernel gradient is
    reproducible bit for bit (fixed-order partial sums, no atomics)."""
    from wavenet_autoencoders_b200.wavenet_vocoder.upsample import UpsampleNetwork
    torch.manual_seed(4)
    net = UpsampleNetwork(scales, cin_channels=C).cuda()
    for p in net.parameters():
        p.data.mul_(1.0 + 0.3 * torch.randn_like(p))            # away from the constant initial filter
    x = torch.randn(B, C, F, device="cuda")
    gy = torch.randn(B, C, F * int(np.prod(scales)), device="cuda")
    res = {}
    for impl in ("autograd", "kernels", "kernels"):
        net.train_impl = impl
        net.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        n0 = _lib.launch_count()
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            y = net(xi)
            (y * gy).sum().backward()
        launched = _lib.launch_count() - n0
        assert (launched >= 3 * len(scales)) if impl == "kernels" else (launched == 0)
        cur = (y.detach().clone(), xi.grad.clone(), [p.grad.clone() for p in net.parameters()])
        if impl in res:
            for a, b in zip([res[impl][0], res[impl][1]] + res[impl][2], [cur[0], cur[1]] + cur[2]):
                assert torch.equal(a, b)
        res[impl] = cur
    (y0, dx0, dp0), (y1, dx1, dp1) = res["autograd"], res["kernels"]
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 1e-5
    assert rel_err(dx1.cpu().numpy(), dx0.cpu().numpy()) < 1e-5
    for a, b in zip(dp0, dp1):
        assert float((a - b).abs().max()) <= 1e-4 * max(float(a.abs().max()), 1e-3), (a.flatten(), b.flatten())   # fp32 sums of up to 4e5 terms in different orders


@pytest.mark.parametrize("hid,c_in,c_out,B,F", [(64, 13, 16, 3, 37), (256, 39, 64, 8, 48), (96, 7, 8, 1, 5)])
def test_encoder_training_kernels_match_torch_autograd(hid, c_in, c_out, B, F):
    """Encoder under autograd on the GPU (vqvae_model.EncoderTrainFunction: wae_enc_layer_forward_train / _backward_input /
    _backward_weight) against the same torch modules differentiated by autograd in float64 on the CPU (the reference's
    vqvae_model.py:9-51 arithmetic without any library's algorithm choice in between): latents, d input and the gradient of every
    weight and bias (odd and even frame counts through the stride-2 blocks); the kernel path is reproducible bit for bit.  The
    cuDNN fp32 path of the same modules on the GPU is printed beside it for scale."""
    import copy
    from wavenet_autoencoders_b200.vqvae_model import Encoder
    torch.manual_seed(8)
    enc = Encoder(hid=hid, c_in=c_in, c_out=c_out).cuda().train()
    x = torch.randn(B, c_in, F, device="cuda")
    gy = torch.randn(B, c_out, enc.out_frames(F), device="cuda")
    ref = copy.deepcopy(enc).cpu().double()
    xr = x.cpu().double().requires_grad_(True)
    (ref(xr) * gy.cpu().double()).sum().backward()
    with torch.no_grad():
        want = [ref(xr).detach(), xr.grad] + [p.grad for p in ref.parameters()]
    res = {}
    for impl in ("autograd", "kernels", "kernels"):
        enc.train_impl = impl
        enc.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        n0 = _lib.launch_count()
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            y = enc(xi)
            (y * gy).sum().backward()
        launched = _lib.launch_count() - n0
        assert (launched >= 3 * 11) if impl == "kernels" else (launched == 0), launched
        cur = [y.detach().clone(), xi.grad.clone()] + [p.grad.clone() for p in enc.parameters()]
        if impl in res:
            for a, b in zip(res[impl], cur):
                assert torch.equal(a, b)
        res[impl] = cur
    names = ["latents", "d input"] + [n for n, _ in enc.named_parameters()]
    errs = {}
    for impl in ("kernels", "autograd"):
        for n, a, b in zip(names, want, res[impl]):
            assert a.shape == b.shape, n
            errs[impl, n] = float((a - b.cpu().double()).norm() / max(float(a.norm()), 1e-20))
    print("relative L2 error vs float64, kernels (cuDNN fp32 in brackets):",
          " ".join(f"{n}={errs['kernels', n]:.1e}({errs['autograd', n]:.1e})" for n in names))
    worst = max(((n, errs["kernels", n]) for n in names), key=lambda kv: kv[1])
    assert worst[1] < 5e-6, worst
