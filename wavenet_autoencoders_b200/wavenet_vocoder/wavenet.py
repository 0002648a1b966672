"""WaveNet decoder with the reference's module API (wavenet_vocoder/wavenet.py), backed by libwae_b200.

Same constructor, sub-module names (hence ``state_dict`` keys), ``forward(x, c, g, softmax)``,
``incremental_forward(...)``, ``clear_buffer``, ``make_generation_fast_``, ``receptive_field``.
What differs is where the arithmetic runs:

* ``forward`` (no autograd)        -> wae_stack_forward_f32 / wae_stack_forward_bf16: one fused kernel per layer
* ``incremental_forward``          -> wae_ar_generate: ONE persistent cluster kernel for all T steps
* ``forward`` when a gradient is needed (grad mode on and any of x, c, g or the parameters requires it, in train OR eval
  mode like the reference) -> ``training.stack_forward_train``: the tcgen05 forward that keeps its activations + the
  hand-derived backward on this library's kernels (``precision="bf16"``, CUDA).  The torch-op composite
  (``_forward_autograd``) is test infrastructure / an fp32 debugging aid and runs only when ``train_impl="autograd"`` is
  set explicitly; with the default ``train_impl="kernels"`` a gradient request on a CPU tensor or with
  ``precision="fp32"`` raises instead of silently taking another path.

``precision`` ("fp32" | "bf16", default from $WAE_B200_PRECISION or "fp32") selects the fp32-faithful
CUDA-core kernels (reference parity to ~1e-5) or the tcgen05 tensor-core kernels.
There is no CPU fallback: inference on a CPU tensor, or without the built library, raises.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn
from torch.nn import functional as F

from .. import _lib, packing
from . import upsample
from .modules import Conv1d1x1, Embedding, ResidualConv1dGLU


def _expand_global_features(B, T, g, bct=True):
    """(B,C) or (B,C,1) -> (B,C,T) / (B,T,C) (wavenet.py:19-39). Only the autograd path needs the expansion."""
    if g is None:
        return None
    g = g.unsqueeze(-1) if g.dim() == 2 else g
    g = g.expand(B, -1, T)
    return g.contiguous() if bct else g.transpose(1, 2).contiguous()


def receptive_field_size(total_layers, num_cycles, kernel_size, dilation=lambda x: 2 ** x):
    assert total_layers % num_cycles == 0
    per = total_layers // num_cycles
    return (kernel_size - 1) * sum(dilation(i % per) for i in range(total_layers)) + 1


class WaveNet(nn.Module):
    def __init__(self, out_channels=256, layers=20, stacks=2, residual_channels=512, gate_channels=512,
                 skip_out_channels=512, kernel_size=3, dropout=1 - 0.95, cin_channels=-1, gin_channels=-1,
                 n_speakers=None, upsample_conditional_features=False, upsample_net="ConvInUpsampleNetwork",
                 upsample_params={"upsample_scales": [4, 4, 4, 4]}, scalar_input=False,
                 use_speaker_embedding=False, output_distribution="Logistic", cin_pad=0):
        super().__init__()
        self.scalar_input = scalar_input
        self.out_channels = out_channels
        self.cin_channels = cin_channels
        self.output_distribution = output_distribution
        assert layers % stacks == 0
        layers_per_stack = layers // stacks
        self.first_conv = Conv1d1x1(1 if scalar_input else out_channels, residual_channels)
        self.conv_layers = nn.ModuleList([
            ResidualConv1dGLU(residual_channels, gate_channels, kernel_size=kernel_size,
                              skip_out_channels=skip_out_channels, bias=True,
                              dilation=2 ** (layer % layers_per_stack), dropout=dropout,
                              cin_channels=cin_channels, gin_channels=gin_channels)
            for layer in range(layers)])
        self.last_conv_layers = nn.ModuleList([
            nn.ReLU(inplace=True), Conv1d1x1(skip_out_channels, skip_out_channels),
            nn.ReLU(inplace=True), Conv1d1x1(skip_out_channels, out_channels)])
        if gin_channels > 0 and use_speaker_embedding:
            assert n_speakers is not None
            self.embed_speakers = Embedding(n_speakers, gin_channels, padding_idx=None, std=0.1)
        else:
            self.embed_speakers = None
        if upsample_conditional_features:
            self.upsample_net = getattr(upsample, upsample_net)(**upsample_params)
        else:
            self.upsample_net = None
        self.receptive_field = receptive_field_size(layers, stacks, kernel_size)

        # ---- B200 execution knobs (not part of the reference API; not in state_dict) ----
        self.precision = os.environ.get("WAE_B200_PRECISION", "fp32")
        self.ar_cluster = None            # None -> 16 CTAs for fp32 weights, 8 for bf16
        self.ar_utts_per_cluster = None   # None -> 8 for the tensor-core AR kernel, 2 for the SIMT kernel
        self.ar_impl = "mma"               # "mma" | "simt" (bf16 precision only)
        self.fuse_frontend = True          # bf16 inference: conv_in + all upsampler stages inside the stack's conditioning kernel
        self.train_impl = "kernels"        # "kernels": tcgen05 forward + GEMM backward (bf16, CUDA) | "autograd": torch ops
        self.last_sampled_indices = None  # (B,T) int32 of the last categorical incremental_forward
        self.last_waveform = None         # (B,T) fp32 of the last incremental_forward(wave_postprocess=...)
        self.last_ar_variant = None       # (weight type, cluster size, utterances per cluster) the last synthesis ran with
        self._packs = {}
        self._ws = packing.WorkspaceCache()

    # ------------------------------------------------------------------ reference API
    def has_speaker_embedding(self):
        return self.embed_speakers is not None

    def local_conditioning_enabled(self):
        return self.cin_channels > 0

    def clear_buffer(self):
        """The reference drops its per-layer shift buffers here (wavenet.py:348-356); the fused AR kernel keeps its
        dilation ring in a scratch workspace that is re-zeroed by every call, so there is nothing to drop."""
        self.first_conv.clear_buffer()
        for f in self.conv_layers:
            f.clear_buffer()
        for f in self.last_conv_layers:
            if hasattr(f, "clear_buffer"):
                f.clear_buffer()

    def make_generation_fast_(self):
        def remove_weight_norm(m):
            try:
                nn.utils.remove_weight_norm(m)
            except ValueError:
                return
        self.apply(remove_weight_norm)
        self._packs.clear()

    # ------------------------------------------------------------------ helpers
    def _speaker_vectors(self, g, B):
        """g: speaker ids (B,)/(B,1) with embed_speakers, else float (B,Gi) / (B,Gi,1)  ->  (B,Gi) fp32 or None."""
        if g is None:
            return None
        if self.embed_speakers is not None:
            g = self.embed_speakers(g.view(B, -1))      # (B,1,Gi)   (wavenet.py:186-191)
            g = g.transpose(1, 2)
        if g.dim() == 3:
            if g.size(-1) != 1:
                raise _lib.WaeError("time-varying global conditioning is not supported (g must be (B,Gi) or (B,Gi,1))")
            g = g[:, :, 0]
        return g

    def _fe_with_speakers(self, fe, g, B):
        """(front-end struct for this call, gvec): integer speaker ids with ``embed_speakers`` are looked up by the kernel that
        builds the per-utterance gate bias (no index_select / transpose launches); anything else goes through
        ``_speaker_vectors``."""
        if g is not None and self.embed_speakers is not None and not torch.is_floating_point(g) and g.is_cuda and g.numel() == B:
            st = type(fe.struct).from_buffer_copy(fe.struct)
            ids = g.detach().reshape(B).long().contiguous()
            tab = self.embed_speakers.weight.detach().float().contiguous()
            st.speaker_ids, st.speaker_table, st.n_speakers = ids.data_ptr(), tab.data_ptr(), tab.shape[0]
            st._keep = (ids, tab, fe)
            return st, None
        return fe.struct, self._speaker_vectors(g, B)

    def _pack(self, kind, **kw):
        key = (kind,) + tuple(sorted(kw.items()))
        fp = packing.params_fingerprint(self)
        hit = self._packs.get(key)
        if hit is None or hit[0] != fp:
            fn = {"f32": packing.pack_f32, "bf16": packing.pack_bf16, "ar": packing.pack_ar, "fe": packing.pack_frontend}[kind]
            with torch.no_grad():
                hit = (fp, fn(self, **kw))
            self._packs[key] = hit
        return hit[1]

    def _require_cuda(self, t, what):
        if not t.is_cuda:
            raise _lib.WaeError(f"{what}: tensor is on {t.device}; wavenet_autoencoders_b200 runs on CUDA sm_100 only "
                                "(no CPU fallback)")

    # ------------------------------------------------------------------ teacher-forced forward
    def _upsample(self, c):
        """The conditioning upsampler; under autograd its stages follow this model's ``train_impl`` ("kernels":
        upsample.UpsampleStageFunction, "autograd": the torch composite of the reference)."""
        inner = getattr(self.upsample_net, "upsample", self.upsample_net)
        if isinstance(inner, upsample.UpsampleNetwork):
            inner.train_impl = self.train_impl
        return self.upsample_net(c)

    def forward(self, x, c=None, g=None, softmax=False):
        """x (B,O,T) one-hot / (B,1,T) scalar; c (B,C,Tc); g ids or (B,Gi[,1])  ->  (B,O,T) (wavenet.py:164-216).
        Additive: x may also be the (B,T) integer mu-law classes themselves (what the one-hot tensor is built from); the
        bf16 inference path then gathers first-conv rows by index and never materialises the (B,O,T) one-hot."""
        # the reference stays differentiable in eval mode too: pick the path by who needs a gradient, not by self.training
        autograd = torch.is_grad_enabled() and (
            (torch.is_floating_point(x) and x.requires_grad) or (c is not None and c.requires_grad)
            or (g is not None and torch.is_floating_point(g) and g.requires_grad)
            or any(p.requires_grad for p in self.parameters()))
        x_idx = None
        if not torch.is_floating_point(x) and x.dim() == 2 and not self.scalar_input:
            if autograd or self.precision != "bf16" or not x.is_cuda:
                x = F.one_hot(x.long(), self.out_channels).float().transpose(1, 2)
            else:
                x_idx = x = x.long().contiguous()     # the kernels read int64 classes
        B, T = x.size(0), x.size(-1)
        last_stage = None
        if (c is not None and self.upsample_net is not None and not autograd and self.precision == "bf16" and x.is_cuda
                and self.fuse_frontend and c.is_cuda and c.dim() == 3):
            fe = self._pack("fe")
            if fe is not None:
                # inference: conv_in + the whole upsampler are evaluated inside the stack's conditioning pass (row f1)
                if c.size(-1) * fe.total_scale != x.size(-1):
                    print(f"c {c.size() } x {x.size()}")
                    raise Exception
                with torch.no_grad():
                    fe_struct, gvec = self._fe_with_speakers(fe, g, B)
                    out = self.stack_forward(x, c, gvec, frontend=fe_struct, x_is_index=x_idx is not None)
                return F.softmax(out, dim=1) if softmax else out
        gvec = self._speaker_vectors(g, B)
        if c is not None and self.upsample_net is not None:
            if not autograd and self.precision == "bf16" and x.is_cuda and isinstance(self.upsample_net, (upsample.UpsampleNetwork, upsample.ConvInUpsampleNetwork)):
                # inference: the last upsampler stage is fused into the stack's conditioning pass
                with torch.no_grad():
                    deferred = self.upsample_net(c, defer_last=True)
                if deferred is not None:
                    c, up_w, up_s = deferred
                    last_stage = (up_w, up_s)
            if last_stage is None:
                c = self._upsample(c)
            if c.size(-1) * (last_stage[1] if last_stage else 1) != x.size(-1):
                print(f"c {c.size() } x {x.size()}")
                raise Exception
        if autograd:
            if self.train_impl == "autograd":        # explicit opt-in only: torch-op composite (tests, fp32 debugging)
                return self._forward_autograd(x, c, gvec, softmax)
            if self.train_impl != "kernels":
                raise ValueError(f"train_impl must be 'kernels' or 'autograd', got {self.train_impl!r}")
            if not x.is_cuda or self.precision != "bf16":
                raise _lib.WaeError(
                    "WaveNet.forward: a gradient is required (grad mode is on and an input or parameter requires grad) but the "
                    f"differentiable kernels need CUDA tensors and precision='bf16' (got {x.device}, precision={self.precision!r}). "
                    "Wrap inference in torch.no_grad(), set model.precision='bf16', or opt into the torch-op composite "
                    "with model.train_impl='autograd' -- there is no silent fallback.")
            if self.training and any(f.dropout > 0 for f in self.conv_layers):
                raise _lib.WaeError(
                    "WaveNet.forward: training with dropout > 0 (modules.py:128) is not implemented in the kernel path "
                    "(every hps/*.json preset uses dropout 0.0); construct the model with dropout=0.0 or set "
                    "model.train_impl='autograd'")
            # tcgen05 forward that keeps its activations + hand-derived backward on them (training.py)
            from .. import training
            out = training.stack_forward_train(self, x, c, gvec)
            return F.softmax(out, dim=1) if softmax else out
        with torch.no_grad():
            out = self.stack_forward(x, c, gvec, last_stage=last_stage, x_is_index=x_idx is not None)
        return F.softmax(out, dim=1) if softmax else out

    def forward_nll(self, x, c=None, g=None, target=None, shift=1):
        """Additive: the reference's training criterion in one call -- mean over b, t < T - shift of the cross-entropy between
        ``forward(x, c, g)[:, :, t]`` and ``target[:, t + shift]`` (vqwae_train.py:760-766 with a mask of ones) -- as a 0-dim
        tensor.  With gradients required (bf16 kernels) loss and backward run fused (training.StackNLLFunction): the
        (B,O,T) log-softmax and its gradient are never materialised.  Without, it is ``losses.teacher_forced_nll``."""
        from .. import losses, training
        needs_grad = torch.is_grad_enabled() and ((c is not None and c.requires_grad) or any(p.requires_grad for p in self.parameters()))
        fused = needs_grad and self.fused_training_ok(x)
        if not fused:
            self._prep = None
            if (not needs_grad and self.precision == "bf16" and x.is_cuda and not torch.is_floating_point(x) and x.dim() == 2
                    and not self.scalar_input):
                return self._nll_from_indices(x, c, g, target, shift)
            y = self.forward(x, c, g)
            if needs_grad:
                return F.cross_entropy(y[:, :, :y.size(-1) - shift], target[:, shift:])
            return losses.teacher_forced_nll(y, target, shift)
        # (B,T) integer classes of a one-hot-input model go to the kernels as they are (first conv = row gather,
        # wae_stack_forward_bf16_save_idx): the (B,256,T) fp32 one-hot of the loader (vqwae_train.py:509-520) never exists
        B = x.size(0)
        gvec = self._speaker_vectors(g, B)
        if c is not None and self.upsample_net is not None:
            c = self._upsample(c)
            if c.size(-1) != x.size(-1):
                print(f"c {c.size() } x {x.size()}")
                raise Exception
        return training.stack_nll_train(self, x, c, gvec, target, shift)

    def fused_training_ok(self, x):
        """True if a gradient-requiring teacher-forced step on ``x`` takes the fused kernel path (training.StackNLLFunction)."""
        from .. import training
        return (self.train_impl == "kernels" and self.precision == "bf16" and x.is_cuda
                and training.tc_backward_supported(packing.stack_shape(self))
                and not (self.training and any(f.dropout > 0 for f in self.conv_layers))
                and (torch.is_floating_point(x) or (x.dim() == 2 and not self.scalar_input)))

    def _nll_from_indices(self, x_idx, c, g, target, shift, logits_out=None):
        """Inference, bf16, class-index input: the NLL comes out of the head kernel's accumulator (wae_stack_nll_bf16_idx);
        the (B,O,T) logits are written only if ``logits_out`` is given."""
        B, T = x_idx.shape
        up_w, up_s = None, 0
        fe = self._pack("fe") if (c is not None and self.upsample_net is not None and self.fuse_frontend and c.dim() == 3) else None
        gvec = None if fe is not None else self._speaker_vectors(g, B)
        with torch.no_grad():
            if fe is not None:
                fe_struct, gvec = self._fe_with_speakers(fe, g, B)
                if c.size(-1) * fe.total_scale != T:
                    print(f"c {c.size() } x {x_idx.size()}")
                    raise Exception
                xi = x_idx.detach().long().contiguous()
                tg = target.detach().long().contiguous()
                lat = c.detach().float().contiguous()
                gv = None if gvec is None else gvec.detach().float().contiguous()
                L, st = _lib.lib(), _lib.stream_ptr(xi.device)
                pk = self._pack("bf16")
                ws = self._ws.get(L.wae_stack_workspace_bf16(pk.struct.d, B, T), xi.device)
                out = torch.zeros(1, dtype=torch.float64, device=xi.device)
                _lib.check(L.wae_stack_forward_bf16_lat(pk.struct, None, _lib.ptr(xi), _lib.ptr(lat), lat.shape[-1], fe_struct, _lib.ptr(gv), B, T,
                                                        _lib.ptr(logits_out), _lib.ptr(tg), int(shift), _lib.ptr(out), _lib.ptr(ws),
                                                        ws.numel(), st), "wae_stack_forward_bf16_lat")
                return (out[0] / float(B * (T - shift))).float()
            if c is not None and self.upsample_net is not None:
                deferred = None
                if isinstance(self.upsample_net, (upsample.UpsampleNetwork, upsample.ConvInUpsampleNetwork)):
                    deferred = self.upsample_net(c, defer_last=True)
                if deferred is not None:
                    c, up_w, up_s = deferred
                else:
                    c = self.upsample_net(c)
                if c.size(-1) * (up_s if up_s else 1) != T:
                    print(f"c {c.size() } x {x_idx.size()}")
                    raise Exception
            xi = x_idx.detach().long().contiguous()
            tg = target.detach().long().contiguous()
            cf = None if c is None else c.detach().float().contiguous()
            gv = None if gvec is None else gvec.detach().float().contiguous()
            uw = None if up_w is None else up_w.detach().float().contiguous()
            L, st = _lib.lib(), _lib.stream_ptr(xi.device)
            pk = self._pack("bf16")
            n = L.wae_stack_workspace_bf16(pk.struct.d, B, T)
            ws = self._ws.get(n, xi.device)
            out = torch.zeros(1, dtype=torch.float64, device=xi.device)
            _lib.check(L.wae_stack_nll_bf16_idx(pk.struct, _lib.ptr(xi), _lib.ptr(cf), 0 if cf is None else cf.shape[-1], int(up_s), _lib.ptr(uw),
                                                _lib.ptr(gv), B, T, _lib.ptr(tg), int(shift), _lib.ptr(out), _lib.ptr(logits_out), _lib.ptr(ws),
                                                ws.numel(), st), "wae_stack_nll_bf16_idx")
            return (out[0] / float(B * (T - shift))).float()

    def stack_forward(self, x, c_up, gvec, precision=None, last_stage=None, x_is_index=False, frontend=None):
        """The hot path proper: first_conv + residual stack + head on already-upsampled conditioning (or, with
        ``last_stage=(filter, scale)``, on the frames entering the last upsampler stage; bf16 only).  ``x_is_index``: x is
        the (B,T) int64 class tensor (bf16 only)."""
        self._require_cuda(x, "WaveNet.forward")
        precision = precision or self.precision
        if (last_stage is not None or x_is_index or frontend is not None) and precision != "bf16":
            raise ValueError("front-end / last_stage fusion / index input exist for precision='bf16' only")
        B, T = x.shape[0], x.shape[-1]
        x = x.detach().contiguous() if x_is_index else x.detach().float().contiguous()
        c_up = None if c_up is None else c_up.detach().float().contiguous()
        gvec = None if gvec is None else gvec.detach().float().contiguous()
        logits = torch.empty(B, self.out_channels, T, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        st = _lib.stream_ptr(x.device)
        if precision == "fp32":
            pk = self._pack("f32")
            n = L.wae_stack_workspace_f32(pk.struct.d, B, T)
            ws = self._ws.get(n, x.device)
            _lib.check(L.wae_stack_forward_f32(pk.struct, _lib.ptr(x), _lib.ptr(c_up), _lib.ptr(gvec), B, T,
                                               _lib.ptr(logits), _lib.ptr(ws), ws.numel(), st), "wae_stack_forward_f32")
        elif precision == "bf16":
            pk = self._pack("bf16")
            n = L.wae_stack_workspace_bf16(pk.struct.d, B, T)
            ws = self._ws.get(n, x.device)
            if frontend is not None:
                # c_up holds the LATENT frames (B, C, F); ``frontend``: the _lib.CondFrontend struct of this call
                _lib.check(L.wae_stack_forward_bf16_lat(pk.struct, None if x_is_index else _lib.ptr(x), _lib.ptr(x) if x_is_index else None,
                                                        _lib.ptr(c_up), c_up.shape[-1], frontend, _lib.ptr(gvec), B, T,
                                                        _lib.ptr(logits), None, 0, None, _lib.ptr(ws), ws.numel(), st),
                           "wae_stack_forward_bf16_lat")
            elif x_is_index:
                up_w, up_s = last_stage if last_stage is not None else (None, 0)
                up_w = None if up_w is None else up_w.detach().float().contiguous()
                _lib.check(L.wae_stack_forward_bf16_idx(pk.struct, _lib.ptr(x), _lib.ptr(c_up), 0 if c_up is None else c_up.shape[-1],
                                                        int(up_s), _lib.ptr(up_w), _lib.ptr(gvec), B, T, _lib.ptr(logits),
                                                        _lib.ptr(ws), ws.numel(), st), "wae_stack_forward_bf16_idx")
            elif last_stage is not None:
                up_w, up_s = last_stage
                up_w = up_w.detach().float().contiguous()
                _lib.check(L.wae_stack_forward_bf16_up(pk.struct, _lib.ptr(x), _lib.ptr(c_up), c_up.shape[-1], int(up_s),
                                                       _lib.ptr(up_w), _lib.ptr(gvec), B, T, _lib.ptr(logits),
                                                       _lib.ptr(ws), ws.numel(), st), "wae_stack_forward_bf16_up")
            else:
                _lib.check(L.wae_stack_forward_bf16(pk.struct, _lib.ptr(x), _lib.ptr(c_up), _lib.ptr(gvec), B, T,
                                                    _lib.ptr(logits), _lib.ptr(ws), ws.numel(), st), "wae_stack_forward_bf16")
        else:
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
        return logits

    def _forward_autograd(self, x, c, gvec, softmax):
        B, _, T = x.size()
        g_bct = _expand_global_features(B, T, gvec, bct=True)
        x = self.first_conv(x)
        skips = 0
        for f in self.conv_layers:
            x, h = f(x, c, g_bct)
            skips = skips + h
        x = skips * math.sqrt(1.0 / len(self.conv_layers))
        for f in self.last_conv_layers:
            x = f(x)
        return F.softmax(x, dim=1) if softmax else x

    # ------------------------------------------------------------------ autoregressive synthesis
    def incremental_forward(self, initial_input=None, c=None, g=None, T=100, test_inputs=None,
                            tqdm=lambda x: x, softmax=True, quantize=True, log_scale_min=-50.0,
                            uniforms=None, generator=None, return_indices=False, wave_postprocess=None):
        """wavenet.py:218-346.  Extra (additive) keywords: ``uniforms`` -- the (T,B[,n]) random draws the fused
        sampler consumes (default: torch.rand with ``generator``); ``return_indices`` -- return the (B,T) int32 sampled
        classes instead of materialising the (B,O,T) one-hot tensor; ``wave_postprocess`` -- a dict with the keywords of
        ``postprocess.waveform_from_synthesis`` (input_type, quantize_channels, postprocess, preemphasis_coef,
        global_gain_scale): the waveform post-processing of synthesis.py:382-394 then runs INSIDE the synthesis kernel, sample
        by sample, and the (B,T) fp32 waveform is left in ``self.last_waveform``."""
        if self.training:
            raise RuntimeError("incremental_forward only supports eval mode")
        self.clear_buffer()
        dev = next(self.parameters()).device
        Oin = 1 if self.scalar_input else self.out_channels
        B = 1
        if test_inputs is not None:
            if test_inputs.size(1) == Oin:          # (B,C,T) -> (B,T,C)   (wavenet.py:249-255)
                test_inputs = test_inputs.transpose(1, 2)
            test_inputs = test_inputs.contiguous()
            B = test_inputs.size(0)
            T = test_inputs.size(1) if T is None else max(T, test_inputs.size(1))
        elif c is not None:
            B = c.shape[0]          # the reference only learns B here, after it already used B=1 for g
        elif initial_input is not None:
            B = initial_input.size(0)
        T = int(T)
        with torch.no_grad():
            gvec = self._speaker_vectors(g, B)
            c_btc = None
            if c is not None:
                if self.upsample_net is not None:
                    c = self.upsample_net(c)
                    assert c.size(-1) == T, f"c {c.size()} != T {T}"
                if c.dim() != 3 or c.size(0) != B:
                    raise ValueError(f"incremental_forward: c must be (B={B}, C, T) or (B, T, C), got {tuple(c.shape)}")
                if c.size(-1) == T:                  # (B,C,T) -> (B,T,C)   (wavenet.py:276-280)
                    c = c.transpose(1, 2)
                if c.size(1) < T:
                    raise ValueError(f"incremental_forward: conditioning covers {c.size(1)} steps but T = {T}")
                # the reference indexes c[:, t] and so tolerates a longer c; the kernel's rows are exactly (B,T,C)
                c_btc = c[:, :T].float().contiguous()
                if self.cin_channels > 0 and c_btc.size(2) != self.cin_channels:
                    raise ValueError(f"incremental_forward: c has {c_btc.size(2)} channels, the model expects {self.cin_channels}")
                self._require_cuda(c_btc, "WaveNet.incremental_forward")
            if initial_input is None:
                init = torch.zeros(B, Oin, device=dev)
                if not self.scalar_input:
                    init[:, 127] = 1
            else:
                init = initial_input
                if init.dim() == 3:
                    if init.size(1) == self.out_channels:   # (B,C,1) -> (B,1,C)   (wavenet.py:292-294)
                        init = init.transpose(1, 2)
                    init = init[:, -1, :]
                init = init.to(dev).float().contiguous()
            if tuple(init.shape) != (B, Oin):
                raise ValueError(f"incremental_forward: initial_input must reduce to (B={B}, {Oin}), got {tuple(init.shape)}")
            if gvec is not None and gvec.shape[0] != B:
                raise ValueError(f"incremental_forward: g has batch {gvec.shape[0]}, expected {B}")
            if dev.type != "cuda":
                raise _lib.WaeError("WaveNet.incremental_forward: parameters are not on a CUDA device (no CPU fallback)")

            if self.scalar_input:
                mode = {"Logistic": _lib.AR_SAMPLE_MOL, "Normal": _lib.AR_SAMPLE_GAUSS}[self.output_distribution]
                nmix = 1 if self.out_channels == 2 else self.out_channels // 3
                if uniforms is None:
                    uniforms = torch.rand(T, B, nmix + 1, device=dev, generator=generator)
                    if mode == _lib.AR_SAMPLE_GAUSS:
                        uniforms[:, :, nmix] = torch.randn(T, B, device=dev, generator=generator)
            elif quantize:
                if not softmax:
                    raise ValueError("quantize=True needs softmax=True (sampling from unnormalised logits is undefined)")
                mode = _lib.AR_SAMPLE_CATEGORICAL
                if uniforms is None:
                    uniforms = torch.rand(T, B, device=dev, generator=generator)
            else:
                mode = _lib.AR_SAMPLE_NONE
            if uniforms is not None:
                want = (T, B, nmix + 1) if self.scalar_input else (T, B)
                if tuple(uniforms.shape) != want:
                    raise ValueError(f"incremental_forward: uniforms must have shape {want}, got {tuple(uniforms.shape)}")
                uniforms = uniforms.to(dev).float().contiguous()

            # fp32 weights -> SIMT kernel (reference parity); bf16 -> tensor-core (mma.sync) kernel with up to 8 utterances
            # per cluster (categorical, mixture-of-logistics and Gaussian samplers alike)
            if self.precision == "fp32":
                wtype = "fp32"
            else:
                wtype = "bf16" if self.ar_impl == "simt" else "bf16mma"
            cluster = self.ar_cluster or (16 if wtype == "fp32" else 8)
            if self.ar_cluster is None and wtype == "fp32" and self.ar_utts_per_cluster in (None, 1, 2):
                # fp32 weights: 16 CTAs per cluster halve every CTA's slice, but only ~7 such clusters are co-resident; more
                # utterance groups than that run in waves (32 utterances: 366 us per step), while 8-CTA clusters (18 co-resident,
                # two utterances each still fit shared memory) take them in one (228 us) -- tools/ar_fp32_sweep.py
                if -(-B // (self.ar_utts_per_cluster or 2)) > 7:
                    cluster = 8
            if wtype == "bf16mma":
                sh = packing.stack_shape(self)     # the tensor-core kernel exchanges bf16 pairs: slice boundaries must be even
                even = lambda cs: not any(packing.part(n, r, cs) % 2 for n in (sh.H, sh.R, sh.S) for r in range(1, cs))
                # clusters of 16 halve every CTA's weight slice (56 vs 60 us per step at the vqwae shape) but only 9 of them
                # are co-resident on 148 SMs, 8 utterances each: take them when the batch fits
                if self.ar_cluster is None and B <= 72 and even(16):
                    cluster = 16
                if not even(cluster):
                    wtype = "bf16"
            upc = self.ar_utts_per_cluster or (8 if wtype == "bf16mma" else 2)
            if wtype != "bf16mma" and upc not in (1, 2, 4):
                upc = 2
            pk = self._pack("ar", cluster=cluster, wtype=wtype, utts_per_cluster=upc)
            if wtype == "bf16mma" and c_btc is not None:
                c_btc = c_btc.to(torch.bfloat16).contiguous()
            L = _lib.lib()
            n = L.wae_ar_workspace(pk.struct, B, T)
            ws = self._ws.get(n, dev)
            forced = None if test_inputs is None else test_inputs.to(dev).float().contiguous()
            if forced is not None and (forced.dim() != 3 or forced.size(0) != B or forced.size(2) != Oin):
                raise ValueError(f"incremental_forward: test_inputs must be (B={B}, T, {Oin}) or (B, {Oin}, T), got {tuple(forced.shape)}")
            Tf = 0 if forced is None else forced.size(1)
            self.last_ar_variant = (wtype, cluster, upc)
            out_idx = torch.empty(B, T, dtype=torch.int32, device=dev) if mode == _lib.AR_SAMPLE_CATEGORICAL else None
            if mode == _lib.AR_SAMPLE_NONE:
                out_dense = torch.empty(B, T, self.out_channels, dtype=torch.float32, device=dev)
            elif mode in (_lib.AR_SAMPLE_MOL, _lib.AR_SAMPLE_GAUSS):
                out_dense = torch.empty(B, T, dtype=torch.float32, device=dev)
            else:
                out_dense = None
            gv_ar = None if gvec is None else gvec.float().contiguous()
            self.last_waveform = None
            if wave_postprocess is not None:
                if mode == _lib.AR_SAMPLE_NONE:
                    raise ValueError("wave_postprocess needs sampling (quantize=True or a scalar-output model)")
                from .. import postprocess as _pp
                post, wave, keep = _pp.ar_post_struct(dev, B, T, categorical=(mode == _lib.AR_SAMPLE_CATEGORICAL), **wave_postprocess)
                _lib.check(L.wae_ar_generate_wave(pk.struct, _lib.ptr(c_btc), _lib.ptr(gv_ar), _lib.ptr(init), _lib.ptr(forced), Tf,
                                                  _lib.ptr(uniforms), B, T, mode, 1 if softmax else 0, _lib.ptr(out_idx),
                                                  _lib.ptr(out_dense), post, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                           "wae_ar_generate_wave")
                self.last_waveform = wave
            else:
                _lib.check(L.wae_ar_generate(pk.struct, _lib.ptr(c_btc), _lib.ptr(gv_ar),
                                             _lib.ptr(init), _lib.ptr(forced), Tf, _lib.ptr(uniforms), B, T, mode,
                                             1 if softmax else 0, _lib.ptr(out_idx), _lib.ptr(out_dense),
                                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "wae_ar_generate")
            if mode == _lib.AR_SAMPLE_CATEGORICAL:
                self.last_sampled_indices = out_idx
                if return_indices:
                    outputs = out_idx
                else:
                    outputs = torch.zeros(B, self.out_channels, T, dtype=torch.float32, device=dev)
                    outputs.scatter_(1, out_idx.long().unsqueeze(1), 1.0)
            elif mode == _lib.AR_SAMPLE_NONE:
                outputs = out_dense.transpose(1, 2).contiguous()
            else:
                outputs = out_dense.unsqueeze(1)
        self.clear_buffer()
        return outputs
