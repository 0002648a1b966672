"""Encoder -> VQ -> WaveNet composition with the reference's names (vqvae_model.py:9-84).

Inference: encoder, its final Linear and the VQ search are ONE kernel (wae_encoder_vq_forward, a cluster of 8 CTAs per
utterance with the activations in shared memory); the decoder is the tcgen05 stack.  Under autograd the encoder is
the torch modules (frame rate, < 2 % of the FLOPs).
"""
import os

import torch
from torch import nn

from .vector_quantization import VectorQuantize


class ConvReLURes(nn.Module):
    def __init__(self, dim_in, dim_out, kernel_size, stride=1):
        super().__init__()
        self.stride, self.dim_in, self.dim_out = stride, dim_in, dim_out
        self.conv = nn.Conv1d(dim_in, dim_out, kernel_size, stride, padding=kernel_size // 2, bias=True)
        self.relu = nn.ReLU()

    def forward(self, x):
        out = self.relu(self.conv(x))
        if self.stride == 1 and self.dim_in == self.dim_out:
            out = out + x   # the reference adds in place (vqvae_model.py:20), which modern autograd rejects
        return out


def _enc_splits(B, cin, cout, k, Tout):
    """Split-K factor of one encoder-layer launch: a layer is a few dozen 64 x 64 output tiles with a serial reduction of
    cin*k/32 chunks (~1.9 us each) -- cut the reduction into slices until there are enough blocks for the GPU or a slice is only
    2 chunks long (the latent-rate layers of a training step are 8 tiles: 17-21 us per launch with 4-chunk slices, ncu)."""
    tiles = -(-(B * Tout) // 64) * -(-cout // 64)
    chunks = -(-(cin * k) // 32)
    return max(1, min(chunks // int(os.environ.get("WAE_ENC_MIN_CHUNKS", "1")), -(-148 // tiles), int(os.environ.get("WAE_ENC_MAX_SPLITS", "16"))))


class EncoderTrainFunction(torch.autograd.Function):
    """The encoder (ConvReLURes blocks + Linear, vqvae_model.py:9-51) under autograd on this library's kernels: per layer
    wae_enc_layer_forward_train, then wae_enc_layer_backward_weight / wae_enc_layer_backward_input in the backward -- fp32 FMA
    SGEMMs reading the parameters in place, the ReLU output kept as the mask.  The reference's training step runs these frame-rate
    convolutions through cuDNN (43 us per layer forward at 8 x 48 frames; here ~10)."""

    @staticmethod
    def forward(ctx, enc, x, *params):
        from . import _lib
        lib, st = _lib.lib(), _lib.stream_ptr(x.device)
        specs = enc.layer_specs()
        cur = x.detach().float().contiguous()
        B = cur.shape[0]
        ins, rs = [], []
        for i, (cin, cout, k, s, relu, res) in enumerate(specs):
            w, b = params[2 * i], params[2 * i + 1]
            T = cur.shape[-1]
            Tout = (T - 1) // s + 1
            out = torch.empty(B, cout, Tout, dtype=torch.float32, device=cur.device)
            r = torch.empty_like(out) if relu else None
            splits = _enc_splits(B, cin, cout, k, Tout)
            part = torch.empty(splits * out.numel(), dtype=torch.float32, device=cur.device) if splits > 1 else None
            _lib.check(lib.wae_enc_layer_forward_train(_lib.ptr(cur), _lib.ptr(w.detach()), _lib.ptr(None if b is None else b.detach()), B, cin,
                                                       T, cout, k, s, relu, res, _lib.ptr(out), _lib.ptr(r), splits, _lib.ptr(part), st),
                       "wae_enc_layer_forward_train")
            ins.append(cur)
            rs.append(r)
            cur = out
        ctx.specs, ctx.n = specs, len(specs)
        ctx.has_r = [r is not None for r in rs]
        ctx.save_for_backward(*ins, *[r for r in rs if r is not None], *[p.detach() if p is not None else None for p in params])
        return cur

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        lib, st = _lib.lib(), _lib.stream_ptr(g.device)
        n, saved = ctx.n, ctx.saved_tensors
        ins = saved[:n]
        r_list = list(saved[n:n + sum(ctx.has_r)])
        params = saved[n + sum(ctx.has_r):]
        rs = [r_list.pop(0) if h else None for h in ctx.has_r]
        grads = [None] * (2 * n)
        g = g.contiguous().float()
        B = g.shape[0]
        dx = None
        for i in range(n - 1, -1, -1):
            cin, cout, k, s, relu, res = ctx.specs[i]
            w, b, x, r = params[2 * i], params[2 * i + 1], ins[i], rs[i]
            T = x.shape[-1]
            dw = torch.empty_like(w, dtype=torch.float32)
            db = torch.empty(cout, dtype=torch.float32, device=g.device) if b is not None else None
            part = torch.empty(B * cout * (cin * k + 1), dtype=torch.float32, device=g.device)
            _lib.check(lib.wae_enc_layer_backward_weight(_lib.ptr(g), _lib.ptr(r), _lib.ptr(x), B, cin, T, cout, k, s, _lib.ptr(dw),
                                                         _lib.ptr(db), _lib.ptr(part), st), "wae_enc_layer_backward_weight")
            grads[2 * i], grads[2 * i + 1] = dw, db
            if i == 0 and not ctx.needs_input_grad[1]:
                break
            dx = torch.empty_like(x)
            splits = _enc_splits(B, cout, cin, k, T)
            part = torch.empty(splits * dx.numel(), dtype=torch.float32, device=g.device) if splits > 1 else None
            _lib.check(lib.wae_enc_layer_backward_input(_lib.ptr(g), _lib.ptr(r), _lib.ptr(w), B, cin, T, cout, k, s, res, _lib.ptr(dx),
                                                        splits, _lib.ptr(part), st), "wae_enc_layer_backward_input")
            g = dx
        return (None, dx if ctx.needs_input_grad[1] else None, *grads)


class Encoder(nn.Module):
    def __init__(self, hid=768, c_in=39, c_out=64):
        super().__init__()
        spec = [(c_in, 3, 1), (hid, 3, 1), (hid, 5, 2), (hid, 5, 2), (hid, 3, 1), (hid, 3, 1)] + [(hid, 1, 1)] * 4
        self.net = nn.Sequential(*[ConvReLURes(cin, hid, k, s) for cin, k, s in spec])
        self.lin = nn.Linear(hid, c_out)

    def _forward_kernels(self, x):
        """Inference on CUDA: every ConvReLURes block and the final Linear are one launch of wae_conv1d_relu_res each."""
        from . import _lib
        L, st = _lib.lib(), _lib.stream_ptr(x.device)
        cur = x.detach().float().contiguous()
        B = cur.shape[0]

        def run(inp, w, b, cin, T, cout, k, s, relu, res, out):
            # a layer is a few dozen 64 x 64 output tiles with a serial reduction of Cin*k/32 chunks (~1.9 us each): cut the
            # reduction into slices until there are enough blocks for the GPU or a slice is only 4 chunks long
            tiles = -(-(B * out.shape[-1]) // 64) * -(-cout // 64)
            chunks = -(-(cin * k) // 32)
            splits = max(1, min(chunks // 4, -(-148 // tiles), 8))
            part = None
            if splits > 1:
                need = splits * out.numel()
                part = self.__dict__.get("_partial")
                if part is None or part.numel() < need or part.device != out.device:
                    part = torch.empty(need, dtype=torch.float32, device=out.device)
                    self.__dict__["_partial"] = part
            _lib.check(L.wae_conv1d_relu_res(_lib.ptr(inp), _lib.ptr(w), _lib.ptr(b), B, cin, T, cout, k, s, relu, res, _lib.ptr(out),
                                             splits, _lib.ptr(part), st), "wae_conv1d_relu_res")

        for m in self.net:
            conv = m.conv
            k, s = conv.kernel_size[0], conv.stride[0]
            res = int(m.stride == 1 and m.dim_in == m.dim_out)
            T = cur.shape[-1]
            out = torch.empty(B, m.dim_out, (T - 1) // s + 1, dtype=torch.float32, device=cur.device)
            w, b = self._wt(conv.weight), conv.bias.detach().float().contiguous()
            run(cur, w, b, m.dim_in, T, m.dim_out, k, s, 1, res, out)
            cur = out
        T = cur.shape[-1]
        out = torch.empty(B, self.lin.out_features, T, dtype=torch.float32, device=cur.device)
        w, b = self._wt(self.lin.weight), self.lin.bias.detach().float().contiguous()
        run(cur, w, b, self.lin.in_features, T, self.lin.out_features, 1, 1, 0, 0, out)
        return out

    def layer_specs(self):
        """[(cin, cout, k, stride, relu, residual)] of the ConvReLURes blocks and the final Linear (a k = 1 layer without ReLU)."""
        specs = [(m.dim_in, m.dim_out, m.conv.kernel_size[0], m.conv.stride[0], 1, int(m.stride == 1 and m.dim_in == m.dim_out))
                 for m in self.net]
        return specs + [(self.lin.in_features, self.lin.out_features, 1, 1, 0, 0)]

    def _train_kernels_ok(self, x):
        """Differentiable kernel path (``train_impl`` = "kernels", the default on CUDA; "autograd": torch modules / cuDNN)."""
        return (x.is_cuda and x.dim() == 3 and getattr(self, "train_impl", "kernels") == "kernels" and x.shape[0] <= 65535
                and not getattr(self, "train_tf32", False) and not getattr(self, "train_channels_last", False)
                and all(m.conv.kernel_size[0] % 2 == 1 and m.conv.padding[0] == m.conv.kernel_size[0] // 2 and m.conv.dilation[0] == 1
                        and m.conv.groups == 1 and m.conv.padding_mode == "zeros" for m in self.net))

    def out_frames(self, F):
        """Frames after the stride-2 blocks (F -> (F-1)//2+1 per block)."""
        for m in self.net:
            s = m.conv.stride[0]
            F = (F - 1) // s + 1
        return F

    def fused_struct(self):
        """The layer list as the plain-array struct wae_encoder_vq_forward takes, or None if that kernel does not support this
        encoder (then every block is its own launch, ``_forward_kernels``).  Weights come from the ``_wt`` cache."""
        from . import _lib
        if len(self.net) > 16 or not all(m.conv.kernel_size[0] // 2 == m.conv.padding[0] and m.conv.dilation[0] == 1 and m.conv.groups == 1
                                         for m in self.net):
            return None
        enc = _lib.Encoder()
        keep = []
        enc.n_layers = len(self.net)
        for i, m in enumerate(self.net):
            w = self._wt(m.conv.weight)
            b = None if m.conv.bias is None else m.conv.bias.detach().float().contiguous()
            keep += [w, b]
            ly = enc.layer[i]
            ly.w, ly.bias = w.data_ptr(), (None if b is None else b.data_ptr())
            ly.cin, ly.cout, ly.k, ly.stride = m.dim_in, m.dim_out, m.conv.kernel_size[0], m.conv.stride[0]
            ly.relu, ly.residual = 1, int(m.stride == 1 and m.dim_in == m.dim_out)
        wl = self._wt(self.lin.weight)[:, 0, :]                      # (hid, 1, D) -> (hid, D)
        bl = None if self.lin.bias is None else self.lin.bias.detach().float().contiguous()
        keep += [wl, bl]
        enc.lin_w_t, enc.lin_b = wl.data_ptr(), (None if bl is None else bl.data_ptr())
        enc.hid, enc.D = self.lin.in_features, self.lin.out_features
        if not _lib.lib().wae_encoder_vq_supported(enc):
            return None
        enc._keep = keep
        return enc

    def _wt(self, weight):
        """(Cout, Cin[, k]) -> (Cin, k, Cout) fp32 for the kernel, cached until the parameter is modified in place or replaced."""
        from . import packing
        cache = self.__dict__.setdefault("_wt_cache", {})
        key = id(weight)
        hit = cache.get(key)
        ver = (weight._version, weight.data_ptr(), packing.generation())   # generation: raw-pointer updates (FlatAdam, graph replay)
        if hit is not None and hit[0] == ver and hit[1].device == weight.device and hit[2] is weight:
            return hit[1]
        w3 = weight.detach().float()
        if w3.dim() == 2:
            w3 = w3.unsqueeze(-1)
        wt = w3.permute(1, 2, 0).contiguous()
        cache[key] = (ver, wt, weight)
        return wt

    def _kernels_ok(self, x):
        if not x.is_cuda or (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            return False
        return all(m.conv.kernel_size[0] % 2 == 1 and m.conv.padding[0] == m.conv.kernel_size[0] // 2 and m.conv.dilation[0] == 1
                   and m.conv.groups == 1 for m in self.net)

    def forward(self, x):
        if self._kernels_ok(x):
            return self._forward_kernels(x)
        if self._train_kernels_ok(x) and torch.is_grad_enabled():
            params = []
            for m in self.net:
                params += [m.conv.weight, m.conv.bias]
            params += [self.lin.weight, self.lin.bias]
            return EncoderTrainFunction.apply(self, x, *params)
        # keep the (out-of-scope, cuDNN) encoder in true fp32: TF32 convolutions would move latents by ~1e-3 and
        # flip VQ codes relative to the reference's fp32 path
        # benchmark=True: cuDNN's heuristic pick for these frame-rate shapes (16 x 256 x 100) is an implicit-GEMM kernel that
        # takes 60-90 us per layer; the autotuned choice is several times faster (shapes are static per model)
        # ``train_tf32`` (default off): TF32 tensor-core convolutions for the TRAINING forward / backward of the encoder -- what
        # PyTorch >= 1.7 does by default on Ampere and later; only meaningful next to a bf16 decoder (precision="bf16")
        tf32 = bool(getattr(self, "train_tf32", False)) and self.training and torch.is_grad_enabled()
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32, benchmark=True):
            if getattr(self, "train_channels_last", False) and x.is_cuda and self.training and torch.is_grad_enabled():
                # experiment switch: the same convolutions as (B, C, 1, T) channels-last 2-D convolutions, so that cuDNN's NHWC
                # forward / wgrad kernels need no layout conversion around every layer
                h = x.unsqueeze(2).contiguous(memory_format=torch.channels_last)
                for m in self.net:
                    conv = m.conv
                    y = torch.relu(torch.nn.functional.conv2d(h, conv.weight.unsqueeze(2), conv.bias, stride=(1, conv.stride[0]),
                                                              padding=(0, conv.padding[0])))
                    h = y + h if (m.stride == 1 and m.dim_in == m.dim_out) else y
                out = h.squeeze(2)
            else:
                out = self.net(x)
        return self.lin(out.permute(0, 2, 1)).permute(0, 2, 1)


class VQVAE(nn.Module):
    def __init__(self, c_in=39, hid=64, K=256, wavenet=None, encoder_hid=768):
        super().__init__()
        self.wavenet = wavenet
        self.encoder = Encoder(c_in=c_in, c_out=hid, hid=encoder_hid)
        self.vq = VectorQuantize(K=K, D=hid)
        self.fuse_encoder = True       # inference: encoder + Linear + VQ search as one launch (wae_encoder_vq_forward)

    def _encode_quantize(self, c):
        """(quant, vq_loss, perplexity) of ``vq(encoder(c))`` (vqvae_model.py:66-69).  Inference on CUDA: ONE launch for the
        encoder, its Linear and the nearest-codeword search; otherwise module by module."""
        if (self.fuse_encoder and c.is_cuda and c.dim() == 3 and self.encoder._kernels_ok(c) and hasattr(self.vq, "_fusable")
                and self.vq._fusable(c)):
            enc = self.encoder.fused_struct()
            if enc is not None:
                with torch.no_grad():
                    return self.vq._forward_fused(enc, c, self.encoder.out_frames(c.shape[-1]))
        return self.vq(self.encoder(c))

    def forward(self, x, c, g, softmax=False):
        quant, vq_loss, perp = self._encode_quantize(c)
        return self.wavenet(x, quant, g, softmax), vq_loss, perp

    def forward_nll(self, x, c, g, target, shift=1):
        """Additive: (teacher-forced NLL of the decoder, vq_loss, perplexity) -- the three quantities the reference's training
        step combines (vqwae_train.py:752-766) -- with loss and backward of the decoder fused (WaveNet.forward_nll)."""
        wn = self.wavenet
        if (torch.is_grad_enabled() and x.is_cuda and any(p.requires_grad for p in wn.parameters()) and wn.fused_training_ok(x)):
            # the decoder's weight preparation depends on the parameters only: start it on a side stream now, beside the encoder
            from . import training
            training.begin_weight_prep(wn, x.device)
        try:
            quant, vq_loss, perp = self._encode_quantize(c)
            return wn.forward_nll(x, quant, g, target, shift), vq_loss, perp
        finally:
            wn._prep = None

    def incremental_forward(self, initial_input, c, g, T, softmax, quantize, tqdm, log_scale_min, **extra):
        """vqvae_model.py:74-80.  ``extra``: the additive keywords of ``WaveNet.incremental_forward`` (``uniforms``,
        ``generator``, ``return_indices``), passed through."""
        with torch.no_grad():
            quant, _, _ = self._encode_quantize(c)
            return self.wavenet.incremental_forward(initial_input, c=quant, g=g, T=T, softmax=softmax,
                                                    quantize=quantize, tqdm=tqdm, log_scale_min=log_scale_min, **extra)

    def encode_batch(self, feats):
        """Additive (inference_2019.py:225-262 batched): ``feats`` is a list of (n_frames_i, n_feat) arrays /
        tensors, one per utterance, of ANY lengths; returns the list of (T'_i, D) float32 numpy arrays ``encode`` would give
        utterance by utterance (``rep_tensor.cpu().numpy()[0].transpose()``) -- from ONE launch of the fused encoder + VQ kernel
        over the zero-padded batch with per-utterance lengths."""
        import numpy as np
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            from . import _lib
            raise _lib.WaeError("VQVAE.encode_batch: parameters are not on a CUDA device (no CPU fallback)")
        lens = [int(f.shape[0]) for f in feats]
        Fmax, nfeat = max(lens), int(feats[0].shape[1])
        x = torch.zeros(len(feats), nfeat, Fmax, dtype=torch.float32)
        for i, f in enumerate(feats):
            x[i, :, :lens[i]] = torch.as_tensor(np.asarray(f), dtype=torch.float32).t()
        x = x.to(dev)
        with torch.no_grad():
            enc = self.encoder.fused_struct() if (self.fuse_encoder and hasattr(self.vq, "_fusable") and self.vq._fusable(x)
                                                  and self.encoder._kernels_ok(x)) else None
            if enc is None:      # encoder the fused kernel does not cover: utterance by utterance, as the reference
                return [self.encode(x[i:i + 1, :, :n])[0].t().contiguous().cpu().numpy() for i, n in enumerate(lens)]
            quant, _, _ = self.vq._forward_fused(enc, x, self.encoder.out_frames(Fmax), lengths=torch.tensor(lens, dtype=torch.int32))
            q = quant.permute(0, 2, 1).contiguous().cpu().numpy()
        return [q[i, :self.encoder.out_frames(n)] for i, n in enumerate(lens)]

    @staticmethod
    def dump_representation(path, rep, decimals=6):
        """``np.savetxt(path, rep, fmt='%.6f')`` (inference_2019.py:262), byte-identical, written by the library."""
        import numpy as np
        from . import _lib
        a = np.ascontiguousarray(rep, dtype=np.float32)
        if a.ndim != 2:
            raise ValueError(f"dump_representation: expected a (frames, D) matrix, got shape {a.shape}")
        _lib.check(_lib.lib().wae_dump_text(str(path).encode(), a.ctypes.data, a.shape[0], a.shape[1], int(decimals)), "wae_dump_text")

    def encode(self, x):
        with torch.no_grad():
            quant, _, _ = self._encode_quantize(x)
        return quant
