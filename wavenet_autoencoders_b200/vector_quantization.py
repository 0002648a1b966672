"""Drop-in for the reference's ``vector_quantization.py``: same classes, constructor arguments,
parameter/buffer names and ``forward(x:(B,D,T)) -> (quant (B,D,T), vq_loss, perplexity)``.

The nearest-codeword search -- the reference's (N,K) ``addmm`` distance matrix, ``argmin``, (N,K) one-hot
``scatter_`` and one-hot ``@`` codebook gather (vector_quantization.py:27-38, :85-110, :166-221, :267-296) --
is ONE CUDA kernel (wae_vq_search) that emits indices, the straight-through forward value, the squared
error and the code histogram without materialising anything of size N*K.  Loss, perplexity and EMA
bookkeeping stay a few K-sized torch ops around the indices.

Additive API: every module records ``last_codes`` (int64 (B,T) or (B,T,2) for the sliced variants) and
offers ``encode_indices(x)``; the returned 3-tuple is unchanged.

No CPU fallback: a CPU input raises WaeError.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


def _search(x: torch.Tensor, codebook: torch.Tensor, d0: int, sub_d: int, quant: torch.Tensor | None,
            want_stats: bool, sqerr: torch.Tensor | None = None, counts: torch.Tensor | None = None):
    """x (B,D,T) fp32 cuda contiguous -> idx (B*T,) int64 [, sqerr (1,) f64, counts (K,) i32]; fills quant rows."""
    if not x.is_cuda:
        raise _lib.WaeError(f"vector quantization input is on {x.device}; wavenet_autoencoders_b200 runs on CUDA "
                            "sm_100 only (no CPU fallback)")
    B, D, T = x.shape
    K = codebook.shape[0]
    cb = codebook.detach().float().contiguous()
    idx = torch.empty(B * T, dtype=torch.int64, device=x.device)
    if want_stats and sqerr is None:               # else: zeroed views into the caller's per-slice buffers
        sqerr = torch.zeros(1, dtype=torch.float64, device=x.device)
        counts = torch.zeros(K, dtype=torch.int32, device=x.device)
    _lib.check(_lib.lib().wae_vq_search(_lib.ptr(x), B, D, T, d0, sub_d, _lib.ptr(cb), K, _lib.ptr(idx),
                                        _lib.ptr(quant), _lib.ptr(sqerr), _lib.ptr(counts),
                                        _lib.stream_ptr(x.device)), "wae_vq_search")
    return idx, sqerr, counts


def _perplexity(counts: torch.Tensor, n: int) -> torch.Tensor:
    p = counts.float() / float(n)                       # == torch.mean(one_hot, dim=0)
    return torch.exp(-torch.sum(p * torch.log(p + 1e-10)))


def _needs_grad(x, *params):
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))


class _VQBase(nn.Module):
    last_codes = None

    def _slices(self):
        raise NotImplementedError

    def encode_indices(self, x):
        """(B,D,T) -> int64 codes (B,T) (or (B,T,n_slices))."""
        with torch.no_grad():
            xin = x.detach().float().contiguous()
            B, D, T = xin.shape
            cols = [_search(xin, emb.weight, d0, sd, None, False)[0].view(B, T) for d0, sd, emb in self._slices()]
        return cols[0] if len(cols) == 1 else torch.stack(cols, dim=-1)

    def _fusable(self, c):
        """True when this module's forward on the encoder output of ``c`` can run inside wae_encoder_vq_forward: inference only
        (no gradient, no EMA codebook update), one or two slices with one codebook size."""
        slices = self._slices()
        is_ema = hasattr(self, "_ema_update")
        return (not _needs_grad(c, *[emb.weight for _, _, emb in slices]) and not (is_ema and self.training)
                and len(slices) <= 2 and len({emb.weight.shape[0] for _, _, emb in slices}) == 1)

    def _forward_fused(self, enc_struct, c, F4, want_latents=False, lengths=None):
        """Encoder + Linear + search in one launch: same (quant, vq_loss, perplexity) as
        ``self(encoder(c))``; the latents stay on chip unless ``want_latents``."""
        cin = c.detach().float().contiguous()
        B, _, F = cin.shape
        slices = self._slices()
        D = sum(sd for _, sd, _ in slices)
        K = slices[0][2].weight.shape[0]
        dev = cin.device
        alloc = torch.zeros if lengths is not None else torch.empty      # ragged: latents past an utterance's end stay 0
        quant = alloc(B, D, F4, dtype=torch.float32, device=dev)
        lat = alloc(B, D, F4, dtype=torch.float32, device=dev) if want_latents else None
        if lengths is not None:
            lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
            assert lengths.numel() == B
        idx = (torch.full((len(slices), B * F4), -1, dtype=torch.int64, device=dev) if lengths is not None
               else torch.empty(len(slices), B * F4, dtype=torch.int64, device=dev))
        sq_all = torch.zeros(len(slices), dtype=torch.float64, device=dev)
        cn_all = torch.zeros(len(slices), K, dtype=torch.int32, device=dev)
        arr = (_lib.VqSlice * len(slices))()
        keep = []
        for si, (d0, sd, emb) in enumerate(slices):
            cb = emb.weight.detach().float().contiguous()
            keep.append(cb)
            arr[si].codebook, arr[si].K, arr[si].d0, arr[si].sub_d = cb.data_ptr(), K, d0, sd
            arr[si].idx_out = idx[si].data_ptr()
            arr[si].counts_out = cn_all[si].data_ptr()
            arr[si].sqerr_out = sq_all[si:si + 1].data_ptr()
        L, st = _lib.lib(), _lib.stream_ptr(dev)
        from .packing import WorkspaceCache
        if self.__dict__.get("_fused_ws") is None:
            self.__dict__["_fused_ws"] = WorkspaceCache()
        ws = self.__dict__["_fused_ws"].get(L.wae_encoder_vq_workspace(B, F), dev)
        _lib.check(L.wae_encoder_vq_forward(enc_struct, _lib.ptr(cin), _lib.ptr(lengths), B, F, len(slices), arr, _lib.ptr(lat),
                                            _lib.ptr(quant), _lib.ptr(ws), ws.numel(), st), "wae_encoder_vq_forward")
        out2 = torch.empty(2, dtype=torch.float32, device=dev)
        N = B * F4 if lengths is None else int(cn_all[0].sum().item())      # ragged: vectors actually quantised
        _lib.check(L.wae_vq_stats(_lib.ptr(sq_all), _lib.ptr(cn_all), len(slices), K, max(N, 1), max(N, 1) * D, _lib.ptr(out2), st),
                   "wae_vq_stats")
        codes = [idx[si].view(B, F4) for si in range(len(slices))]
        self.last_codes = codes[0] if len(codes) == 1 else torch.stack(codes, dim=-1)
        self.last_latents = lat
        return quant, self._loss(out2[0], out2[0]), out2[1]

    def _forward_common(self, x, loss_fn, training_hook=None):
        xin = x.detach().float().contiguous()
        B, D, T = xin.shape
        N = B * T
        grad = _needs_grad(x, *[emb.weight for _, _, emb in self._slices()])
        quant = torch.empty_like(xin)
        idxs, counts, sqerr_total = [], [], 0.0
        slices = self._slices()
        Ks = [emb.weight.shape[0] for _, _, emb in slices]
        fused_stats = not (grad or training_hook) and len(set(Ks)) == 1     # one statistics launch instead of ~15 tiny ones
        if fused_stats:
            sq_all = torch.zeros(len(slices), dtype=torch.float64, device=xin.device)
            cn_all = torch.zeros(len(slices), Ks[0], dtype=torch.int32, device=xin.device)
        for si, (d0, sd, emb) in enumerate(slices):
            idx, sqerr, cnt = _search(xin, emb.weight, d0, sd, None if (grad or training_hook) else quant, True,
                                      sq_all[si:si + 1] if fused_stats else None, cn_all[si] if fused_stats else None)
            idxs.append(idx)
            counts.append(cnt)
            sqerr_total = sqerr_total + sqerr
        if training_hook is not None:                  # EMA variants rewrite the codebook before the gather
            training_hook(xin, idxs, counts)
        if grad or training_hook is not None:
            # differentiable gather (replaces one_hot @ codebook); forward values identical to the kernel's
            q = torch.cat([emb(idx).view(B, T, sd) for (d0, sd, emb), idx in zip(self._slices(), idxs)], dim=2)
            xt = x.permute(0, 2, 1)
            vq_loss = loss_fn(torch.mean((q.detach() - xt) ** 2), torch.mean((q - xt.detach()) ** 2))
            quant = (xt + (q - xt).detach()).permute(0, 2, 1)
            perp = sum(_perplexity(c, N) for c in counts)
        elif fused_stats:
            out2 = torch.empty(2, dtype=torch.float32, device=xin.device)
            _lib.check(_lib.lib().wae_vq_stats(_lib.ptr(sq_all), _lib.ptr(cn_all), len(slices), Ks[0], N, N * D, _lib.ptr(out2),
                                               _lib.stream_ptr(xin.device)), "wae_vq_stats")
            vq_loss = loss_fn(out2[0], out2[0])
            perp = out2[1]
        else:
            mse = (sqerr_total / float(N * D)).float().squeeze(0)
            vq_loss = loss_fn(mse, mse)
            perp = sum(_perplexity(c, N) for c in counts)
        codes = [i.view(B, T) for i in idxs]
        self.last_codes = codes[0] if len(codes) == 1 else torch.stack(codes, dim=-1)
        return quant, vq_loss, perp


class VectorQuantize(_VQBase):
    """vector_quantization.py:10-49."""

    def __init__(self, K, D, beta=0.25):
        super().__init__()
        self.K, self.D = K, D
        self.embedding = nn.Embedding(K, D)
        self.embedding.weight.data.uniform_(-1.0 / K, 1.0 / K)
        self.beta = beta

    def _slices(self):
        return [(0, self.D, self.embedding)]

    def _loss(self, e, c):
        return self.beta * e + c          # vector_quantization.py:41-43

    def forward(self, inputs):
        quant, loss, perp = self._forward_common(inputs, self._loss)
        return quant.contiguous(), loss, perp


class SlicedVectorQuantize(_VQBase):
    """vector_quantization.py:51-128 (two half-vectors, two codebooks; beta on the commitment term)."""

    def __init__(self, K, D, beta=0.25, decay=0.99, n_d=2, dropout=False, dropout_rate=0.25, K1=None):
        super().__init__()
        self.K = K
        self.K1 = K1 if K1 is not None else K
        self.D = D
        self.sub_D = self.D // n_d
        self.embedding1 = nn.Embedding(K, self.sub_D)
        self.embedding1.weight.data.uniform_(-1.0 / K, 1.0 / K)
        self.embedding2 = nn.Embedding(self.K1, self.sub_D)
        self.embedding2.weight.data.uniform_(-1.0 / self.K1, 1.0 / self.K1)
        self.decay, self.beta = decay, beta
        self.dropout, self.dropout_rate = dropout, dropout_rate

    def _slices(self):
        return [(0, self.sub_D, self.embedding1), (self.sub_D, self.D - self.sub_D, self.embedding2)]

    def _loss(self, e, c):
        return e + self.beta * c          # vector_quantization.py:114-118 (beta on the other term)

    def forward(self, x):
        assert x.size(1) == self.D
        return self._forward_common(x, self._loss)


class _EMAMixin:
    def _ema_update(self, xin, idx, cnt, d0, sd, emb, size_name, w_name):
        """vector_quantization.py:190-217 / :282-294: EMA of cluster sizes (Laplace smoothed) and of the
        per-code input sums, then the codebook is overwritten BEFORE the gather."""
        B, D, T = xin.shape
        K = emb.weight.shape[0]
        dw = torch.zeros(K, sd, dtype=torch.float32, device=xin.device)
        _lib.check(_lib.lib().wae_vq_ema_stats(_lib.ptr(xin), B, D, T, d0, sd, _lib.ptr(idx), K, _lib.ptr(dw),
                                               _lib.stream_ptr(xin.device)), "wae_vq_ema_stats")
        size = getattr(self, size_name) * self.decay + (1.0 - self.decay) * cnt.float()
        n = torch.sum(size)
        size = (size + 1e-5) / (n + K * 1e-5) * n
        w = getattr(self, w_name) * self.decay + (1 - self.decay) * dw
        setattr(self, size_name, size)
        setattr(self, w_name, w)
        emb.weight.data.copy_(w / size.unsqueeze(1))


class SlicedVectorQuantizeEMA(_VQBase, _EMAMixin):
    """vector_quantization.py:132-235."""

    def __init__(self, K, D, beta=0.25, decay=0.99, n_d=2):
        super().__init__()
        self.K, self.D = K, D
        self.sub_D = self.D // n_d
        self.embedding1 = nn.Embedding(K, self.sub_D)
        self.embedding1.weight.data.uniform_(-1.0 / K, 1.0 / K)
        self.embedding2 = nn.Embedding(K, self.sub_D)
        self.embedding2.weight.data.uniform_(-1.0 / K, 1.0 / K)
        if self.training:
            self.register_buffer("ema_cluster_size1", torch.zeros(K))
            self.register_buffer("ema_w1", torch.zeros(K, self.sub_D))
            self.register_buffer("ema_cluster_size2", torch.zeros(K))
            self.register_buffer("ema_w2", torch.zeros(K, self.sub_D))
        self.decay, self.beta = decay, beta

    def _slices(self):
        return [(0, self.sub_D, self.embedding1), (self.sub_D, self.D - self.sub_D, self.embedding2)]

    def forward(self, x):
        assert x.size(1) == self.D

        def hook(xin, idxs, counts):
            with torch.no_grad():
                for n, ((d0, sd, emb), idx, cnt) in enumerate(zip(self._slices(), idxs, counts), start=1):
                    self._ema_update(xin, idx, cnt, d0, sd, emb, f"ema_cluster_size{n}", f"ema_w{n}")

        return self._forward_common(x, self._loss, hook if self.training else None)

    def _loss(self, e, c):
        return self.beta * e              # vector_quantization.py:224 (commitment term only)


class VectorQuantizeEMA(_VQBase, _EMAMixin):
    """vector_quantization.py:239-306."""

    def __init__(self, K, D, beta=0.25, decay=0.99):
        super().__init__()
        self.K, self.D = K, D
        self.embedding = nn.Embedding(K, D)
        self.embedding.weight.data.uniform_(-1.0 / K, 1.0 / K)
        if self.training:
            self.register_buffer("ema_cluster_size", torch.zeros(K))
            self.register_buffer("ema_w", torch.zeros(K, D))
        self.decay, self.beta = decay, beta

    def _slices(self):
        return [(0, self.D, self.embedding)]

    def forward(self, x):
        assert x.size(1) == self.D

        def hook(xin, idxs, counts):
            with torch.no_grad():
                self._ema_update(xin, idxs[0], counts[0], 0, self.D, self.embedding, "ema_cluster_size", "ema_w")

        return self._forward_common(x, self._loss, hook if self.training else None)

    def _loss(self, e, c):
        return self.beta * e              # vector_quantization.py:298
