"""Mixture-density helpers with the reference's names (wavenet_vocoder/mixture.py).

Sampling inside autoregressive synthesis is fused into the AR kernel (csrc/wn_ar.cu); the functions
below are the stand-alone torch versions that training scripts import (losses) or that callers may
use on logits they already hold.
"""
import math

import numpy as np
import torch
from torch.nn import functional as F


def log_sum_exp(x):
    m = x.max(dim=-1, keepdim=True)[0]
    return (m + torch.log(torch.exp(x - m).sum(dim=-1, keepdim=True))).squeeze(-1)


def to_one_hot(tensor, n, fill_with=1.0):
    out = torch.zeros(tensor.size() + (n,), dtype=torch.float32, device=tensor.device)
    return out.scatter_(tensor.dim(), tensor.unsqueeze(-1), fill_with)


def _split_params(y_hat, log_scale_min):
    nr_mix = y_hat.size(1) // 3
    y_hat = y_hat.transpose(1, 2)
    return (nr_mix, y_hat[:, :, :nr_mix], y_hat[:, :, nr_mix:2 * nr_mix],
            torch.clamp(y_hat[:, :, 2 * nr_mix:3 * nr_mix], min=log_scale_min))


def discretized_mix_logistic_loss(y_hat, y, num_classes=256, log_scale_min=-7.0, reduce=True):
    """mixture.py:26-106 of the reference: y_hat (B,3*nmix,T), y (B,T,1) in [-1,1]."""
    assert y_hat.dim() == 3 and y_hat.size(1) % 3 == 0
    _, logit_probs, means, log_scales = _split_params(y_hat, log_scale_min)
    y = y.expand_as(means)
    centered = y - means
    inv_stdv = torch.exp(-log_scales)
    half_bin = 1.0 / (num_classes - 1)
    plus_in = inv_stdv * (centered + half_bin)
    min_in = inv_stdv * (centered - half_bin)
    cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    mid_in = inv_stdv * centered
    log_pdf_mid = mid_in - log_scales - 2.0 * F.softplus(mid_in)
    big = (cdf_delta > 1e-5).float()
    inner = big * torch.log(torch.clamp(cdf_delta, min=1e-12)) + (1.0 - big) * (log_pdf_mid - np.log((num_classes - 1) / 2))
    hi = (y > 0.999).float()
    inner = hi * log_one_minus_cdf_min + (1.0 - hi) * inner
    lo = (y < -0.999).float()
    log_probs = lo * log_cdf_plus + (1.0 - lo) * inner + F.log_softmax(logit_probs, -1)
    if reduce:
        return -torch.sum(log_sum_exp(log_probs))
    return -log_sum_exp(log_probs).unsqueeze(-1)


def sample_from_discretized_mix_logistic(y, log_scale_min=-7.0, clamp_log_scale=False):
    """mixture.py:118-156: gumbel-max over mixture logits, then inverse-logistic sampling."""
    assert y.size(1) % 3 == 0
    nr_mix = y.size(1) // 3
    y = y.transpose(1, 2)
    logit_probs = y[:, :, :nr_mix]
    u = torch.empty_like(logit_probs).uniform_(1e-5, 1.0 - 1e-5)
    argmax = (logit_probs.detach() - torch.log(-torch.log(u))).max(dim=-1)[1]
    one_hot = to_one_hot(argmax, nr_mix)
    means = torch.sum(y[:, :, nr_mix:2 * nr_mix] * one_hot, dim=-1)
    log_scales = torch.sum(y[:, :, 2 * nr_mix:3 * nr_mix] * one_hot, dim=-1)
    if clamp_log_scale:
        log_scales = torch.clamp(log_scales, min=log_scale_min)
    u = torch.empty_like(means).uniform_(1e-5, 1.0 - 1e-5)
    x = means + torch.exp(log_scales) * (torch.log(u) - torch.log(1.0 - u))
    return torch.clamp(x, min=-1.0, max=1.0)


def mix_gaussian_loss(y_hat, y, log_scale_min=-7.0, reduce=True):
    """mixture.py:161-218."""
    assert y_hat.dim() == 3
    C = y_hat.size(1)
    nr_mix = 1 if C == 2 else C // 3
    y_t = y_hat.transpose(1, 2)
    if C == 2:
        logit_probs, means, log_scales = None, y_t[:, :, 0:1], torch.clamp(y_t[:, :, 1:2], min=log_scale_min)
    else:
        assert C % 3 == 0
        logit_probs, means = y_t[:, :, :nr_mix], y_t[:, :, nr_mix:2 * nr_mix]
        log_scales = torch.clamp(y_t[:, :, 2 * nr_mix:3 * nr_mix], min=log_scale_min)
    y = y.expand_as(means)
    log_probs = torch.distributions.Normal(loc=0.0, scale=torch.exp(log_scales)).log_prob(y - means)
    if nr_mix > 1:
        log_probs = log_probs + F.log_softmax(logit_probs, -1)
    if nr_mix == 1:
        return -torch.sum(log_probs) if reduce else -log_probs
    return -torch.sum(log_sum_exp(log_probs)) if reduce else -log_sum_exp(log_probs).unsqueeze(-1)


def sample_from_mix_gaussian(y, log_scale_min=-7.0):
    """mixture.py:221-270."""
    C = y.size(1)
    nr_mix = 1 if C == 2 else C // 3
    y = y.transpose(1, 2)
    if nr_mix > 1:
        logit_probs = y[:, :, :nr_mix]
        u = torch.empty_like(logit_probs).uniform_(1e-5, 1.0 - 1e-5)
        argmax = (logit_probs.detach() - torch.log(-torch.log(u))).max(dim=-1)[1]
        one_hot = to_one_hot(argmax, nr_mix)
        means = torch.sum(y[:, :, nr_mix:2 * nr_mix] * one_hot, dim=-1)
        log_scales = torch.sum(y[:, :, 2 * nr_mix:3 * nr_mix] * one_hot, dim=-1)
    elif C == 2:
        means, log_scales = y[:, :, 0], y[:, :, 1]
    else:
        assert C == 3, "shouldn't happen"
        means, log_scales = y[:, :, 1], y[:, :, 2]
    x = torch.distributions.Normal(loc=means, scale=torch.exp(log_scales)).sample()
    return torch.clamp(x, min=-1.0, max=1.0)
