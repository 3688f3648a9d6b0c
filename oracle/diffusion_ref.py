"""CPU fp32 oracle for DiffNet + GaussianDiffusion sampling (TEST INFRASTRUCTURE ONLY).

Restates, over plain state dicts:
  DiffNet.forward / ResidualBlock.forward   NeuralSeq/modules/diff/net.py:66-78,107-130
  SinusoidalPosEmb                          net.py:37-44
  Mish                                      NeuralSeq/modules/diff/diffusion.py:68-70
  schedule buffers                          NeuralSeq/modules/diff/shallow_diffusion_tts.py:44-62,82-123
  predict_start_from_noise / q_posterior    shallow_diffusion_tts.py:134-147
  p_sample / p_mean_variance                shallow_diffusion_tts.py:149-166
  p_sample_plms                             shallow_diffusion_tts.py:168-204
  q_sample, norm_spec/denorm_spec           shallow_diffusion_tts.py:206-211,279-283

Pinned by tests/test_oracle_golden.py against tests/golden/diffusion_*.npz
(outputs of the reference modules, tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ schedule
def linear_betas(timesteps, max_beta=0.01):
    # shallow_diffusion_tts.py:44-49 (float64 linspace from 1e-4)
    return np.linspace(1e-4, max_beta, timesteps)


def cosine_betas(timesteps, s=0.008):
    # shallow_diffusion_tts.py:52-62
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    return np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)


def schedule_tables(betas):
    """All tables in float64 numpy, cast to fp32 at the end, exactly as
    shallow_diffusion_tts.py:90-123 does (cast order matters for bit parity)."""
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    t = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(
        betas=t(betas), alphas_cumprod=t(ac), alphas_cumprod_prev=t(ac_prev),
        sqrt_alphas_cumprod=t(np.sqrt(ac)),
        sqrt_one_minus_alphas_cumprod=t(np.sqrt(1.0 - ac)),
        sqrt_recip_alphas_cumprod=t(np.sqrt(1.0 / ac)),
        sqrt_recipm1_alphas_cumprod=t(np.sqrt(1.0 / ac - 1)),
        posterior_variance=t(post_var),
        posterior_log_variance_clipped=t(np.log(np.maximum(post_var, 1e-20))),
        posterior_mean_coef1=t(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        posterior_mean_coef2=t((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    )


# ------------------------------------------------------------------ DiffNet
def step_embedding(t, dim):
    # SinusoidalPosEmb: sin||cos, divisor (half-1)   net.py:37-44
    half = dim // 2
    e = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half) * -e)
    a = t[:, None] * f[None, :]
    return torch.cat((a.sin(), a.cos()), dim=-1)


def mish(x):
    return x * torch.tanh(F.softplus(x))


def diffnet_forward(sd, cfg, spec, t, cond):
    """spec [B,1,M,T], t [B] (int or float), cond [B,H,T] -> eps [B,1,M,T]."""
    C = cfg["residual_channels"]
    L = cfg["residual_layers"]
    cyc = cfg["dilation_cycle_length"]
    with torch.no_grad():
        x = F.relu(F.conv1d(spec[:, 0], sd["input_projection.weight"], sd["input_projection.bias"]))
        e = step_embedding(t.float() if torch.is_tensor(t) else torch.tensor(t).float(), C)
        e = F.linear(mish(F.linear(e, sd["mlp.0.weight"], sd["mlp.0.bias"])),
                     sd["mlp.2.weight"], sd["mlp.2.bias"])
        skip_sum = None
        for i in range(L):
            p = f"residual_layers.{i}"
            d = 2 ** (i % cyc)
            dp = F.linear(e, sd[f"{p}.diffusion_projection.weight"], sd[f"{p}.diffusion_projection.bias"])
            y = x + dp[:, :, None]
            y = F.conv1d(y, sd[f"{p}.dilated_conv.weight"], sd[f"{p}.dilated_conv.bias"],
                         padding=d, dilation=d)
            y = y + F.conv1d(cond, sd[f"{p}.conditioner_projection.weight"],
                             sd[f"{p}.conditioner_projection.bias"])
            gate, filt = y[:, :C], y[:, C:]
            y = torch.sigmoid(gate) * torch.tanh(filt)
            y = F.conv1d(y, sd[f"{p}.output_projection.weight"], sd[f"{p}.output_projection.bias"])
            x = (x + y[:, :C]) / math.sqrt(2.0)
            skip_sum = y[:, C:] if skip_sum is None else skip_sum + y[:, C:]
        x = skip_sum / math.sqrt(L)
        x = F.relu(F.conv1d(x, sd["skip_projection.weight"], sd["skip_projection.bias"]))
        x = F.conv1d(x, sd["output_projection.weight"], sd["output_projection.bias"])
        return x[:, None]


def diffnet_flops_per_frame(cfg):
    C, H, M, L = cfg["residual_channels"], cfg["hidden_size"], cfg["in_dims"], cfg["residual_layers"]
    per_layer = 2 * (C * 2 * C * 3 + H * 2 * C + C * 2 * C)
    return 2 * M * C + L * per_layer + 2 * C * C + 2 * C * M


# ------------------------------------------------------------------ sampling
def _ex(a, t, ndim):
    return a.gather(-1, t).reshape(t.shape[0], *([1] * (ndim - 1)))


def p_sample(tab, eps_fn, x, t, cond, noise, clip_denoised=True):
    """One ancestral step.  ``noise`` is passed in (the reference draws
    torch.randn at shallow_diffusion_tts.py:163)."""
    eps = eps_fn(x, t, cond)
    x0 = _ex(tab["sqrt_recip_alphas_cumprod"], t, x.dim()) * x - \
        _ex(tab["sqrt_recipm1_alphas_cumprod"], t, x.dim()) * eps
    if clip_denoised:
        x0 = x0.clamp(-1.0, 1.0)
    mean = _ex(tab["posterior_mean_coef1"], t, x.dim()) * x0 + \
        _ex(tab["posterior_mean_coef2"], t, x.dim()) * x
    logvar = _ex(tab["posterior_log_variance_clipped"], t, x.dim())
    nz = (1 - (t == 0).float()).reshape(x.shape[0], *([1] * (x.dim() - 1)))
    return mean + nz * (0.5 * logvar).exp() * noise


def plms_x_pred(tab, x, eps, t, interval):
    # get_x_pred, shallow_diffusion_tts.py:174-185 (clamp_min(0) semantics for t-interval)
    ac = tab["alphas_cumprod"]
    a_t = _ex(ac, t, x.dim())
    if int(t[0]) < interval:
        a_prev = torch.ones_like(a_t)
    else:
        a_prev = _ex(ac, torch.clamp(t - interval, min=0), x.dim())
    a_t_sq, a_prev_sq = a_t.sqrt(), a_prev.sqrt()
    delta = (a_prev - a_t) * ((1 / (a_t_sq * (a_t_sq + a_prev_sq))) * x -
                              1 / (a_t_sq * (((1 - a_prev) * a_t).sqrt() + ((1 - a_t) * a_prev).sqrt())) * eps)
    return x + delta


def p_sample_plms(tab, eps_fn, x, t, interval, cond, hist):
    """One PLMS step; ``hist`` is the list of previous eps (newest last, max 4)."""
    eps = eps_fn(x, t, cond)
    n = len(hist)
    if n == 0:
        xp = plms_x_pred(tab, x, eps, t, interval)
        eps_prev = eps_fn(xp, torch.clamp(t - interval, min=0), cond)
        prime = (eps + eps_prev) / 2
    elif n == 1:
        prime = (3 * eps - hist[-1]) / 2
    elif n == 2:
        prime = (23 * eps - 16 * hist[-1] + 5 * hist[-2]) / 12
    else:
        prime = (55 * eps - 59 * hist[-1] + 37 * hist[-2] - 9 * hist[-3]) / 24
    out = plms_x_pred(tab, x, prime, t, interval)
    hist.append(eps)
    if len(hist) > 4:
        del hist[0]
    return out


def q_sample(tab, x0, t, noise):
    return _ex(tab["sqrt_alphas_cumprod"], t, x0.dim()) * x0 + \
        _ex(tab["sqrt_one_minus_alphas_cumprod"], t, x0.dim()) * noise


def norm_spec(x, smin, smax):
    return (x - smin) / (smax - smin) * 2 - 1


def denorm_spec(x, smin, smax):
    return (x + 1) / 2 * (smax - smin) + smin


def sample_loop(sd, cfg, tab, x, cond, noises, t_start=None):
    """Ancestral loop of shallow_diffusion_tts.py:270-271 from t_start-1 down to 0.
    ``noises`` [steps,B,1,M,T] indexed by the step value t."""
    T = tab["betas"].shape[0] if t_start is None else t_start
    fn = lambda xx, tt, cc: diffnet_forward(sd, cfg, xx, tt, cc)
    for i in reversed(range(T)):
        t = torch.full((x.shape[0],), i, dtype=torch.long)
        x = p_sample(tab, fn, x, t, cond, noises[i])
    return x
