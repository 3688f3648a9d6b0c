"""CPU fp32 oracle for the Make-An-Audio UNet + DDIM sampler (TEST INFRASTRUCTURE ONLY).

Restates, over a plain state dict (all paths under
/root/reference/text_to_audio/Make_An_Audio/):
  UNetModel.forward                 ldm/modules/diffusionmodules/openaimodel.py:711-744
  ResBlock._forward                 openaimodel.py:255-275
  Downsample / Upsample             openaimodel.py:134-160, 91-119
  timestep_embedding (cos||sin)     ldm/modules/diffusionmodules/util.py:151-171
  GroupNorm32 (eps 1e-5)            util.py:214-216
  SpatialTransformer.forward        ldm/modules/attention.py:250-261  (Normalize eps 1e-6, :76-77)
  BasicTransformerBlock._forward    attention.py:211-215
  CrossAttention.forward            attention.py:170-193
  GEGLU / FeedForward               attention.py:37-64
  make_beta_schedule('linear')      util.py:21-26
  make_ddim_timesteps / _sampling_parameters   util.py:46-74
  DDIMSampler.make_schedule/sample/ddim_sampling/p_sample_ddim   ldm/models/diffusion/ddim.py:27-225

The block list is re-derived here from the config (independent of
audiogpt_b200.specs.unet_plan) by walking the same constructor rules
(openaimodel.py:516-693).  Pinned by tests/test_oracle_golden.py against
tests/golden/ldm_*.npz produced by the reference classes.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ pieces
def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def resblock(sd, p, x, emb):
    h = F.conv2d(F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)),
                 sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)),
                 sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def attention(sd, p, x, ctx, heads):
    B, N, _ = x.shape
    c = x if ctx is None else ctx
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(c, sd[p + ".to_k.weight"])
    v = F.linear(c, sd[p + ".to_v.weight"])
    d = q.shape[-1] // heads
    sp = lambda t: t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3)
    q, k, v = sp(q), sp(k), sp(v)
    sim = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    o = torch.matmul(sim.softmax(dim=-1), v)
    o = o.permute(0, 2, 1, 3).reshape(B, N, heads * d)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def transformer_block(sd, p, x, ctx, heads):
    x = attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads) + x
    x = attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), ctx, heads) + x
    hcat = F.linear(_ln(sd, p + ".norm3", x), sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"])
    a, g = hcat.chunk(2, dim=-1)
    x = F.linear(a * F.gelu(g), sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"]) + x
    return x


def spatial_transformer(sd, p, x, ctx, heads, depth):
    B, C, H, W = x.shape
    h = _gn(sd, p + ".norm", x, 1e-6)
    h = F.conv2d(h, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    h = h.reshape(B, h.shape[1], H * W).permute(0, 2, 1)
    for d in range(depth):
        h = transformer_block(sd, f"{p}.transformer_blocks.{d}", h, ctx, heads)
    h = h.permute(0, 2, 1).reshape(B, -1, H, W)
    h = F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return h + x


# ------------------------------------------------------------------ UNet
def _layout(cfg):
    mc, mult, nres = cfg["model_channels"], list(cfg["channel_mult"]), cfg["num_res_blocks"]
    ares = set(cfg["attention_resolutions"])

    def nheads(ch):
        if cfg.get("num_head_channels", -1) == -1:
            return cfg["num_heads"]
        return ch // cfg["num_head_channels"]

    ins, stack, ch, ds = [["conv"]], [mc], mc, 1
    for lvl, m in enumerate(mult):
        for _ in range(nres):
            ch = m * mc
            ins.append(["res"] + ([("st", nheads(ch))] if ds in ares else []))
            stack.append(ch)
        if lvl != len(mult) - 1:
            ins.append(["down"])
            stack.append(ch)
            ds *= 2
    mid_heads = nheads(ch)
    outs = []
    for lvl, m in list(enumerate(mult))[::-1]:
        for i in range(nres + 1):
            stack.pop()
            ch = mc * m
            blk = ["res"] + ([("st", nheads(ch))] if ds in ares else [])
            if lvl and i == nres:
                blk.append("up")
                ds //= 2
            outs.append(blk)
    return ins, mid_heads, outs


def unet_forward(sd, cfg, x, t, context):
    """x [N,4,H,W], t [N], context [N,S,ctx] -> eps [N,4,H,W]."""
    depth = cfg.get("transformer_depth", 1)
    ins, mid_heads, outs = _layout(cfg)
    with torch.no_grad():
        emb = timestep_embedding(t, cfg["model_channels"])
        emb = F.linear(F.silu(F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                       sd["time_embed.2.weight"], sd["time_embed.2.bias"])

        def run(prefix, blk, h):
            for j, l in enumerate(blk):
                p = f"{prefix}.{j}"
                if l == "conv":
                    h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
                elif l == "res":
                    h = resblock(sd, p, h, emb)
                elif l == "down":
                    h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
                elif l == "up":
                    h = F.interpolate(h, scale_factor=2, mode="nearest")
                    h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
                else:
                    h = spatial_transformer(sd, p, h, context, l[1], depth)
            return h

        hs, h = [], x
        for i, blk in enumerate(ins):
            h = run(f"input_blocks.{i}", blk, h)
            hs.append(h)
        h = run("middle_block", ["res", ("st", mid_heads), "res"], h)
        for i, blk in enumerate(outs):
            h = torch.cat([h, hs.pop()], dim=1)
            h = run(f"output_blocks.{i}", blk, h)
        h = F.silu(_gn(sd, "out.0", h, 1e-5))
        return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ------------------------------------------------------------------ DDIM
def ldm_schedule(timesteps=1000, linear_start=0.00085, linear_end=0.012):
    """betas / alphas_cumprod / alphas_cumprod_prev as the LDM registers them:
    float64 linspace of sqrt-betas squared (util.py:21-26), cumprod in float64,
    cast to fp32 (ddpm.py register_schedule)."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    ac = np.cumprod(1.0 - betas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(betas=f(betas), alphas_cumprod=f(ac), alphas_cumprod_prev=f(ac_prev))


def ddim_tables(alphas_cumprod, S, eta=0.0, ddpm_steps=None):
    """ddim.py:27-56 with util.py:46-74; alphas_cumprod is the model's fp32 buffer."""
    n = alphas_cumprod.shape[0] if ddpm_steps is None else ddpm_steps
    c = n // S
    steps = np.asarray(list(range(0, n, c))) + 1
    ac = alphas_cumprod.cpu()
    alphas = ac[steps]
    alphas_prev = np.asarray([ac[0]] + ac[steps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return dict(timesteps=steps, alphas=alphas, alphas_prev=alphas_prev,
                sigmas=sigmas, sqrt_one_minus_alphas=np.sqrt(1.0 - alphas))


def ddim_step(x, e_t, a_t, a_prev, sigma_t, sqrt_om, noise=None, temperature=1.0):
    # ddim.py:208-225 with torch.full((b,1,1,1), value) fp32 scalars
    b = x.shape[0]
    full = lambda v: torch.full((b, 1, 1, 1), float(v))
    a_t, a_prev, sigma_t, sqrt_om = full(a_t), full(a_prev), full(sigma_t), full(sqrt_om)
    pred_x0 = (x - sqrt_om * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    nz = sigma_t * (noise if noise is not None else torch.zeros_like(x)) * temperature
    return a_prev.sqrt() * pred_x0 + dir_xt + nz, pred_x0


def ddim_sample(eps_fn, alphas_cumprod, S, x_T, cond, uncond=None, scale=1.0, eta=0.0,
                noises=None, steps_limit=None):
    """eps_fn(x, t, context) -> eps.  Classifier-free guidance as ddim.py:177-198
    (uncond first in the doubled batch)."""
    tab = ddim_tables(alphas_cumprod, S, eta)
    x = x_T
    order = np.flip(tab["timesteps"])
    total = len(order)
    for i, step in enumerate(order):
        if steps_limit is not None and i >= steps_limit:
            break
        idx = total - i - 1
        t = torch.full((x.shape[0],), int(step), dtype=torch.long)
        if uncond is None or scale == 1.0:
            e = eps_fn(x, t, cond)
        else:
            e2 = eps_fn(torch.cat([x, x]), torch.cat([t, t]), torch.cat([uncond, cond]))
            eu, ec = e2.chunk(2)
            e = eu + scale * (ec - eu)
        x, _ = ddim_step(x, e, tab["alphas"][idx], tab["alphas_prev"][idx], tab["sigmas"][idx],
                         tab["sqrt_one_minus_alphas"][idx],
                         None if noises is None else noises[i])
    return x
