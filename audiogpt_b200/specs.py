"""State-dict layouts of the hot-path networks + seeded synthetic weights.

Host-side logic only (no arithmetic of the path lives here).  The layouts are
the ones the reference's checkpoints carry, so that the drop-in classes accept
them unchanged:

* HiFi-GAN generator   NeuralSeq/modules/hifigan/hifigan.py:104-142 (ctor)
* DiffNet              NeuralSeq/modules/diff/net.py:58-105
* UNetModel            text_to_audio/Make_An_Audio/ldm/modules/diffusionmodules/openaimodel.py:443-693
  (+ SpatialTransformer ldm/modules/attention.py:152-248)

There is no network in the build/bench environment and the reference ships no
weights (download.sh), so parity tests and the benchmark use *seeded random*
state dicts produced by :func:`synth_state_dict`; zero-initialised layers of
the reference are re-randomised so that parity is not vacuous.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch

# --------------------------------------------------------------------------
# configs named by BASELINE.json / SURVEY.md section 8
# --------------------------------------------------------------------------

HIFIGAN_V1 = dict(  # NeuralSeq/egs/egs_bases/tts/vocoder/hifigan.yaml:3-12
    resblock="1",
    upsample_rates=[8, 8, 2, 2],
    upsample_kernel_sizes=[16, 16, 4, 4],
    upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    use_pitch_embed=False,
    audio_sample_rate=22050,
)

HIFIGAN_SMALL = dict(  # same topology, 8x narrower: CPU-second parity fixture
    resblock="1",
    upsample_rates=[8, 8, 2, 2],
    upsample_kernel_sizes=[16, 16, 4, 4],
    upsample_initial_channel=64,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    use_pitch_embed=False,
    audio_sample_rate=22050,
)

BIGVGAN_BASE = dict(  # BigVGAN-base 22 kHz / 80 band (the args.yml that ships with the Make-An-Audio vocoder
    # checkpoint is download-only; topology of text_to_audio/Make_An_Audio/vocoder/bigvgan/models.py:133-203)
    resblock="1", num_mels=80,
    upsample_rates=[8, 8, 2, 2],
    upsample_kernel_sizes=[16, 16, 4, 4],
    upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    activation="snakebeta", snake_logscale=True,
)

BIGVGAN_SMALL = dict(BIGVGAN_BASE, upsample_initial_channel=64)   # CPU-second parity fixture

VAE_TXT2AUDIO = dict(  # first_stage_config.ddconfig, configs/text_to_audio/txt2audio_args.yaml:52-68
    embed_dim=4, z_channels=4, resolution=848, in_channels=1, out_ch=1, ch=128, ch_mult=[1, 2, 2, 4],
    num_res_blocks=2, attn_resolutions=[106, 212], dropout=0.0, double_z=True,
)

VAE_SMALL = dict(VAE_TXT2AUDIO, ch=32)   # same topology (attention at the same levels), 4x narrower

DIFFNET_BASE = dict(  # egs/egs_bases/svs/base.yaml:3-9, configs/tts/fs2.yaml:5
    in_dims=80, hidden_size=256, residual_layers=20, residual_channels=256,
    dilation_cycle_length=1,
)
DIFFNET_SMALL = dict(in_dims=80, hidden_size=32, residual_layers=4,
                     residual_channels=32, dilation_cycle_length=2)

UNET_TXT2AUDIO = dict(  # configs/text_to_audio/txt2audio_args.yaml:31-50
    in_channels=4, out_channels=4, model_channels=320,
    attention_resolutions=[1, 2], num_res_blocks=2, channel_mult=[1, 2],
    num_heads=8, num_head_channels=-1, use_spatial_transformer=True,
    transformer_depth=1, context_dim=1024, legacy=False,
)
UNET_SMALL = dict(
    in_channels=4, out_channels=4, model_channels=32,
    attention_resolutions=[1, 2], num_res_blocks=2, channel_mult=[1, 2],
    num_heads=4, num_head_channels=-1, use_spatial_transformer=True,
    transformer_depth=1, context_dim=48, legacy=False,
)

# lj_ds_beta6.yaml:5-24 (DiffSpeech mel normalisation range)
SPEC_MIN = [-4.7574, -4.6783, -4.6431, -4.5832, -4.5390, -4.6771, -4.8089, -4.7672,
            -4.5784, -4.7755, -4.7150, -4.8919, -4.8271, -4.7389, -4.6047, -4.7759,
            -4.6799, -4.8201, -4.7823, -4.8262, -4.7857, -4.7545, -4.9358, -4.9733,
            -5.1134, -5.1395, -4.9016, -4.8434, -5.0189, -4.8460, -5.0529, -4.9510,
            -5.0217, -5.0049, -5.1831, -5.1445, -5.1015, -5.0281, -4.9887, -4.9916,
            -4.9785, -4.9071, -4.9488, -5.0342, -4.9332, -5.0650, -4.8924, -5.0875,
            -5.0483, -5.0848, -5.1809, -5.0677, -5.0015, -5.0792, -5.0636, -5.2413,
            -5.1421, -5.1710, -5.3256, -5.0511, -5.1186, -5.0057, -5.0446, -5.1173,
            -5.0325, -5.1085, -5.0053, -5.0755, -5.1176, -5.1004, -5.2153, -5.2757,
            -5.3025, -5.2867, -5.2918, -5.3328, -5.2731, -5.2985, -5.2400, -5.2211]
SPEC_MAX = [-0.5982, -0.0778, 0.1205, 0.2747, 0.4657, 0.5123, 0.5684, 0.7093,
            0.6461, 0.6420, 0.7316, 0.7715, 0.7681, 0.8349, 0.7815, 0.7591,
            0.7910, 0.7433, 0.7352, 0.6869, 0.6854, 0.6623, 0.5353, 0.6492,
            0.6909, 0.6106, 0.5761, 0.5936, 0.5638, 0.4054, 0.4545, 0.3589,
            0.3037, 0.3380, 0.1599, 0.2433, 0.2741, 0.2130, 0.1569, 0.1911,
            0.2324, 0.1586, 0.1221, 0.0341, -0.0558, 0.0553, -0.1153, -0.0933,
            -0.1171, -0.0050, -0.1519, -0.1629, -0.0522, -0.0739, -0.2069, -0.2405,
            -0.1244, -0.2116, -0.1361, -0.1575, -0.1442, 0.0513, -0.1567, -0.2000,
            0.0086, -0.0698, 0.1385, 0.0941, 0.1864, 0.1225, 0.2176, 0.2566,
            0.1670, 0.1007, 0.1444, 0.0888, 0.1998, 0.2414, 0.2932, 0.3047]


# --------------------------------------------------------------------------
# HiFi-GAN
# --------------------------------------------------------------------------

def hifigan_stage_channels(h) -> List[int]:
    c0 = int(h["upsample_initial_channel"])
    return [c0 // (2 ** (i + 1)) for i in range(len(h["upsample_rates"]))]


def hifigan_param_shapes(h, c_out: int = 1, n_mels: int = 80) -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys/shapes after remove_weight_norm() (hifigan.py:171-178)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    c0 = int(h["upsample_initial_channel"])
    rates = list(h["upsample_rates"])
    s["conv_pre.weight"] = (c0, n_mels, 7)
    s["conv_pre.bias"] = (c0,)
    chans = hifigan_stage_channels(h)
    for i, (u, k) in enumerate(zip(rates, h["upsample_kernel_sizes"])):
        s[f"ups.{i}.weight"] = (chans[i] * 2, chans[i], int(k))
        s[f"ups.{i}.bias"] = (chans[i],)
    nk = len(h["resblock_kernel_sizes"])
    for i, ch in enumerate(chans):
        for j, (ks, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}"
            if str(h["resblock"]) == "1":
                for n in range(len(dil)):
                    s[f"{p}.convs1.{n}.weight"] = (ch, ch, int(ks))
                    s[f"{p}.convs1.{n}.bias"] = (ch,)
                for n in range(len(dil)):
                    s[f"{p}.convs2.{n}.weight"] = (ch, ch, int(ks))
                    s[f"{p}.convs2.{n}.bias"] = (ch,)
            else:
                for n in range(len(dil)):
                    s[f"{p}.convs.{n}.weight"] = (ch, ch, int(ks))
                    s[f"{p}.convs.{n}.bias"] = (ch,)
    s["conv_post.weight"] = (c_out, chans[-1], 7)
    s["conv_post.bias"] = (c_out,)
    if h.get("use_pitch_embed"):
        s["m_source.l_linear.weight"] = (1, 9)
        s["m_source.l_linear.bias"] = (1,)
        for i in range(len(rates)):
            if i + 1 < len(rates):
                st = int(math.prod(rates[i + 1:]))
                s[f"noise_convs.{i}.weight"] = (chans[i], 1, 2 * st)
            else:
                s[f"noise_convs.{i}.weight"] = (chans[i], 1, 1)
            s[f"noise_convs.{i}.bias"] = (chans[i],)
    return s


def kaiser_sinc_filter12() -> torch.Tensor:
    """The 12-tap Kaiser-windowed sinc low-pass (cutoff 0.25, half-width 0.3) that both halves of
    Activation1d register as a buffer (vocoder/bigvgan/alias_free_torch/filter.py:19-47, resample.py:17-19,
    38-41): shape [1, 1, 12]."""
    ks, cutoff, half_width = 12, 0.25, 0.3
    half = ks // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    window = torch.kaiser_window(ks, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    f = f / f.sum()
    return f.view(1, 1, ks)


def bigvgan_param_shapes(h) -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys/shapes of BigVGAN.state_dict() after remove_weight_norm() (vocoder/bigvgan/models.py:133-175;
    AMPBlock1 :29-83, AMPBlock2 :86-130; Activation1d buffers alias_free_torch/resample.py:19,41)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    c0 = int(h["upsample_initial_channel"])
    n_mels = int(h.get("num_mels", 80))
    beta = str(h["activation"]) == "snakebeta"
    s["conv_pre.weight"] = (c0, n_mels, 7)
    s["conv_pre.bias"] = (c0,)
    chans = hifigan_stage_channels(h)
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        s[f"ups.{i}.0.weight"] = (chans[i] * 2, chans[i], int(k))
        s[f"ups.{i}.0.bias"] = (chans[i],)

    def act(prefix, ch):
        s[f"{prefix}.act.alpha"] = (ch,)
        if beta:
            s[f"{prefix}.act.beta"] = (ch,)
        s[f"{prefix}.upsample.filter"] = (1, 1, 12)
        s[f"{prefix}.downsample.lowpass.filter"] = (1, 1, 12)

    nk = len(h["resblock_kernel_sizes"])
    for i, ch in enumerate(chans):
        for j, (ks, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}"
            if str(h["resblock"]) == "1":
                for n in range(len(dil)):
                    s[f"{p}.convs1.{n}.weight"] = (ch, ch, int(ks))
                    s[f"{p}.convs1.{n}.bias"] = (ch,)
                for n in range(len(dil)):
                    s[f"{p}.convs2.{n}.weight"] = (ch, ch, int(ks))
                    s[f"{p}.convs2.{n}.bias"] = (ch,)
                for m in range(2 * len(dil)):
                    act(f"{p}.activations.{m}", ch)
            else:
                for n in range(len(dil)):
                    s[f"{p}.convs.{n}.weight"] = (ch, ch, int(ks))
                    s[f"{p}.convs.{n}.bias"] = (ch,)
                for m in range(len(dil)):
                    act(f"{p}.activations.{m}", ch)
    act("activation_post", chans[-1])
    s["conv_post.weight"] = (1, chans[-1], 7)
    s["conv_post.bias"] = (1,)
    return s


def synth_bigvgan(h, seed: int = 4321):
    """Seeded BigVGAN weights: convs as synth_hifigan; snake alpha/beta ~ N(0, 0.3^2) (log scale) or
    1 + 0.3 N (linear scale); the filter buffers hold the Kaiser-sinc taps."""
    shapes = bigvgan_param_shapes(h)
    # gain 0.7: the snake activations do not attenuate like leaky-relu; keeps tanh unsaturated (rms ~0.1)
    sd = synth_state_dict(shapes, seed, gain=0.7, gains={"conv_post.weight": 0.143})
    logscale = bool(h.get("snake_logscale", False))
    filt = kaiser_sinc_filter12()
    for idx, key in enumerate(shapes):
        if key.endswith(".filter"):
            sd[key] = filt.clone()
        elif key.endswith(".act.alpha") or key.endswith(".act.beta"):
            g = torch.Generator(device="cpu")
            g.manual_seed(int(seed) * 7919 + idx)
            t = 0.3 * torch.randn(shapes[key], generator=g, dtype=torch.float32)
            sd[key] = t if logscale else (1.0 + t).abs() + 0.1
    return sd


# --------------------------------------------------------------------------
# DiffNet
# --------------------------------------------------------------------------

def diffnet_param_shapes(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    C, H, M = cfg["residual_channels"], cfg["hidden_size"], cfg["in_dims"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["input_projection.weight"] = (C, M, 1)
    s["input_projection.bias"] = (C,)
    s["mlp.0.weight"] = (4 * C, C)
    s["mlp.0.bias"] = (4 * C,)
    s["mlp.2.weight"] = (C, 4 * C)
    s["mlp.2.bias"] = (C,)
    for i in range(cfg["residual_layers"]):
        p = f"residual_layers.{i}"
        s[f"{p}.dilated_conv.weight"] = (2 * C, C, 3)
        s[f"{p}.dilated_conv.bias"] = (2 * C,)
        s[f"{p}.diffusion_projection.weight"] = (C, C)
        s[f"{p}.diffusion_projection.bias"] = (C,)
        s[f"{p}.conditioner_projection.weight"] = (2 * C, H, 1)
        s[f"{p}.conditioner_projection.bias"] = (2 * C,)
        s[f"{p}.output_projection.weight"] = (2 * C, C, 1)
        s[f"{p}.output_projection.bias"] = (2 * C,)
    s["skip_projection.weight"] = (C, C, 1)
    s["skip_projection.bias"] = (C,)
    s["output_projection.weight"] = (M, C, 1)
    s["output_projection.bias"] = (M,)
    return s


# --------------------------------------------------------------------------
# UNet (openaimodel.UNetModel with SpatialTransformer blocks)
# --------------------------------------------------------------------------

def unet_plan(cfg) -> dict:
    """Walk the constructor logic of openaimodel.py:516-693 and return the block
    list as plain data: each block is a list of layers
    ('conv_in', cin, cout) | ('res', cin, cout) | ('st', ch, heads, dhead) |
    ('down', ch) | ('up', ch).  Only the options the shipped configs use are
    supported (dims=2, conv_resample, no resblock_updown, no class labels,
    use_scale_shift_norm=False)."""
    mc = cfg["model_channels"]
    mult = list(cfg["channel_mult"])
    nres = cfg["num_res_blocks"]
    attn_res = set(cfg["attention_resolutions"])
    num_heads = cfg.get("num_heads", -1)
    nhc = cfg.get("num_head_channels", -1)
    if not cfg.get("use_spatial_transformer", False):
        raise NotImplementedError("only use_spatial_transformer=True UNets are supported")
    depth = cfg.get("transformer_depth", 1)

    def heads_for(ch):
        if nhc == -1:
            return num_heads, ch // num_heads
        return ch // nhc, nhc

    inp: List[list] = [[("conv_in", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nres):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                nh, dh = heads_for(ch)
                layers.append(("st", ch, nh, dh, depth))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    nh, dh = heads_for(ch)
    mid = [("res", ch, ch), ("st", ch, nh, dh, depth), ("res", ch, ch)]
    out: List[list] = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nres + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m)]
            ch = mc * m
            if ds in attn_res:
                nh, dh = heads_for(ch)
                layers.append(("st", ch, nh, dh, depth))
            if level and i == nres:
                layers.append(("up", ch))
                ds //= 2
            out.append(layers)
    return dict(input_blocks=inp, middle_block=mid, output_blocks=out,
                model_channels=mc, time_embed_dim=4 * mc, final_ch=ch,
                context_dim=cfg["context_dim"], out_channels=cfg["out_channels"],
                in_channels=cfg["in_channels"])


def _res_shapes(s, p, cin, cout, temb):
    s[f"{p}.in_layers.0.weight"] = (cin,)
    s[f"{p}.in_layers.0.bias"] = (cin,)
    s[f"{p}.in_layers.2.weight"] = (cout, cin, 3, 3)
    s[f"{p}.in_layers.2.bias"] = (cout,)
    s[f"{p}.emb_layers.1.weight"] = (cout, temb)
    s[f"{p}.emb_layers.1.bias"] = (cout,)
    s[f"{p}.out_layers.0.weight"] = (cout,)
    s[f"{p}.out_layers.0.bias"] = (cout,)
    s[f"{p}.out_layers.3.weight"] = (cout, cout, 3, 3)
    s[f"{p}.out_layers.3.bias"] = (cout,)
    if cin != cout:
        s[f"{p}.skip_connection.weight"] = (cout, cin, 1, 1)
        s[f"{p}.skip_connection.bias"] = (cout,)


def _st_shapes(s, p, ch, nh, dh, depth, ctx):
    inner = nh * dh
    s[f"{p}.norm.weight"] = (ch,)
    s[f"{p}.norm.bias"] = (ch,)
    s[f"{p}.proj_in.weight"] = (inner, ch, 1, 1)
    s[f"{p}.proj_in.bias"] = (inner,)
    for d in range(depth):
        q = f"{p}.transformer_blocks.{d}"
        for name, cdim in (("attn1", inner), ("attn2", ctx)):
            s[f"{q}.{name}.to_q.weight"] = (inner, inner)
            s[f"{q}.{name}.to_k.weight"] = (inner, cdim)
            s[f"{q}.{name}.to_v.weight"] = (inner, cdim)
            s[f"{q}.{name}.to_out.0.weight"] = (inner, inner)
            s[f"{q}.{name}.to_out.0.bias"] = (inner,)
        s[f"{q}.ff.net.0.proj.weight"] = (8 * inner, inner)
        s[f"{q}.ff.net.0.proj.bias"] = (8 * inner,)
        s[f"{q}.ff.net.2.weight"] = (inner, 4 * inner)
        s[f"{q}.ff.net.2.bias"] = (inner,)
        for n in ("norm1", "norm2", "norm3"):
            s[f"{q}.{n}.weight"] = (inner,)
            s[f"{q}.{n}.bias"] = (inner,)
    s[f"{p}.proj_out.weight"] = (ch, inner, 1, 1)
    s[f"{p}.proj_out.bias"] = (ch,)


def unet_param_shapes(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    plan = unet_plan(cfg)
    mc, temb, ctx = plan["model_channels"], plan["time_embed_dim"], plan["context_dim"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["time_embed.0.weight"] = (temb, mc)
    s["time_embed.0.bias"] = (temb,)
    s["time_embed.2.weight"] = (temb, temb)
    s["time_embed.2.bias"] = (temb,)

    def block(prefix, layers):
        for j, l in enumerate(layers):
            p = f"{prefix}.{j}"
            if l[0] == "conv_in":
                s[f"{p}.weight"] = (l[2], l[1], 3, 3)
                s[f"{p}.bias"] = (l[2],)
            elif l[0] == "res":
                _res_shapes(s, p, l[1], l[2], temb)
            elif l[0] == "st":
                _st_shapes(s, p, l[1], l[2], l[3], l[4], ctx)
            elif l[0] == "down":
                s[f"{p}.op.weight"] = (l[1], l[1], 3, 3)
                s[f"{p}.op.bias"] = (l[1],)
            elif l[0] == "up":
                s[f"{p}.conv.weight"] = (l[1], l[1], 3, 3)
                s[f"{p}.conv.bias"] = (l[1],)

    for i, layers in enumerate(plan["input_blocks"]):
        block(f"input_blocks.{i}", layers)
    block("middle_block", plan["middle_block"])
    for i, layers in enumerate(plan["output_blocks"]):
        block(f"output_blocks.{i}", layers)
    s["out.0.weight"] = (plan["final_ch"],)
    s["out.0.bias"] = (plan["final_ch"],)
    s["out.2.weight"] = (plan["out_channels"], mc, 3, 3)
    s["out.2.bias"] = (plan["out_channels"],)
    return s


# --------------------------------------------------------------------------
# seeded synthetic weights
# --------------------------------------------------------------------------

def synth_state_dict(shapes: Dict[str, Sequence[int]], seed: int, gain: float = 1.0,
                     convtranspose_prefixes: Sequence[str] = ("ups.",),
                     gains: Dict[str, float] = None) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 weights (CPU torch.Generator => identical on every box).

    weights ~ N(0, gain^2 / fan_in); biases ~ N(0, 0.05^2); normalisation
    scales ~ 1 + 0.1 N(0,1).  Every tensor is drawn from its own generator
    seeded by (seed, index) so that adding keys does not reshuffle the rest."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for idx, (key, shape) in enumerate(shapes.items()):
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed) * 100003 + idx)
        shape = tuple(int(v) for v in shape)
        if len(shape) == 1:
            t = torch.randn(shape, generator=g, dtype=torch.float32)
            if key.endswith(".weight"):      # GroupNorm / LayerNorm scale
                t = 1.0 + 0.1 * t
            else:
                t = 0.05 * t
        else:
            if any(key.startswith(p) for p in convtranspose_prefixes):
                # ConvTranspose1d weight [C_in, C_out, k], stride u=k/2: two taps per output
                fan_in = shape[0] * 2
            else:
                fan_in = 1
                for v in shape[1:]:
                    fan_in *= v
            gk = gain
            for pref, mul in (gains or {}).items():
                if key.startswith(pref):
                    gk = gain * mul
            t = torch.randn(shape, generator=g, dtype=torch.float32) * (gk / math.sqrt(fan_in))
        out[key] = t
    return out


def synth_tensor(shape: Sequence[int], seed: int, scale: float = 1.0, shift: float = 0.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32) * scale + shift


def synth_hifigan(h, seed: int = 1234, c_out: int = 1):
    """Seeded generator weights; conv_post is scaled down so that tanh is not
    saturated (pre-tanh rms ~0.35) and waveform RMSE is a meaningful metric."""
    return synth_state_dict(hifigan_param_shapes(h, c_out), seed, gains={"conv_post.weight": 0.35})


def vae_decoder_plan(cfg):
    """Block list of ldm Decoder.__init__ (ldm/modules/diffusionmodules/model.py:462-536): returns
    (block_in at the bottom, [(level, [(cin, cout, has_attn), ...], has_upsample), ...] in execution order)."""
    ch, mult = int(cfg["ch"]), list(cfg["ch_mult"])
    nres = len(mult)
    block_in = ch * mult[-1]
    curr_res = int(cfg["resolution"]) // 2 ** (nres - 1)
    levels = []
    bi = block_in
    for i_level in reversed(range(nres)):
        bo = ch * mult[i_level]
        blocks = []
        for _ in range(int(cfg["num_res_blocks"]) + 1):
            blocks.append((bi, bo, curr_res in cfg["attn_resolutions"]))
            bi = bo
        up = i_level != 0
        levels.append((i_level, blocks, up))
        if up:
            curr_res *= 2
    return block_in, levels


def vae_decoder_param_shapes(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    """post_quant_conv (ldm/models/autoencoder.py:307) + decoder.* (model.py:462-536) of AutoencoderKL."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    zc, ed = int(cfg["z_channels"]), int(cfg["embed_dim"])
    s["post_quant_conv.weight"] = (zc, ed, 1, 1)
    s["post_quant_conv.bias"] = (zc,)
    block_in, levels = vae_decoder_plan(cfg)

    def res(p, cin, cout):
        s[f"{p}.norm1.weight"] = (cin,); s[f"{p}.norm1.bias"] = (cin,)
        s[f"{p}.conv1.weight"] = (cout, cin, 3, 3); s[f"{p}.conv1.bias"] = (cout,)
        s[f"{p}.norm2.weight"] = (cout,); s[f"{p}.norm2.bias"] = (cout,)
        s[f"{p}.conv2.weight"] = (cout, cout, 3, 3); s[f"{p}.conv2.bias"] = (cout,)
        if cin != cout:
            s[f"{p}.nin_shortcut.weight"] = (cout, cin, 1, 1); s[f"{p}.nin_shortcut.bias"] = (cout,)

    def attn(p, c):
        s[f"{p}.norm.weight"] = (c,); s[f"{p}.norm.bias"] = (c,)
        for n in ("q", "k", "v", "proj_out"):
            s[f"{p}.{n}.weight"] = (c, c, 1, 1); s[f"{p}.{n}.bias"] = (c,)

    s["decoder.conv_in.weight"] = (block_in, zc, 3, 3)
    s["decoder.conv_in.bias"] = (block_in,)
    res("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    res("decoder.mid.block_2", block_in, block_in)
    last = block_in
    for i_level, blocks, up in levels:
        for j, (cin, cout, has_attn) in enumerate(blocks):
            res(f"decoder.up.{i_level}.block.{j}", cin, cout)
            if has_attn:
                attn(f"decoder.up.{i_level}.attn.{j}", cout)
            last = cout
        if up:
            s[f"decoder.up.{i_level}.upsample.conv.weight"] = (last, last, 3, 3)
            s[f"decoder.up.{i_level}.upsample.conv.bias"] = (last,)
    s["decoder.norm_out.weight"] = (last,); s["decoder.norm_out.bias"] = (last,)
    s["decoder.conv_out.weight"] = (int(cfg["out_ch"]), last, 3, 3)
    s["decoder.conv_out.bias"] = (int(cfg["out_ch"]),)
    return s


def synth_vae_decoder(cfg, seed: int = 5150):
    return synth_state_dict(vae_decoder_param_shapes(cfg), seed, convtranspose_prefixes=())


# ---------------------------------------------------------------------------------------------- PitchExtractor
PE_BASE = dict(n_mel_bins=80, hidden_size=256, conv_layers=2, predictor_hidden=256, predictor_layers=5, predictor_kernel=5)
PE_SMALL = dict(n_mel_bins=80, hidden_size=32, conv_layers=2, predictor_hidden=32, predictor_layers=5, predictor_kernel=5)


def pe_param_shapes(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    """State-dict layout (incl. the BatchNorm buffers) of NeuralSeq/modules/fastspeech/pe.py:119-134 PitchExtractor:
    mel_prenet (Prenet :7-42), mel_encoder (ConvStacks :82-116) and pitch_predictor (tts_modules.py:217-245)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    H, M, P = int(cfg["hidden_size"]), int(cfg["n_mel_bins"]), int(cfg["predictor_hidden"])
    k = int(cfg["predictor_kernel"])
    cin = M
    for l in range(3):
        s[f"mel_prenet.layers.{l}.0.weight"] = (H, cin, 5); s[f"mel_prenet.layers.{l}.0.bias"] = (H,)
        s[f"mel_prenet.layers.{l}.2.weight"] = (H,); s[f"mel_prenet.layers.{l}.2.bias"] = (H,)
        s[f"mel_prenet.layers.{l}.2.running_mean"] = (H,); s[f"mel_prenet.layers.{l}.2.running_var"] = (H,)
        s[f"mel_prenet.layers.{l}.2.num_batches_tracked"] = ()
        cin = H
    s["mel_prenet.out_proj.weight"] = (H, H); s["mel_prenet.out_proj.bias"] = (H,)
    for l in range(int(cfg["conv_layers"])):
        s[f"mel_encoder.conv.{l}.conv.conv.weight"] = (H, H, 5); s[f"mel_encoder.conv.{l}.conv.conv.bias"] = (H,)
        s[f"mel_encoder.conv.{l}.norm.weight"] = (H,); s[f"mel_encoder.conv.{l}.norm.bias"] = (H,)
    if int(cfg["conv_layers"]) > 0:
        s["mel_encoder.in_proj.weight"] = (H, H); s["mel_encoder.in_proj.bias"] = (H,)
        s["mel_encoder.out_proj.weight"] = (H, H); s["mel_encoder.out_proj.bias"] = (H,)
    s["pitch_predictor.pos_embed_alpha"] = (1,)
    cin = H
    for l in range(int(cfg["predictor_layers"])):
        s[f"pitch_predictor.conv.{l}.1.weight"] = (P, cin, k); s[f"pitch_predictor.conv.{l}.1.bias"] = (P,)
        s[f"pitch_predictor.conv.{l}.3.weight"] = (P,); s[f"pitch_predictor.conv.{l}.3.bias"] = (P,)
        cin = P
    s["pitch_predictor.linear.weight"] = (2, P); s["pitch_predictor.linear.bias"] = (2,)
    s["pitch_predictor.embed_positions._float_tensor"] = (1,)
    return s


def synth_pe(cfg, seed: int = 606):
    """Seeded PitchExtractor weights; BatchNorm running statistics are made valid (var > 0)."""
    shapes = pe_param_shapes(cfg)
    sd = synth_state_dict({k: v for k, v in shapes.items() if len(v) > 0}, seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, (k, shp) in enumerate(shapes.items()):
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(0, dtype=torch.long)
        elif k.endswith("running_var"):
            g = torch.Generator().manual_seed(seed * 7919 + i)
            out[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("running_mean"):
            out[k] = sd[k] * 4.0                  # N(0, 0.2^2)
        elif k.endswith("pos_embed_alpha"):
            out[k] = torch.tensor([0.7])
        elif k.endswith("_float_tensor"):
            out[k] = torch.zeros(1)
        else:
            out[k] = sd[k]
    return out


def synth_diffnet(cfg, seed: int = 2024):
    return synth_state_dict(diffnet_param_shapes(cfg), seed)


def synth_unet(cfg, seed: int = 4040):
    return synth_state_dict(unet_param_shapes(cfg), seed)
