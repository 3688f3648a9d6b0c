"""CPU fp32 oracle for the HiFi-GAN generator (TEST INFRASTRUCTURE ONLY).

Functional restatement of ``HifiGanGenerator.forward``
(/root/reference/NeuralSeq/modules/hifigan/hifigan.py:144-169) over a plain
state dict, using torch's own fp32 CPU conv kernels -- the same third-party
arithmetic the reference calls (torch pinned 1.12.1 at requirements.txt:57).

Pinned by tests/test_oracle_golden.py against tests/golden/hifigan_*.npz, which
were produced by the reference module itself (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

STACK_SLOPE = 0.1   # LRELU_SLOPE, hifigan.py:11
FINAL_SLOPE = 0.01  # F.leaky_relu default used before conv_post, hifigan.py:165


def fold_weight_norm(sd):
    """weight = g * v / ||v||  with the norm over all dims but 0
    (torch.nn.utils.weight_norm, dim=0; applied at hifigan.py:33-48,118,124,140).
    For ConvTranspose1d the weight is [C_in, C_out, k] so the norm is per *input*
    channel.  Accepts dicts that are already folded."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[: -len("_g")]
            wv = sd[base + "_v"]
            norm = wv.reshape(wv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (wv.dim() - 1)))
            out[base] = v * wv / norm
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def _pad(k, d):  # get_padding, hifigan.py:26-27
    return (k * d - d) // 2


def resblock1(sd, prefix, x, ks, dils):
    # hifigan.py:54-61
    for n, d in enumerate(dils):
        t = F.leaky_relu(x, STACK_SLOPE)
        t = F.conv1d(t, sd[f"{prefix}.convs1.{n}.weight"], sd[f"{prefix}.convs1.{n}.bias"],
                     padding=_pad(ks, d), dilation=d)
        t = F.leaky_relu(t, STACK_SLOPE)
        t = F.conv1d(t, sd[f"{prefix}.convs2.{n}.weight"], sd[f"{prefix}.convs2.{n}.bias"],
                     padding=_pad(ks, 1))
        x = t + x
    return x


def resblock2(sd, prefix, x, ks, dils):
    # hifigan.py:82-87
    for n, d in enumerate(dils):
        t = F.leaky_relu(x, STACK_SLOPE)
        t = F.conv1d(t, sd[f"{prefix}.convs.{n}.weight"], sd[f"{prefix}.convs.{n}.bias"],
                     padding=_pad(ks, d), dilation=d)
        x = t + x
    return x


def hifigan_forward(sd, h, mel, har_source=None):
    """mel [B,80,T] fp32 -> wav [B,c_out,T*prod(rates)].

    ``har_source`` [B,1,T*prod(rates)] is the merged NSF excitation
    (SourceModuleHnNSF output, hifigan.py:145-149); when given, the strided
    ``noise_convs`` are added after every upsample (hifigan.py:155-157)."""
    sd = fold_weight_norm(sd)
    rates = list(h["upsample_rates"])
    ksz = list(h["upsample_kernel_sizes"])
    rks = list(h["resblock_kernel_sizes"])
    rds = list(h["resblock_dilation_sizes"])
    nk = len(rks)
    block = resblock1 if str(h["resblock"]) == "1" else resblock2
    with torch.no_grad():
        x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
        for i, (u, k) in enumerate(zip(rates, ksz)):
            x = F.leaky_relu(x, STACK_SLOPE)
            x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"],
                                   stride=u, padding=(k - u) // 2)
            if har_source is not None:
                if i + 1 < len(rates):
                    st = int(math.prod(rates[i + 1:]))
                    x = x + F.conv1d(har_source, sd[f"noise_convs.{i}.weight"],
                                     sd[f"noise_convs.{i}.bias"], stride=st, padding=st // 2)
                else:
                    x = x + F.conv1d(har_source, sd[f"noise_convs.{i}.weight"],
                                     sd[f"noise_convs.{i}.bias"])
            acc = None
            for j in range(nk):
                r = block(sd, f"resblocks.{i * nk + j}", x, rks[j], rds[j])
                acc = r if acc is None else acc + r
            x = acc / nk
        x = F.leaky_relu(x, FINAL_SLOPE)
        x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
        return torch.tanh(x)


def nsf_source(lin_w, lin_b, f0_up, sampling_rate, rand_ini, noise, harmonic_num=8, sine_amp=0.1, noise_std=0.003,
               voiced_threshold=0.0):
    """SineGen.forward + SourceModuleHnNSF.forward (NeuralSeq/modules/parallel_wavegan/models/source.py:346-441,
    517-526) with the three random draws made explicit: f0_up [B, L, 1] (already at the sample rate), rand_ini
    [B, dim] (torch.rand; column 0 is zeroed like :357), noise [B, L, dim] (torch.randn_like(sines)).
    Returns sine_merge [B, L, 1]."""
    import numpy as np
    dim = harmonic_num + 1
    f0_buf = torch.zeros(f0_up.shape[0], f0_up.shape[1], dim)
    f0_buf[:, :, 0] = f0_up[:, :, 0]
    for idx in range(harmonic_num):
        f0_buf[:, :, idx + 1] = f0_buf[:, :, 0] * (idx + 2)
    rad = (f0_buf / sampling_rate) % 1
    ri = rand_ini.clone()
    ri[:, 0] = 0
    rad[:, 0, :] = rad[:, 0, :] + ri
    tmp = torch.cumsum(rad, 1) % 1
    over = (tmp[:, 1:, :] - tmp[:, :-1, :]) < 0
    shift = torch.zeros_like(rad)
    shift[:, 1:, :] = over * -1.0
    sines = torch.sin(torch.cumsum(rad + shift, dim=1) * 2 * np.pi) * sine_amp
    uv = torch.ones_like(f0_up) * (f0_up > voiced_threshold)
    noise_amp = uv * noise_std + (1 - uv) * sine_amp / 3
    sines = sines * uv + noise_amp * noise
    return torch.tanh(F.linear(sines, lin_w, lin_b))


def hifigan_flops(h, T, c_out=1, n_mels=80):
    """2*MAC count of one forward at T frames (245.6 GFLOP at V1, T=400)."""
    c = int(h["upsample_initial_channel"])
    fl = 2 * n_mels * c * 7 * T
    L = T
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        co = c // 2
        L = L * u
        fl += 2 * c * co * (k // u) * L          # polyphase: k/u taps per output sample
        for ks, dil in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
            per = 2 * co * co * ks * L
            fl += per * len(dil) * (2 if str(h["resblock"]) == "1" else 1)
        c = co
    fl += 2 * c * c_out * 7 * L
    return fl
