"""CPU restatement of PitchExtractor.forward (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

SURVEY.md 8f row 3: the network that extracts F0 from a generated mel for the NSF vocoder.  Follows
NeuralSeq/modules/fastspeech/pe.py:119-148 (PitchExtractor), :7-42 (Prenet: Conv1d k5 -> ReLU -> BatchNorm1d (eval) ->
x non-padding mask, out_proj), :44-116 (ConvBlock / ConvStacks: ConvNorm k5 -> GroupNorm(C/16 groups) -> ReLU, residual),
modules/fastspeech/tts_modules.py:217-260 (PitchPredictor: + alpha * sinusoidal positions, 5 x [conv k SAME -> ReLU ->
LayerNorm over channels], Linear -> 2), modules/commons/common_layers.py:87-142 + utils/__init__.py:145-157 (positions:
cumsum of x[..., 0] != 0), utils/pitch_utils.py:63-76 (denorm_f0).  Functional: state dict in, dict out.
Pinned against the reference module in tests/golden/pe_small.npz / pe_base.npz.
"""
import math

import torch
import torch.nn.functional as F


def _posemb(x0, dim, padding_idx=0):
    mask = x0.ne(padding_idx).int()
    pos = (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    n = int(pos.max().item()) + 1
    emb = torch.arange(n, dtype=torch.float).unsqueeze(1) * f.unsqueeze(0)
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if dim % 2 == 1:
        emb = torch.cat([emb, torch.zeros(n, 1)], dim=1)
    emb[padding_idx, :] = 0
    return emb[pos]


def pe_forward(sd, cfg, mel, use_uv=True, pitch_norm="log", f0_mean=0.0, f0_std=1.0):
    """mel [B, T, n_mel_bins] -> {'pitch_pred': [B, T, 2], 'f0_denorm_pred': [B, T]}"""
    H = int(cfg["hidden_size"])
    pad = mel.abs().sum(-1).eq(0)
    nonpad = 1 - pad.float()[:, None, :]                       # [B, 1, T]
    x = mel.transpose(1, 2)
    for l in range(3):
        p = f"mel_prenet.layers.{l}"
        x = F.conv1d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=2)
        x = F.relu(x)
        x = F.batch_norm(x, sd[p + ".2.running_mean"], sd[p + ".2.running_var"], sd[p + ".2.weight"], sd[p + ".2.bias"],
                         training=False, eps=1e-5)
        x = x * nonpad
    x = F.linear(x.transpose(1, 2), sd["mel_prenet.out_proj.weight"], sd["mel_prenet.out_proj.bias"])
    x = x * nonpad.transpose(1, 2)
    if int(cfg["conv_layers"]) > 0:
        x = F.linear(x, sd["mel_encoder.in_proj.weight"], sd["mel_encoder.in_proj.bias"]).transpose(1, -1)
        for l in range(int(cfg["conv_layers"])):
            p = f"mel_encoder.conv.{l}"
            y = F.conv1d(x, sd[p + ".conv.conv.weight"], sd[p + ".conv.conv.bias"], padding=2)
            y = F.relu(F.group_norm(y, H // 16, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5))
            x = x + y
        x = F.linear(x.transpose(1, -1), sd["mel_encoder.out_proj.weight"], sd["mel_encoder.out_proj.bias"])
    xs = x + sd["pitch_predictor.pos_embed_alpha"] * _posemb(x[..., 0], H)
    xs = xs.transpose(1, -1)
    k = int(cfg["predictor_kernel"])
    for l in range(int(cfg["predictor_layers"])):
        p = f"pitch_predictor.conv.{l}"
        xs = F.conv1d(F.pad(xs, ((k - 1) // 2, (k - 1) // 2)), sd[p + ".1.weight"], sd[p + ".1.bias"])
        xs = F.relu(xs)
        xs = F.layer_norm(xs.transpose(1, -1), (xs.shape[1],), sd[p + ".3.weight"], sd[p + ".3.bias"], eps=1e-5).transpose(1, -1)
    pred = F.linear(xs.transpose(1, -1), sd["pitch_predictor.linear.weight"], sd["pitch_predictor.linear.bias"])
    f0 = pred[:, :, 0].clone()
    if pitch_norm == "standard":
        f0 = f0 * f0_std + f0_mean
    if pitch_norm == "log":
        f0 = 2 ** f0
    if use_uv:
        f0[pred[:, :, 1] > 0] = 0
    f0[pad] = 0
    return {"pitch_pred": pred, "f0_denorm_pred": f0}
