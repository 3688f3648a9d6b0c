"""CPU restatement of AutoencoderKL.decode (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

SURVEY.md 8f row 1 ("next"): the step between the DDIM sampler and the vocoder on every text-to-audio call.
Follows text_to_audio/Make_An_Audio/ldm/models/autoencoder.py:351-354 (decode = post_quant_conv -> decoder),
ldm/modules/diffusionmodules/model.py:538-568 (Decoder.forward), :121-143 (ResnetBlock.forward, temb None),
:177-203 (AttnBlock.forward: single head, scale C^-0.5), :43-49 (Upsample: nearest x2 then conv),
:33-39 (swish, GroupNorm(32, eps 1e-6)).  Functional: state dict in, mel-spectrogram image out.
Pinned against the reference's own Decoder in tests/golden/vae_small.npz / vae_txt2audio.npz.
"""
import torch
import torch.nn.functional as F

from audiogpt_b200.specs import vae_decoder_plan


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _res(x, sd, p, cin, cout):
    h = F.conv2d(_swish(_gn(x, sd, p + ".norm1")), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn(h, sd, p + ".norm2")), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if cin != cout:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def _attn(x, sd, p):
    h = _gn(x, sd, p + ".norm")
    q = F.conv2d(h, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    h = F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + h


def vae_decode(sd, cfg, z):
    """z [B, embed_dim, H, W] -> [B, out_ch, 8H, 8W] (for ch_mult of length 4)."""
    block_in, levels = vae_decoder_plan(cfg)
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _res(h, sd, "decoder.mid.block_1", block_in, block_in)
    h = _attn(h, sd, "decoder.mid.attn_1")
    h = _res(h, sd, "decoder.mid.block_2", block_in, block_in)
    for i_level, blocks, up in levels:
        for j, (cin, cout, has_attn) in enumerate(blocks):
            h = _res(h, sd, f"decoder.up.{i_level}.block.{j}", cin, cout)
            if has_attn:
                h = _attn(h, sd, f"decoder.up.{i_level}.attn.{j}")
        if up:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up.{i_level}.upsample.conv.weight"], sd[f"decoder.up.{i_level}.upsample.conv.bias"],
                         padding=1)
    h = _swish(_gn(h, sd, "decoder.norm_out"))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def vae_decode_flops(cfg, H, W):
    """2 x MAC count of one decode (convs + attention GEMMs) for the measurement row."""
    block_in, levels = vae_decoder_plan(cfg)
    hw = H * W
    fl = 2 * hw * (cfg["z_channels"] * cfg["embed_dim"] + 9 * cfg["z_channels"] * block_in)

    def res(cin, cout, n):
        return 2 * n * (9 * cin * cout + 9 * cout * cout + (cin * cout if cin != cout else 0))

    def attn(c, n):
        return 2 * n * (4 * c * c) + 2 * 2 * n * n * c

    fl += 2 * res(block_in, block_in, hw) + attn(block_in, hw)
    last = block_in
    for _, blocks, up in levels:
        for cin, cout, has_attn in blocks:
            fl += res(cin, cout, hw) + (attn(cout, hw) if has_attn else 0)
            last = cout
        if up:
            hw *= 4
            fl += 2 * hw * 9 * last * last
    fl += 2 * hw * 9 * last * cfg["out_ch"]
    return fl
