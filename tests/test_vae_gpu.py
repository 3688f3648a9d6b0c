"""AutoencoderKL.decode parity on the GPU (through the C ABI) vs golden vectors made by the reference's own
Decoder + post_quant_conv (tests/golden/make_golden.py:golden_vae) and vs the CPU oracle.  Stated tolerance:
relative RMSE <= 1e-4 on the decoded mel image (37 stacked convs, 7 attention blocks, 3 x fp16-part tensor-core
products with fp32 accumulation; the attention GEMMs are fp32 FMA)."""
import numpy as np
import pytest
import torch

from audiogpt_b200 import specs
from audiogpt_b200.ldm.models.autoencoder import AutoencoderKL
from conftest import load_golden, rel_rmse

pytestmark = pytest.mark.gpu
T = torch.tensor


def build(cfg, seed=5150):
    dd = {k: v for k, v in cfg.items() if k != "embed_dim"}
    m = AutoencoderKL(ddconfig=dd, lossconfig=None, embed_dim=cfg["embed_dim"])
    sd = specs.synth_vae_decoder(cfg, seed)
    # a real checkpoint also carries encoder / quant_conv / loss entries: the drop-in must ignore them (strict=False)
    sd_ckpt = dict(sd, **{"encoder.conv_in.weight": torch.zeros(1), "quant_conv.weight": torch.zeros(1),
                          "loss.logvar": torch.zeros(())})
    missing = m.load_state_dict(sd_ckpt, strict=False)
    assert not missing.missing_keys and set(missing.unexpected_keys) == {"encoder.conv_in.weight", "quant_conv.weight", "loss.logvar"}
    assert set(m.state_dict().keys()) == set(specs.vae_decoder_param_shapes(cfg).keys())
    return m.eval().to("cuda"), sd


def check(cfg, name, z, tol=1e-4):
    g = load_golden(name)
    m, _ = build(cfg)
    y = m.decode(z.cuda()).cpu()
    assert y.shape == (z.shape[0], 1, 8 * z.shape[2], 8 * z.shape[3])
    e = rel_rmse(y[:, :, ::2, ::3], g["mel"])
    yd = y.double()
    st = g["stats"]
    es = abs(float((yd * yd).sum()) - st[2]) / st[2]
    print(f"{name}: decoded mel rel-RMSE {e:.3e}  sum-of-squares rel {es:.3e}")
    assert e < tol and es < 10 * tol
    return m, y


def test_vae_small_vs_reference():
    """same topology as the shipped config (attention at 10x78 and 20x156, strips at 156 / 312 / 624 columns), ch=32"""
    z = specs.synth_tensor((2, 4, 10, 78), seed=3)
    m, y = check(specs.VAE_SMALL, "vae_small", z)
    # batch independence: row 1 alone == row 1 of the batch
    y1 = m.decode(z[1:2].cuda()).cpu()
    assert torch.allclose(y1[0], y[1], atol=1e-5, rtol=1e-4)


def test_vae_txt2audio_vs_reference():
    """the shipped first_stage_config (41 M parameters, 392.9 GFLOP per clip): 4x10x78 latent -> 1x80x624 mel"""
    z = specs.synth_tensor((2, 4, 10, 78), seed=3)[:1]
    check(specs.VAE_TXT2AUDIO, "vae_txt2audio", z)


@pytest.mark.parametrize("B,H,W", [(1, 2, 3), (2, 4, 13), (1, 6, 25)])
def test_vae_ragged_vs_oracle(B, H, W):
    """odd latent sizes: one strip / several ragged strips (W*8 = 24, 104, 200 columns), tiny token counts"""
    from oracle import vae_ref as vr
    cfg = specs.VAE_SMALL
    m, sd = build(cfg)
    z = specs.synth_tensor((B, 4, H, W), seed=100 + W)
    ref = vr.vae_decode(sd, cfg, z)
    got = m.decode(z.cuda()).cpu()
    e = rel_rmse(got, ref)
    print(f"vae small {B}x4x{H}x{W}: rel-RMSE vs oracle {e:.3e}")
    assert e < 1e-4


def test_cpu_tensor_raises():
    m, _ = build(specs.VAE_SMALL)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.decode(torch.zeros(1, 4, 2, 2))
    with pytest.raises(NotImplementedError):
        m.encode(torch.zeros(1, 1, 16, 16).cuda())
