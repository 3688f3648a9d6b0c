"""GPU parity of the BigVGAN drop-in against the reference-made fixture and the oracle."""
import numpy as np
import pytest
import torch

from audiogpt_b200 import specs
from conftest import load_golden, rmse

pytestmark = pytest.mark.gpu


def _model(h, seed=4321):
    from audiogpt_b200.vocoder.bigvgan.models import BigVGAN
    m = BigVGAN(h)
    m.load_state_dict(specs.synth_bigvgan(h, seed), strict=True)
    return m.eval().cuda()


def test_bigvgan_small_vs_reference_golden():
    g = load_golden("bigvgan_small")
    h = specs.BIGVGAN_SMALL
    m = _model(h)
    mel = specs.synth_tensor((2, 80, 20), seed=5, scale=2.0, shift=-4.0).cuda()
    wav = m(mel)
    assert wav.shape == (2, 1, 20 * 256)
    e = rmse(wav.cpu().numpy(), g["wav"])
    print("bigvgan small RMSE vs reference golden:", e)
    assert e < 2e-5                                  # waveform rms ~0.1; north-star bound 1e-4
    w7 = m(mel[:1, :, :7].contiguous())              # shorter than the replicate pads of the activations
    e7 = rmse(w7.cpu().numpy(), g["wav_t7"])
    print("bigvgan small T=7 RMSE:", e7)
    assert e7 < 2e-5


def test_bigvgan_vs_oracle_ragged_and_vocode_wrapper():
    from oracle import bigvgan_ref as br
    from audiogpt_b200.vocoder.bigvgan.models import VocoderBigVGAN
    h = specs.BIGVGAN_SMALL
    sd = specs.synth_bigvgan(h, 777)
    m = _model(h, 777)
    for B, T in ((3, 33), (1, 1), (2, 129)):
        mel = specs.synth_tensor((B, 80, T), seed=100 + T, scale=2.0, shift=-4.0)
        ref = br.bigvgan_forward(sd, h, mel)
        wav = m(mel.cuda())
        assert wav.shape == ref.shape
        assert rmse(wav.cpu().numpy(), ref.numpy()) < 2e-5, (B, T)
    voc = VocoderBigVGAN(generator=m, device="cuda")
    spec = specs.synth_tensor((80, 17), seed=3, scale=2.0, shift=-4.0).numpy()
    out = voc.vocode(spec)                              # numpy [80, T] -> numpy [T * hop]  (models.py:406-411)
    assert isinstance(out, np.ndarray) and out.shape == (17 * 256,)
    ref = br.bigvgan_forward(sd, h, torch.from_numpy(spec)[None]).reshape(-1).numpy()
    assert rmse(out, ref) < 2e-5


def test_bigvgan_batch_rows_independent():
    h = specs.BIGVGAN_SMALL
    m = _model(h)
    mel = specs.synth_tensor((4, 80, 40), seed=8, scale=2.0, shift=-4.0).cuda()
    wav = m(mel)
    one = m(mel[2:3].contiguous())
    assert torch.isfinite(wav).all()
    assert rmse(one[0].cpu().numpy(), wav[2].cpu().numpy()) < 1e-6
