"""End-to-end and cross-cutting GPU tests: diffusion -> mel -> HiFi-GAN -> waveform pipeline against the
CPU oracle, the vocoder wrapper (numpy in / numpy out), and the three generations of the tcgen05
tap-GEMM kernel against the fp32-FMA kernel on the layer shapes of the BASELINE configs."""
import ctypes as C

import numpy as np
import pytest
import torch

from audiogpt_b200 import _lib, specs
from audiogpt_b200.modules.diff import shallow_diffusion_tts as sdt
from audiogpt_b200.modules.diff.net import DiffNet
from audiogpt_b200.modules.hifigan.hifigan import HifiGanGenerator
from audiogpt_b200.utils.hparams import set_hparams_from_dict
from audiogpt_b200.vocoders.hifigan import HifiGAN, get_vocoder_cls
from conftest import rel_rmse, rmse

pytestmark = pytest.mark.gpu


def test_c3_pipeline_small_vs_oracle():
    """BASELINE configs[2] flow at CPU-second size: 20-step ancestral sampling -> denorm -> HiFi-GAN."""
    from oracle import diffusion_ref as dr
    from oracle import hifigan_ref as hr
    cfg, h = specs.DIFFNET_SMALL, specs.HIFIGAN_SMALL
    set_hparams_from_dict(dict(cfg, keep_bins=80, schedule_type="linear", max_beta=0.06))
    net = DiffNet(80)
    sdn = specs.synth_diffnet(cfg, 2024)
    net.load_state_dict(sdn, strict=True)
    steps = 20
    gd = sdt.GaussianDiffusion(None, 80, net, timesteps=steps, K_step=steps, loss_type="l1",
                               betas=sdt.linear_beta_schedule(steps, 0.06), spec_min=specs.SPEC_MIN,
                               spec_max=specs.SPEC_MAX).eval().to("cuda")
    voc = HifiGanGenerator(h)
    sdh = specs.synth_hifigan(h, 1234)
    voc.load_state_dict(sdh, strict=True)
    voc = voc.eval().to("cuda")
    B, T = 2, 18
    x = specs.synth_tensor((B, 1, 80, T), seed=5)
    cond = specs.synth_tensor((B, cfg["hidden_size"], T), seed=6)
    noises = specs.synth_tensor((steps, B, 1, 80, T), seed=7)
    # ---- GPU path
    xg = gd.sample(cond.cuda(), x_start=x.cuda(), noises=noises.cuda())
    mel = gd.denorm_spec(xg[:, 0].transpose(1, 2))               # [B, T, 80]
    wav = voc(mel.transpose(1, 2).contiguous())
    # ---- oracle path
    tab = dr.schedule_tables(dr.linear_betas(steps, 0.06))
    xo = dr.sample_loop(sdn, cfg, tab, x, cond, noises)
    smin, smax = torch.tensor(specs.SPEC_MIN)[None, None], torch.tensor(specs.SPEC_MAX)[None, None]
    melo = dr.denorm_spec(xo[:, 0].transpose(1, 2), smin, smax)
    wavo = hr.hifigan_forward(sdh, h, melo.transpose(1, 2).contiguous())
    assert rel_rmse(mel.cpu(), melo) < 1e-4
    e = rmse(wav.cpu(), wavo)
    print("C3-small pipeline waveform RMSE:", e)
    assert wav.shape == (B, 1, T * 256) and e < 1e-4


def test_vocoder_wrapper_spec2wav():
    from oracle import hifigan_ref as hr
    h = specs.HIFIGAN_SMALL
    m = HifiGanGenerator(h)
    sd = specs.synth_hifigan(h, 1234)
    m.load_state_dict(sd, strict=True)
    m = m.eval().to("cuda")
    v = HifiGAN(model=m, config=h)
    assert get_vocoder_cls({"vocoder": "hifigan"}) is HifiGAN
    set_hparams_from_dict({})
    mel = specs.synth_tensor((31, 80), seed=9, scale=2.0, shift=-4.0).numpy()      # [T, 80] as the reference passes
    wav = v.spec2wav(mel)
    ref = hr.hifigan_forward(sd, h, torch.from_numpy(mel.T.copy())[None]).reshape(-1)
    assert isinstance(wav, np.ndarray) and wav.shape == (31 * 256,)
    assert rmse(wav, ref) < 2e-5


SHAPES = [  # G, L, Cin, Cout, K, dil, Wreal
    (2, 3000, 256, 256, 11, 5, 0), (2, 5000, 128, 128, 3, 3, 0), (2, 9000, 32, 32, 7, 1, 0),
    (3, 400, 256, 512, 3, 2, 0), (1, 777, 320, 320, 1, 1, 0), (2, 780, 320, 320, 3, 1, 78),
    (2, 195, 640, 640, 3, 1, 39), (1, 130, 1280, 320, 1, 1, 0), (2, 4, 64, 96, 3, 1, 0),
    (2, 300, 80, 256, 7, 1, 0), (1, 780, 4, 320, 3, 1, 78), (2, 500, 96, 40, 5, 2, 0),
]


@pytest.mark.parametrize("ver", [1, 2, 4, 5, 6, 7])
def test_tcgen05_generations_match_fma(ver):
    """agpt_bench_tapconv(check=1) runs the layer with the selected tcgen05 kernel and with the fp32-FMA
    kernel on the same random data and returns max |difference| (outputs are O(1))."""
    L = _lib.lib()
    torch.zeros(1).cuda()
    _lib.check(L.agpt_set_tc_version(ver))
    try:
        for G, Ln, Cin, Cout, K, dil, Wr in SHAPES:
            out = (C.c_double * 3)()
            _lib.check(L.agpt_bench_tapconv(G, Ln, Cin, Cout, K, dil, Wr, 1, 1, 1, 1, out, None))
            assert 0 <= out[2] < 5e-4, (ver, G, Ln, Cin, Cout, K, dil, Wr, out[2])
    finally:
        _lib.check(L.agpt_set_tc_version(-1))
