"""Pin the CPU oracle (oracle/*.py) against fixtures produced by the reference's
own modules (tests/golden/make_golden.py).  CPU-only; no GPU, no /root/reference."""
import numpy as np
import torch

from audiogpt_b200 import specs
from oracle import diffusion_ref as dr
from oracle import hifigan_ref as hr
from oracle import ldm_ref as lr
from conftest import load_golden, rel_rmse, rmse

T = torch.tensor


def test_hifigan_small():
    g = load_golden("hifigan_small")
    h = specs.HIFIGAN_SMALL
    sd = specs.synth_hifigan(h, 1234)
    wav = hr.hifigan_forward(sd, h, T(g["mel"]))
    assert wav.shape == (2, 1, 24 * 256)
    assert rel_rmse(wav, g["wav"]) < 1e-6   # same torch ops; summation order may differ per box


def test_hifigan_weight_norm_fold():
    g = load_golden("hifigan_small")
    h = specs.HIFIGAN_SMALL
    sd = specs.synth_hifigan(h, 1234)
    sd_wn = {}
    for k, v in sd.items():
        if k.endswith(".weight"):
            sd_wn[k[:-6] + "weight_g"] = T(g["wn::" + k[:-6] + "weight_g"])
            sd_wn[k[:-6] + "weight_v"] = v * 3.0
        else:
            sd_wn[k] = v
    assert sorted(sd_wn.keys()) == list(g["wn_keys"])
    wav = hr.hifigan_forward(sd_wn, h, T(g["mel"]))
    # folding g*v/||v|| on our side vs torch's weight-norm hook: same formula, fp32 rounding may differ
    assert rel_rmse(wav, g["wav_wn"]) < 1e-6


def test_hifigan_resblock2():
    g = load_golden("hifigan_small_rb2")
    h = dict(specs.HIFIGAN_SMALL, resblock="2", resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]])
    wav = hr.hifigan_forward(specs.synth_hifigan(h, 4321), h, T(g["mel"]))
    assert rel_rmse(wav, g["wav"]) < 1e-6   # same torch ops; summation order may differ per box


def test_hifigan_nsf():
    g = load_golden("hifigan_small_nsf")
    h = dict(specs.HIFIGAN_SMALL, use_pitch_embed=True, audio_sample_rate=24000)
    wav = hr.hifigan_forward(specs.synth_hifigan(h, 5678), h, T(g["mel"]), T(g["har_source"]))
    assert rel_rmse(wav, g["wav"]) < 1e-6   # same torch ops; summation order may differ per box


def test_hifigan_v1_c1():
    g = load_golden("hifigan_v1_c1")
    h = specs.HIFIGAN_V1
    mel = specs.synth_tensor((1, 80, 400), seed=0, scale=2.0, shift=-4.0)
    wav = hr.hifigan_forward(specs.synth_hifigan(h, 1234), h, mel)
    assert wav.shape == (1, 1, 102400)
    # thread-count dependent summation order inside torch's conv => not bit-exact across boxes
    assert rel_rmse(wav[0, 0, :4096], g["wav_head"]) < 2e-5
    assert rel_rmse(wav[0, 0, ::37], g["wav_stride"]) < 2e-5
    assert abs(hr.hifigan_flops(h, 400) / 245.64e9 - 1) < 0.01


def _nsf_inputs(g, tag, B, Tn, seed):
    f0f = T(g["f0_" + tag])
    f0u = torch.repeat_interleave(f0f[:, None], 256, dim=2).transpose(1, 2)
    nz = specs.synth_tensor((B, Tn * 256, 9), seed=seed + 2)
    return f0u, T(g["rand_ini_" + tag]), nz


def test_nsf_source_oracle_vs_reference():
    """oracle.hifigan_ref.nsf_source == the reference's SourceModuleHnNSF with its random draws pinned
    (fixture: make_golden.py nsf), short and 102 400-sample utterances."""
    g = load_golden("nsf_source")
    h3 = dict(specs.HIFIGAN_SMALL, use_pitch_embed=True, audio_sample_rate=24000)
    sd = specs.synth_hifigan(h3, 5678)
    w, b = sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"]
    f0u, ri, nz = _nsf_inputs(g, "a", 2, 20, 150)
    har = hr.nsf_source(w, b, f0u, 24000, ri, nz)
    assert rmse(har[:, :, 0], g["har_a"]) < 1e-7
    f0u, ri, nz = _nsf_inputs(g, "b", 1, 400, 160)
    har = hr.nsf_source(w, b, f0u, 24000, ri, nz)[0, :, 0]
    assert rmse(har[:8192], g["har_b_head"]) < 1e-7 and rmse(har[::53], g["har_b_stride"]) < 1e-7
    assert rmse(har[-4096:], g["har_b_tail"]) < 1e-7


def test_schedule_tables_bit_exact():
    g = load_golden("diffusion_small")
    tab = dr.schedule_tables(dr.linear_betas(100, 0.06))
    for k, v in tab.items():
        assert torch.equal(v, T(g["tab::" + k])), k
    tab1000 = dr.schedule_tables(dr.linear_betas(1000, 0.02))
    assert torch.equal(tab1000["alphas_cumprod"], T(g["tab1000::alphas_cumprod"]))


def test_diffnet_and_p_sample():
    g = load_golden("diffusion_small")
    cfg = specs.DIFFNET_SMALL
    sd = specs.synth_diffnet(cfg, 2024)
    x, cond, t = T(g["x"]), T(g["cond"]), T(g["t"])
    eps = dr.diffnet_forward(sd, cfg, x, t, cond)
    assert rel_rmse(eps, g["eps"]) < 1e-6
    tab = dr.schedule_tables(dr.linear_betas(100, 0.06))
    bank = specs.synth_tensor((100,) + tuple(x.shape), seed=23)
    fn = lambda a, b, c: dr.diffnet_forward(sd, cfg, a, b, c)
    xx = x
    for n, i in enumerate(g["steps"]):
        xx = dr.p_sample(tab, fn, xx, torch.full((x.shape[0],), int(i), dtype=torch.long), cond, bank[int(i)])
        assert rel_rmse(xx, g["xs"][n + 1]) < 1e-5, i
    xl = dr.sample_loop(sd, cfg, tab, x, cond, bank)
    assert rel_rmse(xl, g["x_loop"]) < 1e-4
    smin, smax = T(specs.SPEC_MIN)[None, None], T(specs.SPEC_MAX)[None, None]
    mel = dr.denorm_spec(xl[:, 0].transpose(1, 2), smin, smax)
    assert rel_rmse(mel, g["mel_out"]) < 1e-4
    xq = dr.q_sample(tab, dr.norm_spec(T(g["fs2_mel"]), smin, smax).transpose(1, 2)[:, None],
                     torch.tensor([70]), T(g["q_noise"]))
    assert rel_rmse(xq, g["x_q"]) < 1e-6


def test_plms():
    g = load_golden("diffusion_small")
    cfg = specs.DIFFNET_SMALL
    sd = specs.synth_diffnet(cfg, 2024)
    tab = dr.schedule_tables(dr.linear_betas(1000, 0.02))
    fn = lambda a, b, c: dr.diffnet_forward(sd, cfg, a, b, c)
    x, cond = T(g["x"])[:1], T(g["cond"])[:1]
    hist, snaps = [], []
    for i in reversed(range(0, 1000, 10)):
        x = dr.p_sample_plms(tab, fn, x, torch.full((1,), i, dtype=torch.long), 10, cond, hist)
        if i in (990, 980, 970, 960, 500, 0):
            snaps.append(x)
    for n, s in enumerate(snaps):
        assert rel_rmse(s, g["plms"][n]) < 1e-4, n
    assert rel_rmse(x, g["plms_final"]) < 1e-4


def test_diffnet_base_forward():
    g = load_golden("diffusion_base_fwd")
    cfg = specs.DIFFNET_BASE
    sd = specs.synth_diffnet(cfg, 2025)
    xb = specs.synth_tensor((2, 1, 80, 100), seed=31)
    cb = specs.synth_tensor((2, 256, 100), seed=32)
    eb = dr.diffnet_forward(sd, cfg, xb, torch.tensor([99, 3]), cb)
    assert rel_rmse(eb, g["eps"]) < 1e-5
    assert abs(dr.diffnet_flops_per_frame(cfg) / 26.44e6 - 1) < 0.01


def test_unet_small_and_ddim():
    g = load_golden("ldm_small")
    cfg = specs.UNET_SMALL
    sd = specs.synth_unet(cfg, 3030)
    eps = lr.unet_forward(sd, cfg, T(g["x"]), T(g["t"]), T(g["ctx"]))
    assert rel_rmse(eps, g["eps"]) < 1e-5
    sch = lr.ldm_schedule()
    assert torch.equal(sch["alphas_cumprod"], T(g["alphas_cumprod"]))
    tab = lr.ddim_tables(sch["alphas_cumprod"], 10)
    assert np.array_equal(tab["timesteps"], g["ddim_timesteps"])
    assert torch.equal(torch.as_tensor(tab["alphas"]), T(g["ddim_alphas"]))
    assert np.array_equal(np.asarray(tab["alphas_prev"], dtype=np.float64), g["ddim_alphas_prev"])
    assert torch.equal(torch.as_tensor(tab["sqrt_one_minus_alphas"]), T(g["ddim_sqrt_one_minus_alphas"]))
    fn = lambda x, t, c: lr.unet_forward(sd, cfg, x, t, c)
    out = lr.ddim_sample(fn, sch["alphas_cumprod"], 10, T(g["x_T"]), T(g["ctx"]), T(g["uc"]), 1.5)
    assert rel_rmse(out, g["ddim10"]) < 1e-4
    out5 = lr.ddim_sample(fn, sch["alphas_cumprod"], 5, T(g["x_T"]), T(g["ctx"]))
    assert rel_rmse(out5, g["ddim5_nocfg"]) < 1e-4


def test_unet_txt2audio_forward():
    g = load_golden("ldm_txt2audio")
    cfg = specs.UNET_TXT2AUDIO
    sd = specs.synth_unet(cfg, 4040)
    xf = torch.tensor(np.random.RandomState(55).randn(1, 4, 10, 78), dtype=torch.float32)
    cf = specs.synth_tensor((1, 77, 1024), seed=5)
    ucf = specs.synth_tensor((1, 77, 1024), seed=6)
    ef = lr.unet_forward(sd, cfg, torch.cat([xf, xf]), torch.tensor([991, 991]), torch.cat([ucf, cf]))
    assert rel_rmse(ef, g["eps_pair"]) < 2e-5
    sch = lr.ldm_schedule()
    fn = lambda x, t, c: lr.unet_forward(sd, cfg, x, t, c)
    out = lr.ddim_sample(fn, sch["alphas_cumprod"], 100, xf, cf, ucf, 1.5, steps_limit=4)
    assert rel_rmse(out, g["ddim100_first4"]) < 1e-4


def test_ddim100_full_chain():
    """The oracle's whole DDIM-100 + CFG chain (200 forwards of the 160 M-param UNet, B = 1) against the end point of
    the reference's own sampler (fixture made by `make_golden.py ldm100`).  This also measures how much this chain
    amplifies a 1e-6-level perturbation (different summation order only): the GPU gate in test_ldm_gpu.py is set
    from it."""
    g = load_golden("ldm_txt2audio_ddim100")
    cfg = specs.UNET_TXT2AUDIO
    sd = specs.synth_unet(cfg, 4040)
    cf = specs.synth_tensor((1, 77, 1024), seed=5)
    ucf = specs.synth_tensor((1, 77, 1024), seed=6)
    sch = lr.ldm_schedule()
    fn = lambda x, t, c: lr.unet_forward(sd, cfg, x, t, c)
    out = lr.ddim_sample(fn, sch["alphas_cumprod"], 100, T(g["x_T"]), cf, ucf, 1.5)
    e = rel_rmse(out, g["ddim100"])
    print("oracle DDIM-100 end point rel-RMSE vs reference:", e)
    assert e < 1e-4


def test_c3_full_chain():
    """The oracle's full-size C3 chain (DiffNet 20 x 256, B = 16, T = 400, 100 ancestral steps) against the
    reference's own GaussianDiffusion (fixture made by `make_golden.py c3`)."""
    g = load_golden("diffusion_c3_full")
    cfg = specs.DIFFNET_BASE
    sd = specs.synth_diffnet(cfg, 2025)
    tab = dr.schedule_tables(dr.linear_betas(100, 0.06))
    B, Tn = 16, 400
    x = specs.synth_tensor((B, 1, 80, Tn), seed=2)
    cond = specs.synth_tensor((B, 256, Tn), seed=3)
    fn = lambda a, b, c: dr.diffnet_forward(sd, cfg, a, b, c)
    for i in reversed(range(100)):
        x = dr.p_sample(tab, fn, x, torch.full((B,), i, dtype=torch.long), cond,
                        specs.synth_tensor((B, 1, 80, Tn), seed=4000 + i))
    e = rel_rmse(x[:, :, :, ::4], g["x_end"])
    print("oracle C3 full chain rel-RMSE vs reference:", e)
    assert e < 1e-4
    smin, smax = T(specs.SPEC_MIN)[None, None], T(specs.SPEC_MAX)[None, None]
    assert rel_rmse(dr.denorm_spec(x[:, 0].transpose(1, 2), smin, smax)[:, ::4, :], g["mel_end"]) < 1e-4


def test_pitch_extractor_oracle_vs_reference():
    """oracle/pe_ref.py == the reference's PitchExtractor (fixtures: make_golden.py pe), padded tail included."""
    from oracle import pe_ref
    for name, cfg, B, Tn in (("pe_small", specs.PE_SMALL, 2, 37), ("pe_base", specs.PE_BASE, 2, 150)):
        g = load_golden(name)
        mel = specs.synth_tensor((B, Tn, 80), seed=71, scale=1.0, shift=-2.5)
        mel[1, -Tn // 5:] = 0
        r = pe_ref.pe_forward(specs.synth_pe(cfg, 606), cfg, mel)
        assert rmse(r["pitch_pred"], g["pitch_pred"]) < 1e-6 and rmse(r["f0_denorm_pred"], g["f0_denorm_pred"]) < 1e-6
    n = sum(int(np.prod(v)) for k, v in specs.pe_param_shapes(specs.PE_BASE).items()
            if not k.endswith(("running_mean", "running_var", "num_batches_tracked", "_float_tensor")))
    assert n == 3_257_091                                  # parameters of the reference module at hidden 256


def test_bigvgan_small():
    """oracle/bigvgan_ref.py == the reference's BigVGAN module (fixture made by make_golden.py bigvgan)."""
    from oracle import bigvgan_ref as br
    g = load_golden("bigvgan_small")
    h = specs.BIGVGAN_SMALL
    sd = specs.synth_bigvgan(h, 4321)
    mel = specs.synth_tensor((2, 80, 20), seed=5, scale=2.0, shift=-4.0)
    wav = br.bigvgan_forward(sd, h, mel)
    assert wav.shape == (2, 1, 20 * 256)
    assert rel_rmse(wav, g["wav"]) < 1e-6
    wav7 = br.bigvgan_forward(sd, h, mel[:1, :, :7])
    assert rel_rmse(wav7, g["wav_t7"]) < 1e-6
    assert float(np.abs(g["wav"]).max()) < 0.9           # tanh not saturated: the comparison is meaningful
    # the filter buffers of the state dict are the published Kaiser-sinc taps
    f = specs.kaiser_sinc_filter12().reshape(-1)
    assert abs(float(f.sum()) - 1.0) < 1e-6 and torch.allclose(f, f.flip(0), atol=1e-7)


def test_vae_decode_small_and_txt2audio():
    """oracle/vae_ref.py == the reference's Decoder + post_quant_conv (fixtures hold a strided
    view and the global sum / |sum| / sum of squares of the full 80 x 624 output)."""
    from oracle import vae_ref as vr
    z = specs.synth_tensor((2, 4, 10, 78), seed=3)
    for name, cfg, zz in (("vae_small", specs.VAE_SMALL, z), ("vae_txt2audio", specs.VAE_TXT2AUDIO, z[:1])):
        g = load_golden(name)
        y = vr.vae_decode(specs.synth_vae_decoder(cfg, 5150), cfg, zz)
        assert y.shape == (zz.shape[0], 1, 80, 624)
        assert rel_rmse(y[:, :, ::2, ::3], g["mel"]) < 2e-5
        yd = y.double()
        st = np.array([yd.sum().item(), yd.abs().sum().item(), (yd * yd).sum().item()])
        assert np.allclose(st[1:], g["stats"][1:], rtol=1e-4)
    n = sum(int(np.prod(s)) for s in specs.vae_decoder_param_shapes(specs.VAE_TXT2AUDIO).values())
    assert abs(n / 41.0e6 - 1) < 0.01                                  # 41 M params
    assert abs(vr.vae_decode_flops(specs.VAE_TXT2AUDIO, 10, 78) / 392.9e9 - 1) < 0.01   # 392.9 GFLOP / clip
