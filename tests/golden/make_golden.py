#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE's own modules on CPU fp32.

Run in the build container only (needs /root/reference, which does not travel
to the GPU box):

    python tests/golden/make_golden.py

It imports the reference classes from where they lie (nothing is copied), loads
seeded synthetic state dicts (audiogpt_b200.specs.synth_state_dict -- the
reference ships no weights), runs them, and writes small .npz files next to
this script.  Weights are NOT stored: tests regenerate them from the seed.

Shims: scipy.signal.kaiser alias; MagicMock for librosa/pycwt;
a stub omegaconf.listconfig module; a 10-line LDM model shim exposing what
DDIMSampler reads.
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from audiogpt_b200 import specs  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(os.cpu_count() or 1)


_SAVE_ONLY = None   # set of fixture names to (re)write; None = all


def save(name, **arrs):
    if _SAVE_ONLY is not None and name not in _SAVE_ONLY:
        print(f"(not rewriting {name}.npz)")
        return
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KB)")


def stats(x):
    x = x.double()
    return np.array([x.sum().item(), x.abs().sum().item(), (x * x).sum().item()], dtype=np.float64)


# --------------------------------------------------------------------------- NeuralSeq
def import_neuralseq():
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "kaiser"):
        scipy.signal.kaiser = scipy.signal.windows.kaiser
    import transformers  # noqa: F401  (must precede the librosa mock)
    for m in ("librosa", "librosa.filters", "librosa.util", "librosa.core", "pycwt"):
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, os.path.join(REF, "NeuralSeq"))


def golden_hifigan(nsf_only=False):
    from modules.hifigan.hifigan import HifiGanGenerator
    global _SAVE_ONLY
    if nsf_only:
        _SAVE_ONLY = {"nsf_source"}

    # ---- small config, B=2, ragged-ish T=24 -----------------------------------------
    h = specs.HIFIGAN_SMALL
    sd = specs.synth_hifigan(h, 1234)
    m = HifiGanGenerator(h)
    m.remove_weight_norm()
    m.load_state_dict(sd, strict=True)
    m.eval()
    mel = specs.synth_tensor((2, 80, 24), seed=11, scale=2.0, shift=-4.0)
    with torch.no_grad():
        wav = m(mel)
    print("hifigan small wav rms", wav.pow(2).mean().sqrt().item(), "absmax", wav.abs().max().item())

    # weight-norm (pre-removal) layout: g, v  -> strict load into the un-folded module
    m2 = HifiGanGenerator(h)
    sd_wn = {}
    g = torch.Generator().manual_seed(77)
    for k, v in sd.items():
        is_wn = k.endswith(".weight") and not k.startswith("noise_convs") and not k.startswith("m_source")
        if is_wn:
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
            gg = nrm * (0.5 + torch.rand(nrm.shape, generator=g))
            sd_wn[k[:-len("weight")] + "weight_g"] = gg
            sd_wn[k[:-len("weight")] + "weight_v"] = v * 3.0   # direction only matters
        else:
            sd_wn[k] = v
    m2.load_state_dict(sd_wn, strict=True)
    m2.eval()
    with torch.no_grad():
        wav_wn = m2(mel)
    save("hifigan_small", mel=mel, wav=wav, wav_wn=wav_wn,
         wn_keys=np.array(sorted(sd_wn.keys())),
         **{"wn::" + k: v for k, v in sd_wn.items() if k.endswith("weight_g")})

    # ---- ResBlock2 variant ----------------------------------------------------------
    h2 = dict(h, resblock="2", resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]])
    sd2 = specs.synth_hifigan(h2, 4321)
    m3 = HifiGanGenerator(h2)
    m3.remove_weight_norm()
    m3.load_state_dict(sd2, strict=True)
    m3.eval()
    mel2 = specs.synth_tensor((1, 80, 17), seed=12, scale=2.0, shift=-4.0)
    with torch.no_grad():
        wav2 = m3(mel2)
    save("hifigan_small_rb2", mel=mel2, wav=wav2)

    # ---- NSF variant (use_pitch_embed): capture the har_source the reference drew ----
    h3 = dict(h, use_pitch_embed=True, audio_sample_rate=24000)
    sd3 = specs.synth_hifigan(h3, 5678)
    m4 = HifiGanGenerator(h3)
    m4.remove_weight_norm()
    m4.load_state_dict(sd3, strict=True)
    m4.eval()
    cap = {}
    m4.m_source.register_forward_hook(lambda mod, inp, out: cap.__setitem__("har", out[0].detach().clone()))
    mel3 = specs.synth_tensor((2, 80, 20), seed=13, scale=2.0, shift=-4.0)
    f0 = 220.0 + 40.0 * specs.synth_tensor((2, 20), seed=14)
    f0[:, :3] = 0.0  # unvoiced head
    with torch.no_grad():
        wav3 = m4(mel3, f0)
    save("hifigan_small_nsf", mel=mel3, f0=f0, har_source=cap["har"].transpose(1, 2), wav=wav3)

    # ---- the NSF source module itself with its random draws pinned --------------------------
    # torch.rand / torch.randn_like are patched for the duration of the call: the reference then consumes
    # exactly these tensors (rand_ini, randn_like(sines), randn_like(uv) -- in that order, source.py:356,433,523)
    from unittest import mock
    def nsf_case(B, T, seed):
        f0f = 180.0 + 60.0 * specs.synth_tensor((B, T), seed=seed)          # frame-level F0 in Hz
        f0f[:, : T // 8] = 0.0                                              # an unvoiced head
        f0f[:, T // 2: T // 2 + max(1, T // 10)] = 0.0                      # and an unvoiced gap
        f0u = torch.repeat_interleave(f0f[:, None], 256, dim=2).transpose(1, 2)      # nn.Upsample(scale_factor=hop), nearest
        ri = torch.rand((B, 9), generator=torch.Generator().manual_seed(seed + 1))
        nz = specs.synth_tensor((B, T * 256, 9), seed=seed + 2)
        nz2 = specs.synth_tensor((B, T * 256, 1), seed=seed + 3)
        draws = [nz, nz2]
        with mock.patch.object(torch, "rand", lambda *a, **k: ri.clone()), \
                mock.patch.object(torch, "randn_like", lambda x: draws.pop(0)):
            with torch.no_grad():
                har, noi, uv = m4.m_source(f0u)
        assert not draws
        return f0f, ri, har
    f0a, ria, hara = nsf_case(2, 20, 150)
    f0b, rib, harb = nsf_case(1, 400, 160)            # 102 400 samples: a 100-chunk prefix sum
    print("nsf har rms", hara.pow(2).mean().sqrt().item(), harb.pow(2).mean().sqrt().item())
    save("nsf_source", f0_a=f0a, rand_ini_a=ria, har_a=hara[:, :, 0], f0_b=f0b, rand_ini_b=rib,
         har_b_head=harb[0, :8192, 0], har_b_stride=harb[0, ::53, 0], har_b_tail=harb[0, -4096:, 0])

    if nsf_only:
        _SAVE_ONLY = None
        return
    # ---- V1 (BASELINE configs[0]: 1x80x400), subsampled -------------------------------
    hv = specs.HIFIGAN_V1
    sdv = specs.synth_hifigan(hv, 1234)
    mv = HifiGanGenerator(hv)
    mv.remove_weight_norm()
    mv.load_state_dict(sdv, strict=True)
    mv.eval()
    melv = specs.synth_tensor((1, 80, 400), seed=0, scale=2.0, shift=-4.0)
    with torch.no_grad():
        wavv = mv(melv)
    print("hifigan V1 wav rms", wavv.pow(2).mean().sqrt().item(), "absmax", wavv.abs().max().item())
    save("hifigan_v1_c1", wav_head=wavv[0, 0, :4096], wav_stride=wavv[0, 0, ::37], stats=stats(wavv))


def golden_diffusion():
    from utils.hparams import hparams
    hparams.clear()
    cfg = specs.DIFFNET_SMALL
    hparams.update(hidden_size=cfg["hidden_size"], residual_layers=cfg["residual_layers"],
                   residual_channels=cfg["residual_channels"],
                   dilation_cycle_length=cfg["dilation_cycle_length"],
                   keep_bins=80, schedule_type="linear", max_beta=0.06)
    from modules.diff.net import DiffNet
    import modules.diff.shallow_diffusion_tts as sdt

    net = DiffNet(80)
    sd = specs.synth_diffnet(cfg, 2024)
    net.load_state_dict(sd, strict=True)
    net.eval()
    B, T = 3, 21
    x = specs.synth_tensor((B, 1, 80, T), seed=21)
    cond = specs.synth_tensor((B, cfg["hidden_size"], T), seed=22)
    t = torch.tensor([0, 37, 99], dtype=torch.long)
    with torch.no_grad():
        eps = net(x, t, cond)
    print("diffnet small eps rms", eps.pow(2).mean().sqrt().item())

    # GaussianDiffusion without building FastSpeech2: construct via __new__ and run the
    # reference's own __init__ body for the buffers with fs2 construction patched out.
    class _NoFS2(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    sdt.FastSpeech2 = _NoFS2
    sdt.FastSpeech2MIDI = _NoFS2
    timesteps = 100
    gd = sdt.GaussianDiffusion(None, 80, net, timesteps=timesteps, K_step=timesteps,
                               loss_type="l1", betas=sdt.linear_beta_schedule(timesteps, 0.06),
                               spec_min=specs.SPEC_MIN, spec_max=specs.SPEC_MAX)
    gd.eval()
    tabs = {k: getattr(gd, k) for k in (
        "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
        "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
        "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
        "posterior_mean_coef2")}

    # ancestral p_sample for a handful of step values with injected noise
    noise_bank = specs.synth_tensor((timesteps, B, 1, 80, T), seed=23)
    cur = {"i": 0}
    sdt.noise_like = lambda shape, device, repeat=False: noise_bank[cur["i"]]
    xs = [x]
    xx = x
    steps = [99, 98, 50, 1, 0]
    for i in steps:
        cur["i"] = i
        xx = gd.p_sample(xx, torch.full((B,), i, dtype=torch.long), cond)
        xs.append(xx)
    # full 100-step loop on the small net
    xl = x
    for i in reversed(range(timesteps)):
        cur["i"] = i
        xl = gd.p_sample(xl, torch.full((B,), i, dtype=torch.long), cond)
    mel_out = gd.denorm_spec(xl[:, 0].transpose(1, 2))
    print("p_sample loop x rms", xl.pow(2).mean().sqrt().item())

    # q_sample + norm_spec (shallow start)
    fs2_mel = specs.synth_tensor((B, T, 80), seed=24, scale=1.0, shift=-2.5)
    qn = specs.synth_tensor((B, 1, 80, T), seed=25)
    xq = gd.q_sample(gd.norm_spec(fs2_mel).transpose(1, 2)[:, None], torch.tensor([70]).long(), noise=qn)

    # PLMS (reference path only works at B=1: Python max() on a tensor)
    x1, c1 = x[:1], cond[:1]
    from collections import deque
    gd.noise_list = deque(maxlen=4)
    gd1000 = sdt.GaussianDiffusion(None, 80, net, timesteps=1000, K_step=1000, loss_type="l1",
                                   betas=sdt.linear_beta_schedule(1000, 0.02),
                                   spec_min=specs.SPEC_MIN, spec_max=specs.SPEC_MAX)
    gd1000.eval()
    gd1000.noise_list = deque(maxlen=4)
    xp = x1
    plms = []
    for i in reversed(range(0, 1000, 10)):
        xp = gd1000.p_sample_plms(xp, torch.full((1,), i, dtype=torch.long), 10, c1)
        if i in (990, 980, 970, 960, 500, 0):
            plms.append(xp)
    print("plms final rms", xp.pow(2).mean().sqrt().item())

    save("diffusion_small", x=x, cond=cond, t=t, eps=eps,
         steps=np.array(steps), xs=torch.stack(xs), x_loop=xl, mel_out=mel_out,
         fs2_mel=fs2_mel, q_noise=qn, x_q=xq,
         plms=torch.stack(plms), plms_final=xp,
         **{"tab::" + k: v for k, v in tabs.items()},
         **{"tab1000::alphas_cumprod": gd1000.alphas_cumprod})

    # ---- full-size DiffNet (C3 shape family), one forward, B=2, T=100 ----------------
    cfgb = specs.DIFFNET_BASE
    hparams.update(hidden_size=cfgb["hidden_size"], residual_layers=cfgb["residual_layers"],
                   residual_channels=cfgb["residual_channels"],
                   dilation_cycle_length=cfgb["dilation_cycle_length"])
    netb = DiffNet(80)
    sdb = specs.synth_diffnet(cfgb, 2025)
    netb.load_state_dict(sdb, strict=True)
    netb.eval()
    xb = specs.synth_tensor((2, 1, 80, 100), seed=31)
    cb = specs.synth_tensor((2, 256, 100), seed=32)
    tb = torch.tensor([99, 3], dtype=torch.long)
    with torch.no_grad():
        eb = netb(xb, tb, cb)
    print("diffnet base eps rms", eb.pow(2).mean().sqrt().item())
    save("diffusion_base_fwd", eps=eb)


def golden_pe():
    """PitchExtractor (NeuralSeq/modules/fastspeech/pe.py:119-148), small and base widths, ragged + padded mels."""
    from utils.hparams import hparams
    from modules.fastspeech.pe import PitchExtractor
    for name, cfg, B, T in (("pe_small", specs.PE_SMALL, 2, 37), ("pe_base", specs.PE_BASE, 2, 150)):
        hparams.clear()
        hparams.update(hidden_size=cfg["hidden_size"], predictor_hidden=cfg["predictor_hidden"], ffn_padding="SAME",
                       predictor_kernel=cfg["predictor_kernel"], pitch_type="frame", use_uv=True, pitch_norm="log", dropout=0.1)
        pe = PitchExtractor(cfg["n_mel_bins"], conv_layers=cfg["conv_layers"])
        print("pe load:", pe.load_state_dict(specs.synth_pe(cfg, 606), strict=True))
        pe.eval()
        mel = specs.synth_tensor((B, T, 80), seed=71, scale=1.0, shift=-2.5)
        mel[1, -T // 5:] = 0                      # padded tail (all-zero frames)
        with torch.no_grad():
            r = pe(mel)
        print(name, "pitch_pred rms", r["pitch_pred"].pow(2).mean().sqrt().item(), "f0 max", r["f0_denorm_pred"].max().item())
        save(name, pitch_pred=r["pitch_pred"], f0_denorm_pred=r["f0_denorm_pred"])


def golden_c3_full():
    """BASELINE configs[2] at full size: DiffNet base (20 x 256), B=16, T=400, 100 ancestral p_sample steps
    with the per-step noise of SURVEY.md 8d (seed 4), run by the reference's own GaussianDiffusion."""
    from utils.hparams import hparams
    hparams.clear()
    cfgb = specs.DIFFNET_BASE
    hparams.update(hidden_size=cfgb["hidden_size"], residual_layers=cfgb["residual_layers"],
                   residual_channels=cfgb["residual_channels"],
                   dilation_cycle_length=cfgb["dilation_cycle_length"],
                   keep_bins=80, schedule_type="linear", max_beta=0.06)
    from modules.diff.net import DiffNet
    import modules.diff.shallow_diffusion_tts as sdt

    class _NoFS2(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    sdt.FastSpeech2 = _NoFS2
    sdt.FastSpeech2MIDI = _NoFS2
    net = DiffNet(80)
    net.load_state_dict(specs.synth_diffnet(cfgb, 2025), strict=True)
    net.eval()
    gd = sdt.GaussianDiffusion(None, 80, net, timesteps=100, K_step=100, loss_type="l1",
                               betas=sdt.linear_beta_schedule(100, 0.06),
                               spec_min=specs.SPEC_MIN, spec_max=specs.SPEC_MAX)
    gd.eval()
    B, T = 16, 400
    x = specs.synth_tensor((B, 1, 80, T), seed=2)
    cond = specs.synth_tensor((B, 256, T), seed=3)
    cur = {"i": 0}
    sdt.noise_like = lambda shape, device, repeat=False: specs.synth_tensor((B, 1, 80, T), seed=4000 + cur["i"])
    import time
    t0 = time.time()
    for i in reversed(range(100)):
        cur["i"] = i
        x = gd.p_sample(x, torch.full((B,), i, dtype=torch.long), cond)
        if i % 10 == 0:
            print(f"  c3 step {i}: rms {x.pow(2).mean().sqrt().item():.4f}  ({time.time() - t0:.0f} s)", flush=True)
    mel = gd.denorm_spec(x[:, 0].transpose(1, 2))
    save("diffusion_c3_full", x_end=x[:, :, :, ::4], mel_end=mel[:, ::4, :], stats=stats(x))


# --------------------------------------------------------------------------- Make-An-Audio
def import_ldm():
    oc = types.ModuleType("omegaconf")
    lc = types.ModuleType("omegaconf.listconfig")

    class ListConfig(list):
        pass

    lc.ListConfig = ListConfig
    oc.listconfig = lc
    oc.OmegaConf = MagicMock()     # vocoder/bigvgan/models.py imports the name (used only by VocoderBigVGAN.__init__)
    sys.modules.setdefault("omegaconf", oc)
    sys.modules.setdefault("omegaconf.listconfig", lc)
    # the NeuralSeq tree also has a top-level 'modules'/'utils'; make sure ldm wins its own names
    sys.path.insert(0, os.path.join(REF, "text_to_audio", "Make_An_Audio"))


class LDMShim:
    """What DDIMSampler reads from the LatentDiffusion object (ddim.py:17,30-36,124,175)."""

    def __init__(self, unet, tab):
        self.unet = unet
        self.num_timesteps = tab["betas"].shape[0]
        self.betas = tab["betas"]
        self.alphas_cumprod = tab["alphas_cumprod"]
        self.alphas_cumprod_prev = tab["alphas_cumprod_prev"]
        self.device = torch.device("cpu")
        self.parameterization = "eps"

    def apply_model(self, x, t, c):
        return self.unet(x, timesteps=t, context=c)


def golden_ldm(only100=False):
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from ldm.modules.diffusionmodules.util import make_beta_schedule
    from ldm.models.diffusion.ddim import DDIMSampler

    betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    ac = np.cumprod(1.0 - betas, axis=0)
    tab = dict(betas=torch.tensor(betas, dtype=torch.float32),
               alphas_cumprod=torch.tensor(ac, dtype=torch.float32),
               alphas_cumprod_prev=torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32))

    def build(cfg, seed):
        u = UNetModel(image_size=32, use_checkpoint=True, **cfg)
        sd = specs.synth_unet(cfg, seed)
        u.load_state_dict(sd, strict=True)
        u.eval()
        return u

    if not only100:
        _golden_ldm_small(build, tab, DDIMSampler)
    _golden_ldm_full(build, tab, DDIMSampler, only100)


def _golden_ldm_small(build, tab, DDIMSampler):
    # ---- small UNet: forward + DDIM-10 with CFG ---------------------------------------
    cfg = specs.UNET_SMALL
    u = build(cfg, 3030)
    N, H, W, S = 2, 6, 10, 7
    x = specs.synth_tensor((N, 4, H, W), seed=41)
    ctx = specs.synth_tensor((N, S, cfg["context_dim"]), seed=42)
    t = torch.tensor([991, 1], dtype=torch.long)
    with torch.no_grad():
        eps = u(x, timesteps=t, context=ctx)
    print("unet small eps rms", eps.pow(2).mean().sqrt().item())
    # odd spatial size (5x39-like after stride 2): H=5, W=7 -> 3x4 -> up 6x8 != 5x7 would break
    # the skip cat, so the reference itself requires even H,W here; keep even sizes.
    smp = DDIMSampler(LDMShim(u, tab))
    uc = specs.synth_tensor((1, S, cfg["context_dim"]), seed=43).expand(N, -1, -1).contiguous()
    xT = torch.tensor(np.random.RandomState(55).randn(N, 4, H, W), dtype=torch.float32)
    out, inter = smp.sample(S=10, batch_size=N, shape=(4, H, W), conditioning=ctx, verbose=False,
                            unconditional_guidance_scale=1.5, unconditional_conditioning=uc,
                            eta=0.0, x_T=xT)
    tabs10 = dict(ddim_timesteps=smp.ddim_timesteps.copy(), ddim_alphas=smp.ddim_alphas.clone(),
                  ddim_alphas_prev=np.asarray(smp.ddim_alphas_prev, dtype=np.float64),
                  ddim_sqrt_one_minus_alphas=smp.ddim_sqrt_one_minus_alphas.clone())
    out_nocfg, _ = smp.sample(S=5, batch_size=N, shape=(4, H, W), conditioning=ctx, verbose=False,
                              eta=0.0, x_T=xT)
    print("ddim small out rms", out.pow(2).mean().sqrt().item())
    save("ldm_small", x=x, ctx=ctx, t=t, eps=eps, uc=uc, x_T=xT, ddim10=out, ddim5_nocfg=out_nocfg,
         **tabs10,
         alphas_cumprod=tab["alphas_cumprod"])



def _golden_ldm_full(build, tab, DDIMSampler, only100):
    # ---- full txt2audio UNet (C4 shape): one CFG-pair forward + DDIM-100 first 4 steps --
    cfgf = specs.UNET_TXT2AUDIO
    uf = build(cfgf, 4040)
    xf = torch.tensor(np.random.RandomState(55).randn(1, 4, 10, 78), dtype=torch.float32)
    cf = specs.synth_tensor((1, 77, 1024), seed=5)
    ucf = specs.synth_tensor((1, 77, 1024), seed=6)
    tf = torch.tensor([991, 991], dtype=torch.long)
    with torch.no_grad():
        ef = uf(torch.cat([xf, xf]), timesteps=tf, context=torch.cat([ucf, cf]))
    print("unet full eps rms", ef.pow(2).mean().sqrt().item())
    smpf = DDIMSampler(LDMShim(uf, tab))
    smpf.make_schedule(ddim_num_steps=100, ddim_eta=0.0, verbose=False)
    img = xf
    for i, step in enumerate(np.flip(smpf.ddim_timesteps)[:4]):
        idx = 100 - i - 1
        ts = torch.full((1,), int(step), dtype=torch.long)
        img, _ = smpf.p_sample_ddim(img, cf, ts, index=idx, unconditional_guidance_scale=1.5,
                                    unconditional_conditioning=ucf)
    print("ddim full 4 steps rms", img.pow(2).mean().sqrt().item())
    if not only100:
        save("ldm_txt2audio", eps_pair=ef, ddim100_first4=img)

    # ---- the whole DDIM-100 + CFG 1.5 chain (C4 per-clip work, B=1): end point and last pred_x0 ----
    out100, inter100 = smpf.sample(S=100, batch_size=1, shape=(4, 10, 78), conditioning=cf, verbose=False,
                                   unconditional_guidance_scale=1.5, unconditional_conditioning=ucf,
                                   eta=0.0, x_T=xf)
    print("ddim-100 end point rms", out100.pow(2).mean().sqrt().item(), "absmax", out100.abs().max().item())
    save("ldm_txt2audio_ddim100", x_T=xf, ddim100=out100, pred_x0_last=inter100["pred_x0"][-1])


def golden_bigvgan():
    """BigVGAN generator of Make-An-Audio's vocoder (vocoder/bigvgan/models.py:133-203), small config."""
    from vocoder.bigvgan.models import BigVGAN

    class AttrDict(dict):
        __getattr__ = dict.__getitem__

    h = specs.BIGVGAN_SMALL
    sd = specs.synth_bigvgan(h, 4321)
    m = BigVGAN(AttrDict(h))
    m.remove_weight_norm()
    missing = m.load_state_dict(sd, strict=True)
    print("bigvgan load:", missing)
    m.eval()
    mel = specs.synth_tensor((2, 80, 20), seed=5, scale=2.0, shift=-4.0)
    with torch.no_grad():
        wav = m(mel)
        wav1 = m(mel[:1, :, :7])      # ragged / shorter than the replicate pads
    print("bigvgan small wav rms", wav.pow(2).mean().sqrt().item(), "absmax", wav.abs().max().item())
    save("bigvgan_small", wav=wav, wav_t7=wav1)


def golden_vae():
    """AutoencoderKL.decode = post_quant_conv -> Decoder (ldm/models/autoencoder.py:351-354; the LightningModule
    itself needs pytorch_lightning/taming, so the two sub-modules it calls are instantiated directly)."""
    from ldm.modules.diffusionmodules.model import Decoder

    def run(cfg, seed, z):
        sd = specs.synth_vae_decoder(cfg, seed)
        dec = Decoder(**{k: v for k, v in cfg.items() if k != "embed_dim"})
        print("vae decoder load:", dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items()
                                                        if k.startswith("decoder.")}, strict=True))
        pq = torch.nn.Conv2d(cfg["embed_dim"], cfg["z_channels"], 1)
        pq.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})
        dec.eval()
        with torch.no_grad():
            return dec(pq(z))

    z = specs.synth_tensor((2, 4, 10, 78), seed=3)
    ys = run(specs.VAE_SMALL, 5150, z)
    print("vae small out rms", ys.pow(2).mean().sqrt().item(), tuple(ys.shape))
    save("vae_small", mel=ys[:, :, ::2, ::3], stats=stats(ys))
    yf = run(specs.VAE_TXT2AUDIO, 5150, z[:1])
    print("vae txt2audio out rms", yf.pow(2).mean().sqrt().item(), tuple(yf.shape))
    save("vae_txt2audio", mel=yf[:, :, ::2, ::3], stats=stats(yf))


if __name__ == "__main__":
    which = sys.argv[1:] or ["hifigan", "diffusion", "c3", "pe", "ldm", "bigvgan", "vae"]   # extra selector: ldm100 (DDIM-100 end point only)
    if "nsf" in which:
        which = list(which) + ["hifigan_nsf_only"]
    if "hifigan" in which or "diffusion" in which or "c3" in which or "pe" in which or "hifigan_nsf_only" in which:
        import_neuralseq()
        cwd = os.getcwd()
        os.chdir(os.path.join(REF, "NeuralSeq"))
        try:
            if "hifigan" in which or "hifigan_nsf_only" in which:
                golden_hifigan(nsf_only="hifigan" not in which)
            if "diffusion" in which:
                golden_diffusion()
            if "c3" in which:
                golden_c3_full()
            if "pe" in which:
                golden_pe()
        finally:
            os.chdir(cwd)
    if "ldm" in which or "ldm100" in which:
        # drop NeuralSeq's top-level packages so that Make-An-Audio's resolve
        for k in [k for k in sys.modules if k.split(".")[0] in ("modules", "utils", "vocoders", "tasks")]:
            del sys.modules[k]
        if os.path.join(REF, "NeuralSeq") in sys.path:
            sys.path.remove(os.path.join(REF, "NeuralSeq"))
        import_ldm()
        golden_ldm(only100="ldm" not in which)
    if "bigvgan" in which:
        for k in [k for k in sys.modules if k.split(".")[0] in ("modules", "utils", "vocoders", "tasks")]:
            del sys.modules[k]
        if os.path.join(REF, "NeuralSeq") in sys.path:
            sys.path.remove(os.path.join(REF, "NeuralSeq"))
        import_ldm()
        golden_bigvgan()
    if "vae" in which:
        for k in [k for k in sys.modules if k.split(".")[0] in ("modules", "utils", "vocoders", "tasks")]:
            del sys.modules[k]
        if os.path.join(REF, "NeuralSeq") in sys.path:
            sys.path.remove(os.path.join(REF, "NeuralSeq"))
        import_ldm()
        golden_vae()
