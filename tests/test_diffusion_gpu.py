"""DiffNet + GaussianDiffusion parity on the GPU (through the C ABI) vs golden vectors from
the reference modules and vs the CPU oracle.  Stated tolerance: relative RMSE <= 1e-4 on
eps / x_t / mel (fp32 everywhere; only the summation order differs)."""
from collections import deque

import numpy as np
import pytest
import torch

from audiogpt_b200 import specs
from audiogpt_b200.modules.diff import shallow_diffusion_tts as sdt
from audiogpt_b200.modules.diff.net import DiffNet
from audiogpt_b200.utils.hparams import set_hparams_from_dict
from conftest import load_golden, rel_rmse

pytestmark = pytest.mark.gpu
TOL = 1e-4
T = torch.tensor


def make(cfg, seed, timesteps=100, max_beta=0.06):
    set_hparams_from_dict(dict(hidden_size=cfg["hidden_size"], residual_layers=cfg["residual_layers"],
                               residual_channels=cfg["residual_channels"],
                               dilation_cycle_length=cfg["dilation_cycle_length"], keep_bins=80,
                               schedule_type="linear", max_beta=max_beta))
    net = DiffNet(80)
    net.load_state_dict(specs.synth_diffnet(cfg, seed), strict=True)
    gd = sdt.GaussianDiffusion(None, 80, net, timesteps=timesteps, K_step=timesteps, loss_type="l1",
                               betas=sdt.linear_beta_schedule(timesteps, max_beta),
                               spec_min=specs.SPEC_MIN, spec_max=specs.SPEC_MAX)
    return gd.eval().to("cuda")


def test_state_dict_layout_matches_reference_keys():
    cfg = specs.DIFFNET_SMALL
    set_hparams_from_dict(cfg)
    net = DiffNet(80)
    assert list(net.state_dict().keys()) == list(specs.diffnet_param_shapes(cfg).keys())


def test_schedule_buffers_bit_exact():
    g = load_golden("diffusion_small")
    gd = make(specs.DIFFNET_SMALL, 2024)
    for k in g.files:
        if k.startswith("tab::"):
            assert torch.equal(getattr(gd, k[5:]).cpu(), T(g[k])), k


def test_diffnet_eps_small():
    g = load_golden("diffusion_small")
    gd = make(specs.DIFFNET_SMALL, 2024)
    eps = gd.denoise_fn(T(g["x"]).cuda(), T(g["t"]).cuda(), T(g["cond"]).cuda()).cpu()
    e = rel_rmse(eps, g["eps"])
    print("diffnet small eps rel-RMSE:", e)
    assert e < TOL


def test_p_sample_steps_and_loop(monkeypatch):
    g = load_golden("diffusion_small")
    gd = make(specs.DIFFNET_SMALL, 2024)
    x, cond = T(g["x"]).cuda(), T(g["cond"]).cuda()
    bank = specs.synth_tensor((100,) + tuple(x.shape), seed=23).cuda()
    cur = {"i": 0}
    monkeypatch.setattr(sdt, "noise_like", lambda shape, device, repeat=False: bank[cur["i"]])
    xx = x
    for n, i in enumerate(g["steps"]):
        cur["i"] = int(i)
        xx = gd.p_sample(xx, torch.full((x.shape[0],), int(i), device="cuda", dtype=torch.long), cond)
        assert rel_rmse(xx.cpu(), g["xs"][n + 1]) < TOL, i
    xl = gd.sample(cond, x_start=x, noises=bank)
    e = rel_rmse(xl.cpu(), g["x_loop"])
    print("100-step p_sample loop rel-RMSE:", e)
    assert e < 1e-3   # 100 recursive steps through clamp(); still fp32-level
    mel = gd.denorm_spec(xl[:, 0].transpose(1, 2))
    assert rel_rmse(mel.cpu(), g["mel_out"]) < 1e-3
    xq = gd.q_sample(gd.norm_spec(T(g["fs2_mel"]).cuda()).transpose(1, 2)[:, None],
                     torch.tensor([70], device="cuda"), noise=T(g["q_noise"]).cuda())
    assert rel_rmse(xq.cpu(), g["x_q"]) < 1e-6


def test_plms_b1_vs_reference_and_b2_extension():
    g = load_golden("diffusion_small")
    gd = make(specs.DIFFNET_SMALL, 2024, timesteps=1000, max_beta=0.02)
    assert torch.equal(gd.alphas_cumprod.cpu(), T(g["tab1000::alphas_cumprod"]))
    x, cond = T(g["x"])[:1].cuda(), T(g["cond"])[:1].cuda()
    gd.noise_list = deque(maxlen=4)
    snaps = []
    for i in reversed(range(0, 1000, 10)):
        x = gd.p_sample_plms(x, torch.full((1,), i, device="cuda", dtype=torch.long), 10, cond)
        if i in (990, 980, 970, 960, 500, 0):
            snaps.append(x.cpu())
    for n, s in enumerate(snaps):
        assert rel_rmse(s, g["plms"][n]) < 1e-3, n
    # B=2 (reference raises here): both rows must follow the B=1 trajectory of their row
    x2, c2 = T(g["x"])[:2].cuda(), T(g["cond"])[:2].cuda()
    gd.noise_list = deque(maxlen=4)
    for i in (990, 980, 970, 960):
        x2 = gd.p_sample_plms(x2, [i, i], 10, c2)
    assert rel_rmse(x2[:1].cpu(), g["plms"][3]) < 1e-3      # snapshot 3 = the reference's x after t = 960
    xb = T(g["x"])[1:2].cuda()
    gd.noise_list = deque(maxlen=4)
    for i in (990, 980, 970, 960):
        xb = gd.p_sample_plms(xb, [i], 10, T(g["cond"])[1:2].cuda())
    assert rel_rmse(x2[1:2].cpu(), xb.cpu()) < 1e-5


def test_diffnet_base_forward_c3_shape():
    g = load_golden("diffusion_base_fwd")
    gd = make(specs.DIFFNET_BASE, 2025)
    xb = specs.synth_tensor((2, 1, 80, 100), seed=31).cuda()
    cb = specs.synth_tensor((2, 256, 100), seed=32).cuda()
    eb = gd.denoise_fn(xb, [99, 3], cb).cpu()
    e = rel_rmse(eb, g["eps"])
    print("diffnet base (20x256) eps rel-RMSE:", e)
    assert e < TOL


@pytest.mark.parametrize("B,Tn", [(1, 1), (2, 5), (5, 131)])
def test_ragged_vs_oracle(B, Tn):
    from oracle import diffusion_ref as dr
    cfg = specs.DIFFNET_SMALL
    gd = make(cfg, 2024)
    sd = specs.synth_diffnet(cfg, 2024)
    x = specs.synth_tensor((B, 1, 80, Tn), seed=50 + Tn)
    cond = specs.synth_tensor((B, cfg["hidden_size"], Tn), seed=60 + Tn)
    t = [(7 * i + 3) % 100 for i in range(B)]
    ref = dr.diffnet_forward(sd, cfg, x, torch.tensor(t), cond)
    got = gd.denoise_fn(x.cuda(), t, cond.cuda()).cpu()
    assert rel_rmse(got, ref) < TOL
    tab = dr.schedule_tables(dr.linear_betas(100, 0.06))
    noise = specs.synth_tensor(tuple(x.shape), seed=70)
    refp = dr.p_sample(tab, lambda a, b, c: dr.diffnet_forward(sd, cfg, a, b, c), x, torch.tensor(t), cond, noise)
    gotp = gd._p_sample_core(x.cuda(), t, cond.cuda(), noise.cuda()).cpu()
    assert rel_rmse(gotp, refp) < TOL


def test_generic_denoise_fn_path():
    """Any callable denoise_fn (here: zero eps) goes through the eps-given entry."""
    gd = make(specs.DIFFNET_SMALL, 2024)
    class Zero(torch.nn.Module):
        def forward(self, x, t, cond=None):
            return torch.zeros_like(x)

    gd.denoise_fn = Zero()
    x = specs.synth_tensor((2, 1, 80, 9), seed=1).cuda()
    out = gd._p_sample_core(x, [5, 5], None, None)
    tb = gd._tables()
    exp = tb["c1"][5] * (tb["A"][5] * x).clamp(-1, 1) + tb["c2"][5] * x
    assert torch.allclose(out, exp, atol=1e-6)


def test_c3_full_chain_vs_reference(monkeypatch):
    """BASELINE configs[2] at FULL size -- DiffNet 20 x 256, B = 16, T = 400, all 100 ancestral steps with the
    per-step noise seeds of tests/golden/make_golden.py:golden_c3_full -- against the end point the reference's own
    GaussianDiffusion produced on CPU.  Stated tolerance: rel-RMSE <= 1e-3 on x_0 and on the de-normalised mel."""
    g = load_golden("diffusion_c3_full")
    gd = make(specs.DIFFNET_BASE, 2025)
    B, Tn = 16, 400
    x = specs.synth_tensor((B, 1, 80, Tn), seed=2).cuda()
    cond = specs.synth_tensor((B, 256, Tn), seed=3).cuda()
    calls = []

    def seeded(shape, device, repeat=False):          # same seeds, same call order (t = 99 .. 0) as the generator script
        i = 99 - len(calls)
        calls.append(i)
        return specs.synth_tensor((B, 1, 80, Tn), seed=4000 + i).to(device)

    monkeypatch.setattr(sdt, "noise_like", seeded)
    xl = gd.sample(cond, x_start=x)                    # on-device graph loop
    assert calls == list(range(99, -1, -1))
    e = rel_rmse(xl[:, :, :, ::4].cpu(), g["x_end"])
    mel = gd.denorm_spec(xl[:, 0].transpose(1, 2))
    em = rel_rmse(mel[:, ::4, :].cpu(), g["mel_end"])
    st = g["stats"]
    xd = xl.double()
    es = abs(float((xd * xd).sum()) - st[2]) / st[2]
    print("C3 full 100-step chain rel-RMSE x_0:", e, " mel:", em, " sum-of-squares rel:", es)
    assert e < 1e-3 and em < 1e-3 and es < 1e-3
    # the step-wise path (one p_sample call per step through the C ABI) must agree with the graph loop
    calls.clear()
    xs = x
    for i in reversed(range(100)):
        xs = gd.p_sample(xs, [i] * B, cond)
    print("graph loop vs step-wise:", rel_rmse(xl.cpu(), xs.cpu()))
    assert rel_rmse(xl.cpu(), xs.cpu()) < 1e-5


def test_cond_cache_survives_freed_source():
    """the hoisted conditioner cache is keyed on the caller's tensor; a transposed view is made
    contiguous for the engine, so the source block could be recycled for the next utterance (same shape,
    version 0) and the stale projection reused."""
    gd = make(specs.DIFFNET_SMALL, 2024)
    x = specs.synth_tensor((2, 1, 80, 33), seed=1).cuda()
    outs = []
    for seed in (1, 2, 3):
        dec = specs.synth_tensor((2, 33, specs.DIFFNET_SMALL["hidden_size"]), seed=seed).cuda()
        cond = dec.transpose(1, 2)                    # what GaussianDiffusion.forward hands over
        ref = cond.contiguous().clone()
        e = gd.denoise_fn(x, [7, 7], cond)
        del dec, cond
        assert torch.equal(e, gd.denoise_fn(x, [7, 7], ref))
        outs.append(e)
    assert not torch.equal(outs[0], outs[1])


def test_forward_infer_shallow_diffusion_vs_oracle(monkeypatch):
    """GaussianDiffusion.forward(infer=True) (shallow_diffusion_tts.py:248-277) end to end: a stub acoustic front-end
    supplies decoder_inp / mel_out, then the shallow-diffusion start x = q_sample(norm(fs2_mel), K-1), the K-step
    ancestral loop, denorm_spec and the mel2ph mask -- against the same composition on the CPU oracle with the SAME
    random draws (torch.randn_like and noise_like are pinned for the call).  Tolerance: rel-RMSE <= 1e-3 on the mel."""
    from oracle import diffusion_ref as dr
    cfg = specs.DIFFNET_SMALL
    K, B, Tn = 30, 2, 23
    gd = make(cfg, 2024)
    gd.K_step = K
    set_hparams_from_dict(dict(cfg, keep_bins=80, schedule_type="linear", max_beta=0.06, gaussian_start=False, pndm_speedup=None))
    dec = specs.synth_tensor((B, Tn, cfg["hidden_size"]), seed=81)
    fs2_mel = specs.synth_tensor((B, Tn, 80), seed=82, scale=1.0, shift=-2.5)
    mel2ph = torch.ones(B, Tn, dtype=torch.long)
    mel2ph[1, -5:] = 0                                  # padded tail of the second utterance

    class StubFS2(torch.nn.Module):
        def forward(self, txt_tokens, mel2ph=None, spk_embed=None, ref_mels=None, f0=None, uv=None, energy=None,
                    skip_decoder=False, infer=False, **kw):
            return {"decoder_inp": dec.cuda(), "mel_out": fs2_mel.cuda()}

    gd.fs2 = StubFS2()
    qn = specs.synth_tensor((B, 1, 80, Tn), seed=83)
    bank = specs.synth_tensor((K, B, 1, 80, Tn), seed=84)
    calls = []

    def pinned_noise(shape, device, repeat=False):
        i = K - 1 - len(calls)
        calls.append(i)
        return bank[i].to(device)

    monkeypatch.setattr(sdt, "noise_like", pinned_noise)
    monkeypatch.setattr(torch, "randn_like", lambda x, **kw: qn.to(x.device))
    ret = gd(torch.zeros(B, 5, dtype=torch.long).cuda(), mel2ph=mel2ph.cuda(), infer=True)
    monkeypatch.undo()
    assert calls == list(range(K - 1, -1, -1)) and ret["mel_out"].shape == (B, Tn, 80)
    assert torch.equal(ret["fs2_mel"].cpu(), fs2_mel)
    # oracle composition
    sd = specs.synth_diffnet(cfg, 2024)
    tab = dr.schedule_tables(dr.linear_betas(100, 0.06))
    smin, smax = T(specs.SPEC_MIN)[None, None], T(specs.SPEC_MAX)[None, None]
    cond = dec.transpose(1, 2)
    x = dr.q_sample(tab, dr.norm_spec(fs2_mel, smin, smax).transpose(1, 2)[:, None], torch.tensor([K - 1]), qn)
    fn = lambda a, b, c: dr.diffnet_forward(sd, cfg, a, b, c)
    for i in reversed(range(K)):
        x = dr.p_sample(tab, fn, x, torch.full((B,), i, dtype=torch.long), cond, bank[i])
    mel = dr.denorm_spec(x[:, 0].transpose(1, 2), smin, smax) * (mel2ph > 0).float()[:, :, None]
    e = rel_rmse(ret["mel_out"].cpu(), mel)
    print("forward(infer=True) mel rel-RMSE vs oracle:", e)
    assert e < 1e-3
    assert float(ret["mel_out"][1, -5:].abs().max()) == 0.0


def test_c3_full_size_properties():
    """BASELINE configs[2] shape: B=16, T=400, full DiffNet: batch independence + finite."""
    gd = make(specs.DIFFNET_BASE, 2025)
    x = specs.synth_tensor((16, 1, 80, 400), seed=2).cuda()
    cond = specs.synth_tensor((16, 256, 400), seed=3).cuda()
    e = gd.denoise_fn(x, [42] * 16, cond)
    assert torch.isfinite(e).all()
    e1 = gd.denoise_fn(x[5:6].contiguous(), [42], cond[5:6].contiguous())
    assert torch.allclose(e1[0], e[5], atol=1e-5, rtol=1e-5)
