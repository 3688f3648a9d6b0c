"""Dev probe: DiffSinger C3 chain timing (16 utt x 400 frames x 100 ancestral steps on the device loop)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib, specs
from audiogpt_b200.modules.diff import shallow_diffusion_tts as sdt
from audiogpt_b200.modules.diff.net import DiffNet
from audiogpt_b200.utils.hparams import set_hparams_from_dict
cfgd = specs.DIFFNET_BASE
set_hparams_from_dict(dict(cfgd, keep_bins=80, schedule_type="linear", max_beta=0.06))
net = DiffNet(80)
net.load_state_dict(specs.synth_diffnet(cfgd, 2025), strict=True)
gd = sdt.GaussianDiffusion(None, 80, net, timesteps=100, K_step=100, loss_type="l1", betas=sdt.linear_beta_schedule(100, 0.06),
                           spec_min=specs.SPEC_MIN, spec_max=specs.SPEC_MAX).eval().cuda()
Bc, Tc = 16, 400
x = specs.synth_tensor((Bc, 1, 80, Tc), seed=2).cuda()
cond = specs.synth_tensor((Bc, 256, Tc), seed=3).cuda()
gd.sample(cond, x_start=x)
torch.cuda.synchronize()
for r in range(3):
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    y = gd.sample(cond, x_start=x)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"C3 16x400x100: {dt*1e3:.1f} ms  {dt*10:.3f} ms/step  {26.44e6*Bc*Tc*100/dt/1e12:.1f} TF alg  launches {_lib.launch_count()-l0}  finite {bool(torch.isfinite(y).all())}", flush=True)
