"""Dev probe: DDIM chain timing on the C4 shard.  usage: ddim_quick.py [S=100] [B=4] [reps=2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audiogpt_b200 import _lib, specs
from audiogpt_b200.ldm.models.diffusion.ddim import DDIMSampler, LatentDiffusionShim
from audiogpt_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = specs.UNET_TXT2AUDIO
u = UNetModel(image_size=32, use_checkpoint=True, **cfg)
u.load_state_dict(specs.synth_unet(cfg, 4040), strict=True)
u = u.eval().cuda()
smp = DDIMSampler(LatentDiffusionShim(u).cuda())
xT = torch.tensor(np.random.RandomState(55).randn(B, 4, 10, 78), dtype=torch.float32).cuda()
c = specs.synth_tensor((B, 77, 1024), seed=5).cuda()
uc = specs.synth_tensor((1, 77, 1024), seed=6).expand(B, -1, -1).contiguous().cuda()
kw = dict(S=S, batch_size=B, shape=(4, 10, 78), conditioning=c, verbose=False, x_T=xT, eta=0.0,
          unconditional_guidance_scale=1.5, unconditional_conditioning=uc)
smp.sample(**kw)
torch.cuda.synchronize()
for r in range(reps):
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    z, _ = smp.sample(**kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"DDIM-{S} B={B}: {dt*1e3:.1f} ms  {dt/S*1e3:.3f} ms/step  {B/dt*S/100:.2f} clips/s (DDIM-100 equiv)  "
          f"{18.66*B*S/100/dt:.1f} TF alg  launches {_lib.launch_count()-l0}  finite {bool(torch.isfinite(z).all())}", flush=True)
