import sys, os
sys.path.insert(0, os.getcwd())
import torch
from audiogpt_b200 import specs
from audiogpt_b200.vocoder.bigvgan.models import BigVGAN
hb = specs.BIGVGAN_BASE
bv = BigVGAN(hb); bv.load_state_dict(specs.synth_bigvgan(hb, 4321), strict=True); bv = bv.eval().cuda()
mel = specs.synth_tensor((8, 80, 400), seed=9, scale=2.0, shift=-4.0).cuda()
for _ in range(2): bv(mel)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): w = bv(mel)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"bigvgan base 8x400: {ms:.2f} ms  {8*400/ms*1e3:.0f} frames/s  finite={bool(torch.isfinite(w).all())} rms={w.pow(2).mean().sqrt().item():.3f}")
