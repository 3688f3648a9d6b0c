"""Dev probe: time the HiFi-GAN V1 forward on one GPU (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import specs, _lib
from audiogpt_b200.modules.hifigan.hifigan import HifiGanGenerator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 400
h = specs.HIFIGAN_V1
m = HifiGanGenerator(h); m.load_state_dict(specs.synth_hifigan(h, 1234)); m = m.eval().cuda()
mel = specs.synth_tensor((B, 80, T), seed=0, scale=2.0, shift=-4.0).cuda()
for _ in range(3): y = m(mel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n): y = m(mel)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 245.64e9 * B * T / 400
print(f"B={B} T={T}: {ms:.3f} ms/fwd  {B*T/ms*1e3:.0f} frames/s  {B*T*256/22050/(ms/1e3):.1f}x RT  {fl/ms/1e9:.2f} TFLOP/s  launches/fwd={_lib.launch_count()//(n+3)}")
