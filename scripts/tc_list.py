import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib, specs
from audiogpt_b200.modules.hifigan.hifigan import HifiGanGenerator
h = specs.HIFIGAN_V1
m = HifiGanGenerator(h); m.load_state_dict(specs.synth_hifigan(h, 1234)); m = m.eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mel = specs.synth_tensor((B, 80, 400), seed=0, scale=2.0, shift=-4.0).cuda()
_lib.lib().agpt_set_tensor_cores(1)
m(mel); torch.cuda.synchronize()
m(mel); torch.cuda.synchronize()
