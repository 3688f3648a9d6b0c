import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import specs
from audiogpt_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
cfg = specs.UNET_TXT2AUDIO
u = UNetModel(image_size=32, use_checkpoint=True, **cfg)
u.load_state_dict(specs.synth_unet(cfg, 4040), strict=True)
u = u.eval().cuda()
N = 8
x = specs.synth_tensor((N, 4, 10, 78), seed=9).cuda()
ctx = specs.synth_tensor((N, 77, 1024), seed=10).cuda()
for _ in range(3):
    u(x, timesteps=[501] * N, context=ctx)
torch.cuda.synchronize()
