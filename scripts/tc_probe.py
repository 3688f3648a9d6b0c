"""Dev probe: validate the tcgen05 tapconv against the fp32-FMA tapconv and the CPU oracle.
Run each stage under `timeout` -- a wrong mbarrier phase hangs the kernel."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib, specs
from audiogpt_b200.modules.hifigan.hifigan import HifiGanGenerator

stage = sys.argv[1] if len(sys.argv) > 1 else "small"
L = _lib.lib()


def rmse(a, b):
    return (a.double() - b.double()).pow(2).mean().sqrt().item()


def build(h, seed=1234):
    m = HifiGanGenerator(h)
    m.load_state_dict(specs.synth_hifigan(h, seed))
    return m.eval().cuda()


if stage == "small":
    h = specs.HIFIGAN_SMALL
    m = build(h)
    mel = specs.synth_tensor((2, 80, 24), seed=11, scale=2.0, shift=-4.0).cuda()
    L.agpt_set_tensor_cores(0)
    ref = m(mel)
    torch.cuda.synchronize()
    print("fma done", flush=True)
    L.agpt_set_tensor_cores(1)
    out = m(mel)
    torch.cuda.synchronize()
    print("small: tc vs fma RMSE", rmse(out, ref), "max", (out - ref).abs().max().item(), flush=True)
elif stage == "v1":
    h = specs.HIFIGAN_V1
    m = build(h)
    B, T = int(sys.argv[2]) if len(sys.argv) > 2 else 1, 400
    mel = specs.synth_tensor((B, 80, T), seed=0, scale=2.0, shift=-4.0).cuda()
    L.agpt_set_tensor_cores(0)
    ref = m(mel)
    torch.cuda.synchronize()
    L.agpt_set_tensor_cores(1)
    out = m(mel)
    torch.cuda.synchronize()
    print("v1: tc vs fma RMSE", rmse(out, ref), "max", (out - ref).abs().max().item(), flush=True)
    for mode in (0, 1):
        L.agpt_set_tensor_cores(mode)
        for _ in range(2):
            m(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m(mel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"v1 B={B} mode={'tc' if mode else 'fma'}: {ms:.3f} ms  {245.64e9 * B / ms / 1e9:.1f} TFLOP/s  "
              f"{B * T * 256 / 22050 / (ms / 1e3):.0f}x RT", flush=True)
