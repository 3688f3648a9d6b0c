"""Dev probe: per-layer tapconv micro-benchmarks, fp16 hi/lo (v5) vs tf32 hi/lo (v2) vs fp32 FMA."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib
L = _lib.lib()
torch.cuda.init(); torch.zeros(1).cuda()
shapes = [  # name, G, L, Cin, Cout, K, dil, Wreal
    ("hifi s0 k11", 8, 3200, 256, 256, 11, 1, 0), ("hifi s0 k3", 8, 3200, 256, 256, 3, 1, 0),
    ("hifi s1 k3", 8, 25600, 128, 128, 3, 1, 0),
    ("hifi s1 k11", 8, 25600, 128, 128, 11, 5, 0), ("hifi s2 k11", 8, 51200, 64, 64, 11, 1, 0),
    ("hifi s2 k3", 8, 51200, 64, 64, 3, 1, 0),
    ("hifi s3 k11", 8, 102400, 32, 32, 11, 1, 0), ("hifi s3 k3", 8, 102400, 32, 32, 3, 1, 0),
    ("diffnet dil", 16, 400, 256, 512, 3, 1, 0), ("diffnet out1x1", 16, 400, 256, 512, 1, 1, 0),
    ("unet lin 320", 1, 6240, 320, 320, 1, 1, 0), ("unet ff1", 1, 6240, 320, 2560, 1, 1, 0),
    ("unet ff2", 1, 6240, 1280, 320, 1, 1, 0), ("unet conv 320", 8, 780, 320, 320, 3, 1, 78),
    ("unet conv 640@5x39", 8, 195, 640, 640, 3, 1, 39), ("unet conv 1280->640", 8, 195, 1280, 640, 3, 1, 39),
    # tap-shift alignment probe: dilation 8 makes every tap's row shift a multiple of the 8-row swizzle atom
    ("probe k11 d1", 8, 3200, 256, 256, 11, 1, 0), ("probe k11 d8", 8, 3200, 256, 256, 11, 8, 0),
    ("probe k3 d1 128", 8, 25600, 128, 128, 3, 1, 0), ("probe k3 d8 128", 8, 25600, 128, 128, 3, 8, 0),
]
sel = sys.argv[1:]
for name, G, Ln, Cin, Cout, K, dil, Wr in shapes:
    if sel and not any(s in name for s in sel): continue
    res = {}
    VA, VB = int(os.environ.get('VA', 7)), int(os.environ.get('VB', 5))
    for ver in (VA, VB):
        _lib.check(L.agpt_set_tc_version(ver))
        out = (C.c_double * 3)(); dbg = (C.c_double * 8)()
        _lib.check(L.agpt_bench_tapconv(G, Ln, Cin, Cout, K, dil, Wr, 1, 1, 5, 1, out, dbg))
        res[ver] = (list(out), list(dbg))
    o2 = (C.c_double * 3)()
    _lib.check(L.agpt_bench_tapconv(G, Ln, Cin, Cout, K, dil, Wr, 1, 0, 3, 0, o2, None))
    o5, d = res[VA]; ov2, dv2 = res[VB]
    print(f"{name:22s} vA {o5[0]*1e3:8.1f} us {o5[1]:6.1f} TF d={o5[2]:.1e} | vB {ov2[0]*1e3:8.1f} us {ov2[1]:6.1f} TF d={ov2[2]:.1e} | fma {o2[0]*1e3:8.1f} us | "
          f"vA dbg " + " ".join(f"{x:.0f}" for x in d) + " | vB dbg " + " ".join(f"{x:.0f}" for x in dv2), flush=True)
_lib.check(L.agpt_set_tc_version(-1))
