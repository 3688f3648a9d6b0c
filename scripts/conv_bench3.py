"""Dev probe: v3 (persistent) tapconv wait accounting."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib
L = _lib.lib(); torch.zeros(1).cuda()
shapes = [("hifi s1 k3", 8, 25600, 128, 128, 3, 1, 0), ("hifi s1 k11", 8, 25600, 128, 128, 11, 5, 0),
          ("hifi s3 k3", 8, 102400, 32, 32, 3, 1, 0), ("unet lin 320", 1, 6240, 320, 320, 1, 1, 0),
          ("unet ff2", 1, 6240, 1280, 320, 1, 1, 0), ("diffnet dil", 16, 400, 256, 512, 3, 1, 0)]
for name, G, Ln, Cin, Cout, K, dil, Wr in shapes:
    out = (C.c_double * 3)(); dbg = (C.c_double * 8)()
    _lib.check(L.agpt_bench_tapconv(G, Ln, Cin, Cout, K, dil, Wr, 1, 1, 5, 0, out, dbg))
    d = list(dbg)
    print(f"{name:14s} {out[0]*1e3:8.1f} us {out[1]:6.1f} TF | total {d[0]:.0f} mmaWaitA {d[1]:.0f} mmaWaitW {d[2]:.0f} mmaWaitAcc {d[3]:.0f} "
          f"xfWaitAempty {d[4]:.0f} epiWaitAcc {d[5]:.0f} epiBusy {d[6]:.0f} prodWait {d[7]:.0f}", flush=True)
