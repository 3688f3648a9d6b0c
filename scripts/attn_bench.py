"""Dev probe: attention kernels on the UNet's shapes, CUDA-event timing.  Modes: 1 fp32-input tcgen05, 2 plane-fed tcgen05
(+ 3 plane-split launches on this route), 0 fp32 FMA."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib, specs
L = _lib.lib()
torch.zeros(1).cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for N, heads, d, Lq, Lk in [(8, 8, 40, 780, 780), (8, 8, 40, 780, 77), (8, 8, 80, 195, 195), (8, 8, 80, 195, 77)]:
    C_ = heads * d
    q = specs.synth_tensor((N, Lq, C_), seed=1).cuda()
    kv = specs.synth_tensor((N, Lk, 2 * C_), seed=2).cuda()
    o = torch.empty((N, Lq, C_), device="cuda")
    res = []
    for mode in (1, 2, 0):
        _lib.check(L.agpt_set_attention_tc(mode))
        def run():
            _lib.check(L.agpt_attention(_lib.fptr(q), C_, _lib.fptr(kv), 2 * C_, C.c_void_p(kv.data_ptr() + 4 * C_), 2 * C_,
                                        _lib.fptr(o), C_, N, heads, d, Lq, Lk, _lib.cur_stream()))
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / reps * 1e3)
    _lib.check(L.agpt_set_attention_tc(-1))
    fl = 4.0 * N * heads * Lq * Lk * d
    print(f"attention N={N} h={heads} d={d} {Lq}x{Lk}: tcgen05 {res[0]:.1f} us ({fl/res[0]*1e-6:.1f} TF)  plane-fed(+3 splits) {res[1]:.1f} us  fma {res[2]:.1f} us", flush=True)
