"""Sanitizer workload for the tcgen05 attention kernel (attention_tc.cu) and the NSF source kernels: small shapes,
parity against torch in fp64 / the CPU oracle; exits non-zero on a failure."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib, specs
L = _lib.lib()
bad = 0
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1      # 1: fp32-input kernel, 2: plane-fed kernel
_lib.check(L.agpt_set_attention_tc(mode))
for N, heads, d, Lq, Lk in [(1, 2, 40, 130, 77), (1, 2, 80, 70, 130), (1, 1, 8, 5, 3), (1, 2, 64, 129, 64)]:
    Cc = heads * d
    q = specs.synth_tensor((N, Lq, Cc), seed=1).cuda()
    kv = specs.synth_tensor((N, Lk, 2 * Cc), seed=2).cuda()
    o = torch.empty((N, Lq, Cc), device="cuda")
    _lib.check(L.agpt_attention(_lib.fptr(q), Cc, _lib.fptr(kv), 2 * Cc, C.c_void_p(kv.data_ptr() + 4 * Cc), 2 * Cc,
                                _lib.fptr(o), Cc, N, heads, d, Lq, Lk, _lib.cur_stream()))
    torch.cuda.synchronize()
    qh = q.double().cpu().reshape(N, Lq, heads, d).permute(0, 2, 1, 3)
    kh = kv[:, :, :Cc].double().cpu().reshape(N, Lk, heads, d).permute(0, 2, 1, 3)
    vh = kv[:, :, Cc:].double().cpu().reshape(N, Lk, heads, d).permute(0, 2, 1, 3)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, dim=-1) @ vh).permute(0, 2, 1, 3).reshape(N, Lq, Cc)
    e = ((o.double().cpu() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    ok = e < 1e-5
    bad += 0 if ok else 1
    print(f"attention mode {mode} h={heads} d={d} {Lq}x{Lk}: rel-RMSE {e:.2e} {'ok' if ok else 'FAIL'}", flush=True)
sys.exit(1 if bad else 0)
