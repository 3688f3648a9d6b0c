"""Dev probe / sanitizer workload for the plane-fed tap-GEMM (tcconv7.cu): parity against the fp32-FMA kernel
(fp32 result and emitted planes) on 1-D layer shapes, then timing against the default schedule on the HiFi-GAN shapes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiogpt_b200 import _lib
L = _lib.lib()
torch.zeros(1).cuda()
CHECK = [(2, 3000, 256, 256, 11, 5), (2, 5000, 128, 128, 3, 3), (2, 9000, 32, 32, 7, 1), (3, 400, 256, 512, 3, 2),
         (1, 777, 320, 320, 1, 1), (1, 130, 1280, 320, 1, 1), (2, 4, 64, 96, 3, 1), (2, 300, 80, 256, 7, 1),
         (2, 500, 96, 40, 5, 2), (1, 100000, 32, 32, 3, 1), (1, 50000, 64, 64, 11, 1)]
bad = 0
_lib.check(L.agpt_set_tc_version(8))
for G, Ln, Cin, Cout, K, dil in CHECK:
    for res in (0, 1):
        rel = (C.c_double * 2)()
        rc = L.agpt_check_tapconv(G, Ln, Cin, Cout, K, dil, 0, res, C.c_double(1.0), C.c_double(1.0), rel)
        if rc != 0:
            print(f"planes G={G} L={Ln} {Cin}->{Cout} k={K} d={dil} res={res}: REJECTED/ERROR {L.agpt_last_error().decode()}", flush=True)
            bad += 1
            continue
        ok = rel[0] < 2e-4 and rel[1] < 2e-5
        bad += 0 if ok else 1
        print(f"planes G={G} L={Ln} {Cin}->{Cout} k={K} d={dil} res={res}: max/rms {rel[0]:.2e} rms/rms {rel[1]:.2e} {'ok' if ok else 'FAIL'}", flush=True)
if "--time" in sys.argv:
    SH = [("s0 k3", 8, 6400, 256, 256, 3, 1), ("s0 k11", 8, 6400, 256, 256, 11, 1), ("s1 k3", 8, 51200, 128, 128, 3, 1),
          ("s1 k7", 8, 51200, 128, 128, 7, 3), ("s1 k11", 8, 51200, 128, 128, 11, 5), ("s2 k3", 8, 102400, 64, 64, 3, 1),
          ("s2 k11", 8, 102400, 64, 64, 11, 1), ("s3 k3", 8, 204800, 32, 32, 3, 1), ("s3 k7", 8, 204800, 32, 32, 7, 1),
          ("s3 k11", 8, 204800, 32, 32, 11, 1)]
    for name, G, Ln, Cin, Cout, K, dil in SH:
        row = []
        for ver in (8, 6):
            _lib.check(L.agpt_set_tc_version(ver))
            for res in (0, 1):
                out = (C.c_double * 3)()
                _lib.check(L.agpt_bench_tapconv(G, Ln, Cin, Cout, K, dil, 0, res, 1, 5, 0, out, None))
                row.append((out[0] * 1e3, out[1]))
        print(f"{name:8s} planes: {row[0][0]:7.1f} us {row[0][1]:6.1f} TF | +res {row[1][0]:7.1f} us {row[1][1]:6.1f} TF || v6: {row[2][0]:7.1f} us {row[2][1]:6.1f} TF | +res {row[3][0]:7.1f} us {row[3][1]:6.1f} TF", flush=True)
_lib.check(L.agpt_set_tc_version(-1))
sys.exit(1 if bad else 0)
