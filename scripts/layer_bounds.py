"""Per-layer ceilings for a layer table written by scripts/layer_profile.py (no GPU needed).

For every tap-GEMM launch: measured time, the tensor-pipe floor of the fp16-hi/lo kernel and the HBM floor.
  MMA floor : row_tiles * co_tiles * chunks * taps * ksteps * 3 MMAs, each >= max(BN / 2, A_FETCH) cycles
              (128 x BN x 16 per instruction; in SS mode the 4 KB A slice is re-read from shared memory for every
              instruction, ~77 cycles measured whatever BN <= 128 is -- profiles/r1e_findings.md), spread over 148 SMs.
  HBM floor : (input + output [+ residual] [+ old accumulator]) * 4 B / measured copy bandwidth.
usage: layer_bounds.py profiles/r1e_layers_hifigan_B8_T800.txt"""
import json, math, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
A_FETCH, SMS, GHZ = 77.0, 148, 1.965
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) * 1e9
except Exception:
    HBM = 6.65e12


def pick_bn(cout):
    if cout <= 32: return 32
    if cout <= 64: return 64
    return 128 if math.ceil(cout / 128) * 128 <= math.ceil(cout / 64) * 64 else 64


pat = re.compile(r"G=\s*(\d+) L=\s*(\d+) Cin=\s*(\d+) Cout=\s*(\d+) taps=\s*(\d+) epi=\s*(\d+) W=\s*(\d+)\s+n=\s*(\d+)\s+([\d.]+) us")
rows, tot, tot_m, tot_h = [], 0.0, 0.0, 0.0
for line in open(sys.argv[1]):
    m = pat.search(line)
    if not m:
        continue
    G, L, Cin, Cout, taps, epi, W, n, us = [float(v) if i == 8 else int(v) for i, v in enumerate(m.groups())]
    us /= n
    bn = pick_bn(Cout)
    Lv = (L // W) * (W + 1) if W else L
    tiles = math.ceil(Lv / 128) * G * math.ceil(Cout / bn)
    mmas = 0
    for c in range(math.ceil(Cin / 64)):
        kv = min(64, Cin - 64 * c)
        mmas += taps * math.ceil(kv / 16) * 3
    cyc = max(bn / 2.0, A_FETCH)
    t_mma = math.ceil(tiles / SMS) * mmas * cyc / (GHZ * 1e3)          # us
    out_c = Cout // 2 if epi in (5, 6) else Cout
    elems = G * L * (Cin + out_c + (Cout if epi in (1, 2, 5, 6) else 0) + (Cout if epi == 2 else 0))
    t_hbm = elems * 4 / HBM * 1e6
    rows.append((n * us, G, L, Cin, Cout, taps, epi, n, us, t_mma, t_hbm))
    tot += n * us; tot_m += n * t_mma; tot_h += n * max(t_hbm, 0)
print(f"{'layer':44s} {'n':>3s} {'meas us':>9s} {'MMA floor':>10s} {'HBM floor':>10s} {'meas/ceil':>9s}")
for _, G, L, Cin, Cout, taps, epi, n, us, tm, th in sorted(rows, reverse=True):
    ceil_ = max(tm, th)
    print(f"G={G:2d} L={L:7d} {Cin:4d}->{Cout:4d} k={taps:2d} epi={epi:2d}        {n:3d} {us:9.1f} {tm:10.1f} {th:10.1f} {us / ceil_:9.2f}")
print(f"total measured {tot / 1e3:.2f} ms; sum of MMA floors {tot_m / 1e3:.2f} ms; sum of HBM floors {tot_h / 1e3:.2f} ms; "
      f"sum of max(floors) {sum(r[7] * max(r[9], r[10]) for r in rows) / 1e3:.2f} ms")
