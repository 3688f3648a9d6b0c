import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "x_realtime", "gpu_launches")})
print("e2e", d["e2e"]["value"])
r = d["roofline"]
print("roofline", {k: r.get(k) for k in ("bound", "achieved", "peak", "frac", "achieved_algorithmic_tflops", "share_of_step")})
print(r["per_variant"]); print("hbm", r["hbm"]["achieved"], r["hbm"]["frac"])
print(d["cpu_baseline"]); print(d["extra"]); print(d["clocks"])
