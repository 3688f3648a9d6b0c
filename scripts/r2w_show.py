import json
for g in ("none", "p2p", "nccl"):
    try:
        d = json.loads(open(f"gpurun_out/r2w_n4_{g}.json").read().strip().splitlines()[-1])
        print(g, round(d["ms_per_step"], 3), [round(x, 3) for x in d["ms_per_step_by_rank"]], d["clocks"])
    except Exception as e:
        print(g, "ERR", e)
