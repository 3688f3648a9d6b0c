"""Key metrics of an ncu report as JSON.  usage: ncu_extract.py <report.ncu-rep> "<command line / description>" > out.json"""
import csv, json, subprocess, sys
rep, desc = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
units = rows[1]
out = []
for r in rows[2:]:
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i] + (" " + units[i] if units[i] else "")).strip()
    out.append(d)
print(json.dumps({"command": desc, "launches": out}, indent=1))
