#!/bin/bash
# compute-sanitizer over the tap-GEMM schedules, the plane-fed kernel and the attention kernels.
# usage: scripts/run_sanitizer.sh <summary.txt>   (per-tool logs go to gpurun_out/san_*.log)
out=${1:-gpurun_out/sanitizer.txt}
mkdir -p gpurun_out
echo "compute-sanitizer $(compute-sanitizer --version 2>/dev/null | tail -1) on $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1); workloads: scripts/sanitize_shapes.py <schedule> (layer shapes x {bias,residual}; 5 = one tile per CTA incl. the 96-wide tile, 7 = persistent kernel forced, 8 = plane-fed kernel incl. 64/96-wide tiles), scripts/sanitize_attn.py <mode> (1 fp32-input, 2 plane-fed)" > $out
run() {
  name=$1; shift
  timeout 900 compute-sanitizer "$@" > gpurun_out/san_$name.log 2>&1
  rc=$?
  echo "$name rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$name.log | tail -1)" >> $out
}
run memcheck_v5 --tool memcheck python scripts/sanitize_shapes.py 5
run memcheck_v7 --tool memcheck python scripts/sanitize_shapes.py 7
run racecheck_v7 --tool racecheck python scripts/sanitize_shapes.py 7
run memcheck_v8 --tool memcheck python scripts/sanitize_shapes.py 8
run racecheck_v8 --tool racecheck python scripts/sanitize_shapes.py 8
run memcheck_attn1 --tool memcheck python scripts/sanitize_attn.py 1
run memcheck_attn2 --tool memcheck python scripts/sanitize_attn.py 2
run racecheck_attn2 --tool racecheck python scripts/sanitize_attn.py 2
for n in memcheck_v5 memcheck_v8 memcheck_attn2; do echo "--- $n" >> $out; grep -E " ok$| FAIL$" gpurun_out/san_$n.log >> $out; done
cat $out | head -12
