#!/bin/bash
# usage: tools/gpu_prof.sh <tag> [bench args]: ncu full capture of the fused kernel + default bench line
TAG=$1; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda $*"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f $BENCH > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-ref-cuda $* 2>/dev/null | tail -1 > gpurun_out/bench_$TAG.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json")); r=d["roofline"]
print(round(d["value"]/1e6,2),"Mframes/s step",round(d["ms_per_step"],4),"fused",round(r["kernel_ms"],4),"fwd",round(r["forward_only_kernel"]["ms"],4),"step frac",round(r["step"]["frac"],3),"kernel frac",round(r["frac"],3))
PY
