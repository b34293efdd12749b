#!/bin/bash
# evidence of the FINAL round-2 build: smoke, ncu full capture + launch list of the fused kernel, the driver's default bench line
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02s -f $BENCH > gpurun_out/ncu_full_r02s.log 2>&1; echo "full exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/r02s_launches.csv $BENCH > gpurun_out/r02s_ncu_launch.log 2>&1; echo "launch-list exit $?"
timeout 600 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_r02s_default.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02s_default.json")); r=d["roofline"]
print(round(d["value"]/1e6,2),"Mframes/s step",round(d["ms_per_step"],4),"kernel",round(r["kernel_ms"],4),"kernel frac",round(r["frac"],3),"step frac",round(r["step"]["frac"],3),"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],"ref_on_cuda",(d.get("ref_on_cuda") or {}).get("value"))
PY
