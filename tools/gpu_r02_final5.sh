#!/bin/bash
# final round-2 build: GPU tests, the driver's default bench line (timed), launch list, paper-batch latency lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
T0=$(date +%s)
timeout 600 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_r02n_default.json
echo "default bench: $(( $(date +%s) - T0 )) s"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02n_default.json")); r=d["roofline"]
print(round(d["value"]/1e6,2),"Mframes/s step",round(d["ms_per_step"],4),"kernel frac",round(r["frac"],3),"step frac",round(r["step"]["frac"],3),"e2e",d["e2e"]["value"],"cpu",d["cpu_baseline"]["value"],"ref_on_cuda",(d.get("ref_on_cuda") or {}).get("value"),"clocks",d.get("clocks"))
PY
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/r02n_launches.csv $BENCH > gpurun_out/r02n_ncu_launch.log 2>&1; echo "launch-list exit $?"
timeout 200 python bench.py --workload sot2048-cut --frames 1024 --steps 200 --warmup 20 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r02n_latency_eager.json
timeout 200 python bench.py --workload sot2048-cut --frames 1024 --steps 200 --warmup 20 --no-e2e --no-cpu --no-ref-cuda --graph 2>/dev/null | tail -1 > gpurun_out/r02n_latency_graph.json
python - <<PY
import json
for f in ("eager","graph"):
    d=json.load(open(f"gpurun_out/r02n_latency_{f}.json")); print(f, round(d["ms_per_step"]*1e3,1), "us/step", (d.get("ref_on_cuda") or {}).get("ms_per_step"))
PY
