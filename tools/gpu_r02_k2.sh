#!/bin/bash
# validation of the final round-2 build: smoke, GPU tests, ncu capture of the fused kernel, bench lines of the four workloads
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
bash tools/gpu_prof.sh r02m
for W in sot2048-nocut-sweep sot2048-cut; do
  timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-e2e --no-cpu 2>/dev/null | tail -1 >> gpurun_out/bench_r02m_workloads.jsonl
done
for W in sot512-cut sot512-logf-cut; do
  timeout 300 python bench.py --workload $W --frames 262144 --steps 20 --warmup 5 --no-e2e --no-cpu 2>/dev/null | tail -1 >> gpurun_out/bench_r02m_workloads.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/bench_r02m_workloads.jsonl"):
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["workload"], round(d["value"]/1e6,2),"Mframes/s step",round(d["ms_per_step"],4),"fused",round(r["kernel_ms"],4),"step frac",round(r["step"]["frac"],3),"kernel frac",round(r["frac"],3), "ref_on_cuda", (d.get("ref_on_cuda") or {}).get("value"))
PY
