#!/bin/bash
# A/B of a kernel change: GPU tests + bench lines of the four workloads (no comparators)
TAG=${1:-r02p}
mkdir -p gpurun_out; rm -f gpurun_out/bench_${TAG}_workloads.jsonl
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
for W in sot2048-nocut-sweep sot2048-cut; do
  timeout 300 python bench.py --workload $W --steps 50 --warmup 5 --no-e2e --no-cpu --no-ref-cuda 2>/dev/null | tail -1 >> gpurun_out/bench_${TAG}_workloads.jsonl
done
for W in sot512-cut sot512-logf-cut; do
  timeout 300 python bench.py --workload $W --frames 262144 --steps 50 --warmup 5 --no-e2e --no-cpu --no-ref-cuda 2>/dev/null | tail -1 >> gpurun_out/bench_${TAG}_workloads.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/bench_${TAG}_workloads.jsonl"):
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["workload"], round(d["value"]/1e6,2),"Mframes/s step",round(d["ms_per_step"],4),"fused",round(r["kernel_ms"],4),"step frac",round(r["step"]["frac"],3),"kernel frac",round(r["frac"],3))
PY
