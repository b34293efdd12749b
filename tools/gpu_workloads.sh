#!/bin/bash
# bench lines for the other BASELINE configurations (1 GPU): SOT-512, SOT-512-LogF, SOT-2048 cut, paper batch
mkdir -p gpurun_out
TAG=${1:-r02}
for W in sot512-cut sot512-logf-cut sot2048-cut; do
  FR=65536; if [ $W != sot2048-cut ]; then FR=262144; fi
  timeout 300 python bench.py --workload $W --frames $FR --steps 20 --warmup 5 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_${TAG}_$W.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_$W.json")); r=d["roofline"]
print("$W", round(d["value"]/1e6,2),"Mframes/s step",round(d["ms_per_step"],4),"fused",round(r["kernel_ms"],4),"step frac",round(r["step"]["frac"],3),"kernel frac",round(r["frac"],3), "ref_on_cuda", d["ref_on_cuda"] and round(d["ref_on_cuda"]["value"]/1e6,3))
PY
done
