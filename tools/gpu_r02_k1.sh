#!/bin/bash
# validation of the one-warp-per-frame default for 1025 bins: smoke, GPU tests, bench lines, ncu capture
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
bash tools/gpu_prof.sh r02k
timeout 300 python bench.py --workload sot2048-cut --steps 20 --warmup 5 --no-e2e --no-cpu --no-ref-cuda 2>/dev/null | tail -1 > gpurun_out/bench_r02k_sot2048-cut.json
python -c "
import json
d=json.load(open('gpurun_out/bench_r02k_sot2048-cut.json')); r=d['roofline']
print('cut', round(d['value']/1e6,2),'Mframes/s step',round(d['ms_per_step'],4),'fused',round(r['kernel_ms'],4),'step frac',round(r['step']['frac'],3),'kernel frac',round(r['frac'],3))"
