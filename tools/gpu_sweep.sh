#!/bin/bash
# kernel-configuration sweep on the bench workload (kernel-only timings come from the roofline block)
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for t in "128,9" "64,17" "32,33" "128,17"; do
  for w in sot2048-nocut-sweep sot2048-cut; do
    timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --tuning $t --workload $w 2>>gpurun_out/sweep.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print(json.dumps({'tuning': '$t', 'workload': '$w', 'ms_step': d['ms_per_step'], 'bwd_ms': r['kernel_ms'], 'fwd_ms': r['forward_kernel']['ms'], 'frac_bwd': r['frac'], 'frac_fwd': r['forward_kernel']['frac'], 'frac_step': r['step']['frac']}))" | tee -a gpurun_out/sweep.jsonl
  done
done
for w in sot512-cut sot512-logf-cut; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $w --frames 262144 2>>gpurun_out/sweep.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print(json.dumps({'workload': '$w', 'ms_step': d['ms_per_step'], 'bwd_ms': r['kernel_ms'], 'fwd_ms': r['forward_kernel']['ms'], 'frac_bwd': r['frac'], 'frac_fwd': r['forward_kernel']['frac'], 'frac_step': r['step']['frac'], 'value': d['value']}))" | tee -a gpurun_out/sweep.jsonl
done
