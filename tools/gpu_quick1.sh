#!/bin/bash
# quick single-GPU check: parity tests + headline and short-row bench numbers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_complex.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for W in sot2048-nocut-sweep sot512-cut; do
  FR=65536; if [ $W = sot512-cut ]; then FR=262144; fi
  python bench.py --workload $W --frames $FR --no-e2e --no-cpu --no-ref-cuda 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('$W', round(d['value']/1e6,2), 'Mframes/s step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'fwd', round(r['forward_only_kernel']['ms'],4))"
done
