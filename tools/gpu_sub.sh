#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "two_frames_per_warp or papers_batch" 2>&1 | tail -8 | cut -c1-300
for T in "" "--tuning 16,17,1"; do
for W in sot512-cut sot512-logf-cut; do
  python bench.py --workload $W --frames 262144 --no-e2e --no-cpu --no-ref-cuda $T 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('$W', '$T', round(d['value']/1e6,2), 'Mframes/s step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'fwd', round(r['forward_only_kernel']['ms'],4), 'loss', d['loss'])"
done; done
