#!/bin/bash
# usage: tools/gpu_ab.sh <workload> <frames> <tuning> [<tuning> ...]: A/B of kernel configurations ("default" = none)
W=$1; FR=$2; shift; shift
for T in "$@"; do
  ARG=""; if [ "$T" != default ]; then ARG="--tuning $T"; fi
  python bench.py --workload $W --frames $FR --no-e2e --no-cpu --no-ref-cuda $ARG 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('$W', '$T', round(d['value']/1e6,2), 'Mframes/s step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'fwd', round(r['forward_only_kernel']['ms'],4))"
done
