#!/bin/bash
# bench every library under _lib/variants (plus the default build) on the headline workload
PKG=1d-spectral-optimal-transport_b200
for lib in default $PKG/_lib/variants/*.so; do
  if [ "$lib" = default ]; then unset SOT_B200_LIBRARY; else export SOT_B200_LIBRARY=$PWD/$lib; fi
  python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$lib', round(d['value']/1e6, 2), 'Mframes/s step', round(d['ms_per_step'], 4), 'bwd', round(r['kernel_ms'], 4), 'fwd', round(r['forward_only_kernel']['ms'], 4))"
done
