#!/bin/bash
# the host-buffer entry point with and without SOT_HOST_ZEROCOPY: parity tests, then the e2e leg of the bench line
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_buffer" -p no:cacheprovider 2>&1 | tail -4
for Z in 0 1; do
  SOT_HOST_ZEROCOPY=$Z timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu --no-ref-cuda 2>gpurun_out/zc_$Z.err | tail -1 > gpurun_out/bench_zerocopy_$Z.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_zerocopy_$Z.json')); print('zerocopy=$Z e2e', round(d['e2e']['value']/1e6,3), 'M frames/s', d['e2e']['api'][-40:], 'step', round(d['value']/1e6,1))" || tail -3 gpurun_out/zc_$Z.err
done
