#!/bin/bash
# Short 8-GPU session: the scaling lines only (weak as the driver runs it, strong eager / graph) + one GPU of the box.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag, args...
  tag=$1; shift
  timeout 240 $TR --master-port 29512 bench.py --gpus $N "$@" > gpurun_out/r02_bench_${tag}_$N.log 2>&1
  echo "$tag exit $?"; tail -1 gpurun_out/r02_bench_${tag}_$N.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$tag', d['n_gpus'], 'gpus', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4), 'ms/step', d['collective']['used'], [round(t,4) for t in d.get('ms_per_step_by_rank')], d.get('host_issue'), 'check', d['value_check'] and d['value_check']['rel_err'], 'e2e', d['e2e'] and round(d['e2e']['value']/1e6,2), d['clocks'])
except Exception as e: print('$tag parse failed', e)"
}
run weak_full
run strong_graph --no-e2e --no-ref-cuda --scaling strong --graph
run strong_eager --no-e2e --no-ref-cuda --scaling strong
python bench.py --no-cpu --no-ref-cuda --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02_bench_single_on_$N.json; python -c "
import json
d = json.load(open('gpurun_out/r02_bench_single_on_$N.json')); print('single', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4))"
