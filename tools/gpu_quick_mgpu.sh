#!/bin/bash
# usage: tools/gpu_quick_mgpu.sh <N> [check]: weak-scaling line on N GPUs (+ the exchange checks when asked)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$2" = check ]; then
timeout 240 $TR --master-port 29511 tools/p2p_check.py > gpurun_out/p2p_check_$N.log 2>&1; echo "p2p_check exit $?"; tail -1 gpurun_out/p2p_check_$N.log | cut -c1-300
fi
for tag in weak strong_graph; do
  EXTRA=""; if [ $tag = strong_graph ]; then EXTRA="--scaling strong --graph"; fi
  timeout 200 $TR --master-port 29512 bench.py --gpus $N --no-e2e --no-ref-cuda $EXTRA > gpurun_out/r02_bench_${tag}_quick_$N.log 2>&1
  tail -1 gpurun_out/r02_bench_${tag}_quick_$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$tag', d['n_gpus'], 'gpus', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4), 'ms/step', d['collective']['used'], 'check', d['value_check'] and d['value_check']['rel_err'])"
done
python bench.py --no-cpu --no-ref-cuda --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('single', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4))"
