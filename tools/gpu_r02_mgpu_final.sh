#!/bin/bash
# usage: tools/gpu_r02_mgpu_final.sh <N>: exchange checks + weak / strong-graph scaling lines of the final build on N GPUs
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/p2p_check.py > gpurun_out/p2p_check_final_$N.log 2>&1; echo "p2p_check exit $?"; tail -1 gpurun_out/p2p_check_final_$N.log | cut -c1-400
timeout 200 python -m pytest tests/test_gpu_mss.py -m gpu -q -k "peer_memory or late or sharded" -p no:cacheprovider 2>&1 | tail -2
for tag in weak strong_graph; do
  EXTRA=""; if [ $tag = strong_graph ]; then EXTRA="--scaling strong --graph"; fi
  timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-e2e --no-cpu --no-ref-cuda $EXTRA > gpurun_out/r02_bench_${tag}_final_$N.log 2>&1
  tail -1 gpurun_out/r02_bench_${tag}_final_$N.log > gpurun_out/r02_bench_${tag}_final_$N.json
  python -c "
import sys, json
d = json.load(open('gpurun_out/r02_bench_${tag}_final_$N.json')); print('$tag', d['n_gpus'], 'gpus', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4), 'ms/step', d['collective']['used'], 'check', d['value_check'] and d['value_check']['rel_err'])"
done
python bench.py --steps 50 --warmup 5 --no-cpu --no-ref-cuda --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('single', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4))"
