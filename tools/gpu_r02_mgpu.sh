#!/bin/bash
# usage: tools/gpu_r02_mgpu.sh <N>: multi-GPU checks and scaling lines on N GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/p2p_check.py > gpurun_out/p2p_check_$N.log 2>&1; echo "p2p_check exit $?"; tail -2 gpurun_out/p2p_check_$N.log | cut -c1-1800
run() {  # tag, args...
  tag=$1; shift
  timeout 150 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-ref-cuda "$@" > gpurun_out/bench_${tag}_$N.log 2>&1
  echo "$tag exit $?"; tail -1 gpurun_out/bench_${tag}_$N.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$tag', d['n_gpus'], 'gpus', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4), 'ms/step', d['collective']['used'], d.get('ms_per_step_by_rank'), d.get('host_issue'), 'check', d['value_check'] and d['value_check']['rel_err'], 'e2e', d['e2e'] and round(d['e2e']['value']/1e6,2))
except Exception as e: print('$tag parse failed', e)"
}
MODE=${2:-quick}
timeout 120 python tools/pcie_probe.py > gpurun_out/pcie_1.log 2>&1; tail -1 gpurun_out/pcie_1.log
timeout 120 $TR --master-port 29513 tools/pcie_probe.py > gpurun_out/pcie_$N.log 2>&1; tail -1 gpurun_out/pcie_$N.log
timeout 120 $TR --master-port 29514 tools/pcie_probe.py --bind > gpurun_out/pcie_bind_$N.log 2>&1; tail -1 gpurun_out/pcie_bind_$N.log
run weak_p2p --no-e2e
run strong_p2p --no-e2e --scaling strong
run strong_p2p_graph --no-e2e --scaling strong --graph
run weak_e2e --no-check
if [ "$MODE" = full ]; then
run weak_nccl --no-e2e --collective nccl
run weak_p2p_instream --no-e2e --no-overlap
fi
python bench.py --steps 30 --warmup 5 --no-cpu --no-ref-cuda --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('single', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4))"
