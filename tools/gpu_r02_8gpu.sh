#!/bin/bash
# One 8-GPU session: exchange checks, host-bandwidth probe, weak and strong scaling lines, DDP training step.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nproc > gpurun_out/host_$N.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> gpurun_out/host_$N.txt; nvidia-smi topo -m >> gpurun_out/host_$N.txt 2>&1
timeout 240 $TR --master-port 29511 tools/p2p_check.py > gpurun_out/p2p_check_$N.log 2>&1; echo "p2p_check exit $?"; tail -1 gpurun_out/p2p_check_$N.log | cut -c1-400
timeout 100 python tools/pcie_probe.py > gpurun_out/pcie_1of$N.log 2>&1; tail -1 gpurun_out/pcie_1of$N.log
timeout 100 $TR --master-port 29513 tools/pcie_probe.py > gpurun_out/pcie_$N.log 2>&1; tail -1 gpurun_out/pcie_$N.log
timeout 100 $TR --master-port 29514 tools/pcie_probe.py --bind > gpurun_out/pcie_bind_$N.log 2>&1; tail -1 gpurun_out/pcie_bind_$N.log
run() {  # tag, args...
  tag=$1; shift
  timeout 240 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 "$@" > gpurun_out/bench_${tag}_$N.log 2>&1
  echo "$tag exit $?"; tail -1 gpurun_out/bench_${tag}_$N.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$tag', d['n_gpus'], 'gpus', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4), 'ms/step', d['collective']['used'], [round(t,4) for t in d.get('ms_per_step_by_rank')], d.get('host_issue'), 'check', d['value_check'] and d['value_check']['rel_err'], 'e2e', d['e2e'] and round(d['e2e']['value']/1e6,2))
except Exception as e: print('$tag parse failed', e)"
}
run weak_full
run strong_graph --no-e2e --no-ref-cuda --scaling strong --graph
run strong_eager --no-e2e --no-ref-cuda --scaling strong
run weak_nccl --no-e2e --no-ref-cuda --collective nccl
timeout 200 $TR --master-port 29515 examples/train_step.py --steps 30 --batch 64 --loss-share > gpurun_out/train_step_$N.log 2>&1; tail -1 gpurun_out/train_step_$N.log
python bench.py --steps 50 --warmup 10 --no-cpu --no-ref-cuda --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_single_on_$N.json; python -c "
import json
d = json.load(open('gpurun_out/bench_single_on_$N.json')); print('single', round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4))"
