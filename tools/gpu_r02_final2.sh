#!/bin/bash
# Final validation of round 2 on 2 GPUs: whole GPU test suite (incl. the 2-GPU exchange test), smoke, default bench.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02_bench_default.log 2>&1; echo "default exit $?"; tail -1 gpurun_out/r02_bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print(round(d['value']/1e6,2), 'Mframes/s', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'step frac', round(r['step']['frac'],3), 'e2e', round(d['e2e']['value']/1e6,2), 'cpu', round(d['cpu_baseline']['value']), 'refcuda', round(d['ref_on_cuda']['value']/1e6,3), d['clocks'])"
bash tools/gpu_workloads.sh r02g
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-ref-cuda > gpurun_out/r02_bench_weak_full_2.log 2>&1; tail -1 gpurun_out/r02_bench_weak_full_2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('2 gpus', round(d['value']/1e6,2), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,2), d['collective'], d['value_check'])"
