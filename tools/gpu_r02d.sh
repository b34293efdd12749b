#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python examples/train_step.py --steps 30 --batch 64 --loss-share > gpurun_out/train_step_1.log 2>&1; tail -1 gpurun_out/train_step_1.log
bash tools/gpu_workloads.sh r02d
