#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
bash tools/gpu_workloads.sh r02h
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02i_sot512 -f $BENCH --workload sot512-cut --frames 262144 > gpurun_out/r02_ncu_full_sot512.log 2>&1; echo "full512 exit $?"
