#!/bin/bash
# tests (all failures), smoke, ncu full capture of the fused kernel
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log | cut -c1-400
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02a -f $BENCH > gpurun_out/ncu_full_r02a.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -6
