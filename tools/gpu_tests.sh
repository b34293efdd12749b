#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log | cut -c1-250
