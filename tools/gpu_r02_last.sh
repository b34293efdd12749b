#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "support_positions" 2>&1 | tail -25 | cut -c1-220
timeout 100 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -rf > gpurun_out/pytest_gpu_last.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_last.log | cut -c1-220
