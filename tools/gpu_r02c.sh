#!/bin/bash
# A/B of kernel variants + tests
mkdir -p gpurun_out
bash tools/bench_variants.sh --no-ref-cuda > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
timeout 1800 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -rf -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-400
