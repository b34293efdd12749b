#!/bin/bash
# Round 2, first contact: GPU tests, smoke, default bench, paper-batch latency line, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02a.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_r02a.log; tail -2 gpurun_out/bench_r02a.log | cut -c1-2500
timeout 300 python bench.py --workload sot2048-cut --frames 1024 --steps 200 --warmup 20 --no-e2e --no-cpu > gpurun_out/bench_r02a_latency.log 2>&1; tail -1 gpurun_out/bench_r02a_latency.log | cut -c1-1500
timeout 300 python bench.py --workload sot2048-cut --frames 1024 --steps 200 --warmup 20 --no-e2e --no-cpu --no-ref-cuda --graph > gpurun_out/bench_r02a_latency_graph.log 2>&1; tail -1 gpurun_out/bench_r02a_latency_graph.log | cut -c1-600
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/launches_r02a.csv $BENCH > gpurun_out/ncu_launch_r02a.log 2>&1
echo "launch-list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02a -f $BENCH > gpurun_out/ncu_full_r02a.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -12
