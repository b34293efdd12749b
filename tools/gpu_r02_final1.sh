#!/bin/bash
# Final single-GPU evidence of round 2: default bench line, reference arm, other workloads, paper-batch latency,
# ncu launch list and full captures.  Everything lands in gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_default.log 2>&1; echo "default exit $?"; tail -1 gpurun_out/r02_bench_default.log | cut -c1-400
timeout 400 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference.log 2>&1; echo "reference exit $?"; tail -1 gpurun_out/r02_bench_reference.log | cut -c1-300
bash tools/gpu_workloads.sh r02f
timeout 300 python bench.py --workload sot2048-cut --frames 1024 --steps 200 --warmup 20 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r02_latency_eager.json
timeout 300 python bench.py --workload sot2048-cut --frames 1024 --steps 200 --warmup 20 --no-e2e --no-cpu --no-ref-cuda --graph 2>/dev/null | tail -1 > gpurun_out/r02_latency_graph.json
python - <<PY
import json
for f in ("r02_latency_eager", "r02_latency_graph"):
    d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["ms_per_step"]*1e3,1), "us/step", d["gpu_launches"], "launches", d.get("step_modes"), d.get("ref_on_cuda") and round(d["ref_on_cuda"]["ms_per_step"]*1e3,1))
PY
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_launch.log 2>&1; echo "launch-list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02f -f $BENCH > gpurun_out/r02_ncu_full.log 2>&1; echo "full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02f_sot512 -f $BENCH --workload sot512-cut --frames 262144 > gpurun_out/r02_ncu_full_sot512.log 2>&1; echo "full512 exit $?"
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
