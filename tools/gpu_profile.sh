#!/bin/bash
# ncu evidence for the bench command: launch list (durations) + one full capture of each hot kernel.
# usage: tools/gpu_profile.sh <tag> [extra bench args]
TAG=${1:-r01}; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu $*"
# every launch of the timed steps (bench.py brackets them with cudaProfilerStart/Stop)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch-list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 6 -c 2 \
    -o gpurun_out/prof_${TAG} -f $BENCH > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -8
