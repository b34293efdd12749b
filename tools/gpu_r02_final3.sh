#!/bin/bash
# ncu evidence of the FINAL build (launch list + full captures for n_fft 2048 and 512)
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-ref-cuda"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_launch.log 2>&1; echo "launch-list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02h -f $BENCH > gpurun_out/r02_ncu_full.log 2>&1; echo "full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sot_frame -s 3 -c 1 \
    -o gpurun_out/prof_r02h_sot512 -f $BENCH --workload sot512-cut --frames 262144 > gpurun_out/r02_ncu_full_sot512.log 2>&1; echo "full512 exit $?"
