#!/bin/bash
# ncu launch list + one --set full capture of the move kernel; R=tag
mkdir -p gpurun_out
R=${ROUND:-r02}
K=${KERNEL:-move_kernel}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/ncu_launch_$R.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 20 -c 1 \
    -o gpurun_out/prof_$R -f python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/ncu_full_$R.log 2>&1; echo "ncu full rc=$?"
