#!/bin/bash
# Bench + profiles on one B200: smoke, bench line, ncu launch list, ncu --set full of the two hot kernels.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-200} --warmup 20 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"
cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
if [ -z "$NO_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/ncu_launch_$R.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'systematic_kernel|step_kernel|tile_sum' -s 60 -c 6 \
    -o gpurun_out/prof_$R -f python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/ncu_full_$R.log 2>&1; echo "ncu full rc=$?"
fi
ls -la gpurun_out
