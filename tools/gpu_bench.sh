#!/bin/bash
# One B200: smoke, GPU parity tests, the default bench line, the ncu launch list and one `--set full` capture of the hot kernels.
# Outputs under gpurun_out/ (scratch); tools/summarize_profile.py + tools/sass_ops.py turn them into the files kept in profiles/.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cat gpurun_out/bench_$R.json
timeout 300 python bench.py --workload smc2 --steps 250 --warmup 10 > gpurun_out/bench_smc2_$R.json 2> gpurun_out/bench_smc2_$R.err; echo "smc2 rc=$?"; cat gpurun_out/bench_smc2_$R.json
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/ncu_launch_$R.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'step_kernel|resample_fused|expand_kernel|normalize_kernel' -s 20 -c 2 \
    -o gpurun_out/prof_$R -f python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/ncu_full_$R.log 2>&1; echo "ncu full rc=$?"
fi
