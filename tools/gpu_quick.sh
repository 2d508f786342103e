#!/bin/bash
# tests + a short bench in one box lease; R=tag
mkdir -p gpurun_out
R=${ROUND:-r02}
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cat gpurun_out/bench_$R.json | cut -c1-1800; tail -3 gpurun_out/bench_$R.err
