#!/bin/bash
# tests + bench (+ optional ncu) in one box lease
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
bash tools/gpu_bench.sh
