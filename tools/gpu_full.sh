#!/bin/bash
# Full single-GPU pass: all GPU parity tests, then the default bench line (extras included).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_full.json
