#!/bin/bash
# the last pass of a round on a small GPU budget: all GPU tests, the bench line, ncu launch list + --set full of the frame kernels
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench_fused.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"
B="python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras --no-parity --profile-frames 1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_rays|k_proj_scatter2|k_resolve_gather|k_fill_list|k_hole_ids|k_list_scatter' -s 30 -c 14 \
    -o gpurun_out/${TAG}_prof_frame -f $B > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
