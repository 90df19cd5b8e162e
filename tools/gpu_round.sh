#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, ncu --set full of the frame kernels.  Outputs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --mode pingpong --no-cpu-baseline > gpurun_out/bench_pingpong.json 2>> gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 16 --warmup 4 --no-cpu-baseline --profile-frames 1 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rays|k_proj_scatter2|k_resolve_gather|k_copy_colorize|k_fill_list' -s 30 -c 12 \
    -o gpurun_out/prof_frame -f python bench.py --steps 8 --warmup 4 --no-cpu-baseline --profile-frames 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_raycast_fine_2' -c 2 \
    -o gpurun_out/prof_fullray -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile-frames 1 > gpurun_out/ncu_fullray.log 2>&1
ls -la gpurun_out
