#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, in-pipeline DRAM traffic, ncu --set full of the frame kernels.
# Outputs in gpurun_out/; summarise with tools/ncu_summary.py and copy into profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_fused.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_fused.json
timeout 300 python bench.py --mode pingpong --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_pingpong.json 2>> gpurun_out/bench_err.log
B="python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras --profile-frames 1"
# every launch with its device time (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
# DRAM bytes per launch with the caches left as the pipeline leaves them (one pass per kernel, no flush)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --cache-control none --clock-control none -s 40 -c 160 --csv \
    --log-file gpurun_out/${TAG}_dram_inpipeline.csv $B > gpurun_out/ncu_dram.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rays|k_proj_scatter2|k_resolve_gather|k_copy_colorize|k_fill_list|k_hole_ids|k_list_scatter' -s 30 -c 14 \
    -o gpurun_out/${TAG}_prof_frame -f $B > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_raycast_fine_2' -c 2 \
    -o gpurun_out/${TAG}_prof_fullray -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --profile-frames 1 > gpurun_out/ncu_fullray.log 2>&1
ls -la gpurun_out | tail -15
