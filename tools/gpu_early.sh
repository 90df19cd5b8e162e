#!/bin/bash
# GPU pass for the early reprojection (svo_frame_early_count): sequence parity tests, then same-box A/B of the bench with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${PYTEST_K:-back_to_back or schedule_switches or golden or cache_rotation or present}" > gpurun_out/pytest_early.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_early.log
tail -15 gpurun_out/pytest_early.log
for rep in 1 2; do
for sw in "" $EXTRA_SW; do
  tag=$(echo "${sw:-default}" | tr ' =,' '___')
  timeout 300 python bench.py --steps 252 --warmup 4 --no-cpu-baseline --no-extras ${sw:+--debug-switches $sw} 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/early_${tag}_$rep.json
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2] or 'default', 'fps',round(d['value'],1), 'e2e',round(d['e2e']['value'],1), 'parity', d.get('parity',{}).get('mismatching_words'), 'launches', d['gpu_launches']); print('  ', {k:round(v*1000,1) for k,v in d['kernel_ms_per_frame'].items()})" gpurun_out/early_${tag}_$rep.json "$sw" || tail -5 gpurun_out/bench_err.log
done; done
python tools/timeline.py 4 fused > gpurun_out/timeline_early.txt 2>&1; tail -45 gpurun_out/timeline_early.txt
