#!/bin/bash
# Quick GPU pass for a scheduling change: the frame-sequence parity tests, then the bench under each variant given as
# arguments (environment assignments, "" = default), then the per-launch timeline of the default.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frame_sequence or golden" > gpurun_out/pytest_seq.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_seq.log
tail -5 gpurun_out/pytest_seq.log
[ $# -eq 0 ] && set -- ""
for v in "$@"; do
  echo "== bench $v"
  tag=$(echo "${v:-default}" | tr ' =' '__')
  env $v python bench.py --steps 252 --warmup 4 --no-cpu-baseline --no-extras 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_ab_$tag.json
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print('fps',round(d['value'],1), 'ms',round(d['ms_per_step'],4), 'e2e',round(d['e2e']['value'],1), 'mrays',round(d['full_raycast_mrays_per_s'],1), 'launches',d['gpu_launches']); print(d['kernel_ms_per_frame'])" gpurun_out/bench_ab_$tag.json
done
python tools/timeline.py 4 fused > gpurun_out/timeline_ab.txt 2>&1; tail -40 gpurun_out/timeline_ab.txt
