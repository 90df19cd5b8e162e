#!/bin/bash
# same-box A/B of the e2e leg: R,G,B bytes stored by the producing kernels (default) vs word image + pack pass (--e2e-pack)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${PYTEST_K:-present or back_to_back or golden}" > gpurun_out/pytest_e2e.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_e2e.log
for rep in 1 2; do
for v in "" "--e2e-pack"; do
  tag=$(echo "${v:-producers}" | tr -d ' -')
  timeout 300 python bench.py --steps 252 --warmup 4 --no-cpu-baseline --no-extras $v 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/e2e_${tag}_$rep.json
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); e=d['e2e']; print(sys.argv[2] or 'producers', 'fps',round(d['value'],1), 'e2e',round(e['value'],1), e['passes_fps'], 'checksum', e['frame_checksum'], 'link', round(e['d2h_link_gbs'],1), 'launches', d['gpu_launches'])" gpurun_out/e2e_${tag}_$rep.json "$v" || tail -5 gpurun_out/bench_err.log
done; done
