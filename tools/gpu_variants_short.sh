#!/bin/bash
# one short same-box pass over library variants (tools/variants.sh): ray parity test + full-raycast rate + frame rate each
mkdir -p gpurun_out
for tag in default "$@"; do
  if [ "$tag" = default ]; then unset SVO_B200_LIB; else export SVO_B200_LIB=$PWD/build/variants/libsvo_b200_$tag.so; fi
  timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "raycast_fine_2_full_screen or golden" 2>&1 | tail -1
  timeout 100 python bench.py --steps 60 --warmup 4 --no-cpu-baseline --no-extras --no-parity 2>/dev/null | tail -1 > gpurun_out/var_${tag}.json
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'fps',round(d['value'],1), 'mrays',round(d['full_raycast_mrays_per_s'],1), {k:round(v*1000,1) for k,v in d['kernel_ms_per_frame'].items() if 'rays' in k})" gpurun_out/var_${tag}.json $tag
done
