#!/bin/bash
# same-box A/B of the full-screen traversal + the frame for library variants (tools/variants.sh): gpu_fullray.sh <tag> ...
mkdir -p gpurun_out
for rep in 1 2; do
for tag in default "$@"; do
  if [ "$tag" = default ]; then unset SVO_B200_LIB; else export SVO_B200_LIB=$PWD/build/variants/libsvo_b200_$tag.so; fi
  python bench.py --steps 128 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/fr_${tag}_$rep.json
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'fps',round(d['value'],1), 'mrays',round(d['full_raycast_mrays_per_s'],1), {k:v for k,v in d['kernel_ms_per_frame'].items() if 'rays' in k})" gpurun_out/fr_${tag}_$rep.json $tag
done; done
