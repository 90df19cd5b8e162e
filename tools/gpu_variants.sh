#!/bin/bash
# same-box A/B of library variants built by tools/variants.sh: gpu_variants.sh <tag> [<tag> ...]  ("default" = the in-tree build)
mkdir -p gpurun_out
for rep in 1 2; do
for tag in default "$@"; do
  if [ "$tag" = default ]; then unset SVO_B200_LIB; else export SVO_B200_LIB=$PWD/build/variants/libsvo_b200_$tag.so; fi
  python bench.py --steps 252 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/var_${tag}_$rep.json
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'fps',round(d['value'],1), 'e2e',round(d['e2e']['value'],1), 'mrays',round(d['full_raycast_mrays_per_s'],1), {k:v for k,v in d['kernel_ms_per_frame'].items() if 'rays' in k})" gpurun_out/var_${tag}_$rep.json $tag
done; done
