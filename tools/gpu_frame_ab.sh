#!/bin/bash
# same-box A/B of the warped frame: gpu_frame_ab.sh "<tag> [ENV=val ...]" ...   (tag = library variant of tools/variants.sh or "default")
mkdir -p gpurun_out
for rep in 1 2; do
for spec in "$@"; do
  set -- $spec; tag=$1; shift
  if [ "$tag" = default ]; then unset SVO_B200_LIB; else export SVO_B200_LIB=$PWD/build/variants/libsvo_b200_$tag.so; fi
  name=$(echo "$spec" | tr ' =' '__')
  env "$@" python bench.py --steps 252 --warmup 4 --no-cpu-baseline --no-extras --no-parity 2>/dev/null | tail -1 > gpurun_out/ab_${name}_$rep.json
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'fps',round(d['value'],1), 'e2e',round(d['e2e']['value'],1), {k:v for k,v in d['kernel_ms_per_frame'].items() if 'rays' in k or 'gather' in k})" gpurun_out/ab_${name}_$rep.json "$spec"
done; done
