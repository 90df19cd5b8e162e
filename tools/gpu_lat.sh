#!/bin/bash
# GPU pass of the latency tools: instruction latencies + cycles per trip of the traversal variants given as arguments (tags of tools/ubench/build.sh)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
build/ubench_lat | tee gpurun_out/ubench_lat.txt
for tag in "" "$@"; do
  t=${tag:+_$tag}
  echo "== raylat$t"
  for f in 3 8 12; do build/raylat$t $f; done | tee gpurun_out/raylat$t.txt
done
