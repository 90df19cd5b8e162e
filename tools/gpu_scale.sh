#!/bin/bash
# scaling pass on an N-GPU box: bench.py at 1, 2, .. N ranks (what the driver's SCALE run does), band tests across real GPUs
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bands.py -x -q -k multi 2>&1 | tail -3
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "n=$n rc=$?"; tail -2 gpurun_out/scale_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    b = d.get("bands_3840x2160", {})
    print("n=$n fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "mrays", round(d["full_raycast_mrays_per_s"], 1),
          "| bands mrays", round(b.get("full_raycast_mrays_per_s", 0), 1), "warped fps", round(b.get("warped_fps", 0), 1),
          "| 64cam Grays/s", round(d.get("view_parallel_64_cameras", {}).get("grays_per_s", 0), 2))
except Exception as e:
    print("n=$n parse error", e)
PY
done
