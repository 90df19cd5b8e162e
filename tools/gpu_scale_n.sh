#!/bin/bash
# one point of the scaling table on an N-GPU box: band parity tests across real GPUs (N=2 and 8), bench.py at N ranks
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = 2 ] || [ "$N" = 8 ]; then timeout 600 python -m pytest tests/test_gpu_bands.py -x -q -k multi 2>&1 | tail -2; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "n=$N rc=$?"; tail -2 gpurun_out/scale_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
b = d.get("bands_3840x2160", {})
print("n=$N fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "mrays", round(d["full_raycast_mrays_per_s"], 1),
      "| bands mrays", round(b.get("full_raycast_mrays_per_s", 0), 1), "ms", round(b.get("full_raycast_ms", 0), 4), "warped fps", round(b.get("warped_fps", 0), 1),
      "| 64cam Grays/s", round(d.get("view_parallel_64_cameras", {}).get("grays_per_s", 0), 2))
PY
if [ -n "$WITH_N1" ]; then
  timeout 600 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_n1_on$N.json 2> gpurun_out/scale_n1_on$N.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n1_on$N.json").read().strip().splitlines()[-1])
b = d.get("bands_3840x2160", {})
print("n=1 (same box) fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "mrays", round(d["full_raycast_mrays_per_s"], 1),
      "| bands mrays", round(b.get("full_raycast_mrays_per_s", 0), 1), "ms", round(b.get("full_raycast_ms", 0), 4), "warped fps", round(b.get("warped_fps", 0), 1),
      "| 64cam Grays/s", round(d.get("view_parallel_64_cameras", {}).get("grays_per_s", 0), 2))
PY
fi
