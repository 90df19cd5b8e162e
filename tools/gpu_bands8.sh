#!/bin/bash
N=${1:-8}; shift
mkdir -p gpurun_out
run() {  # env, ray stripe rows
  env $1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --bands-only --ray-stripe-rows $2 2>/dev/null | tail -1 | python -c "
import json,sys; b=json.loads(sys.stdin.read())['bands_3840x2160']; print('$1 SR=$2', 'mrays', round(b['full_raycast_mrays_per_s'],1), 'ms', round(b['full_raycast_ms'],4), 'warped', round(b['warped_fps'],1))"
}
for sr in ${@:-16}; do run "A=1" $sr; done
