#!/bin/bash
# multi-GPU band pass: parity tests across real GPUs, then the 3840x2160 band measurements at a few stripe heights
N=${1:-2}; shift
SRS=${@:-"64 32 16"}
mkdir -p gpurun_out
[ -z "$SKIP_TESTS" ] && timeout 600 python -m pytest tests/test_gpu_bands.py -x -q 2>&1 | tail -3
for SR in $SRS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --bands-only --stripe-rows $SR > gpurun_out/bands_n${N}_sr${SR}.json 2> gpurun_out/bands_n${N}_sr${SR}.err
  echo "N=$N SR=$SR rc=$?"; tail -c 600 gpurun_out/bands_n${N}_sr${SR}.json; echo; tail -2 gpurun_out/bands_n${N}_sr${SR}.err | cut -c1-200
done
python bench.py --gpus 1 --bands-only 2>/dev/null | tail -c 600
