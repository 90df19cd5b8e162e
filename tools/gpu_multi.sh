#!/bin/bash
# multi-GPU pass: bench at N ranks under torchrun (N = $1), a few stripe heights
N=${1:-2}
mkdir -p gpurun_out
for SR in 64 16 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --stripe-rows $SR > gpurun_out/bench_n${N}_sr${SR}.json 2> gpurun_out/bench_n${N}_sr${SR}.err
  echo "N=$N SR=$SR rc=$?"; tail -c 1500 gpurun_out/bench_n${N}_sr${SR}.json; tail -3 gpurun_out/bench_n${N}_sr${SR}.err
done
