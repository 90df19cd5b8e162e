#!/bin/bash
# Where does the replica loss come from?  (a) one independent N=1 bench alone, (b) two independent N=1 benches at once on two GPUs
# (no torch.distributed, no NCCL, no peer mappings), (c) the torchrun N=2 job.  Same box, 252 frames.
mkdir -p gpurun_out
B="python bench.py --steps 252 --warmup 4 --no-cpu-baseline --no-extras --no-parity"
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], 'fps/gpu', round(d['value']/d['n_gpus'],1), 'ms', round(d['ms_per_step'],5), 'by rank', d['details'].get('ms_per_step_by_rank'), 'enq', round(d['host_enqueue_ms_per_frame'],4), 'e2e/gpu', round(d['e2e']['value']/d['n_gpus'],1))" $1 "$2"; }
CUDA_VISIBLE_DEVICES=0 $B 2>/dev/null | tail -1 > gpurun_out/rp_alone.json; show gpurun_out/rp_alone.json "alone(gpu0)"
CUDA_VISIBLE_DEVICES=0 taskset -c 0-7 $B 2>/dev/null | tail -1 > gpurun_out/rp_pair0.json &
CUDA_VISIBLE_DEVICES=1 taskset -c 8-15 $B 2>/dev/null | tail -1 > gpurun_out/rp_pair1.json &
wait
show gpurun_out/rp_pair0.json "independent pair, gpu0"; show gpurun_out/rp_pair1.json "independent pair, gpu1"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 252 --warmup 4 --no-cpu-baseline --no-extras --no-parity 2>/dev/null | tail -1 > gpurun_out/rp_torchrun2.json
show gpurun_out/rp_torchrun2.json "torchrun N=2"
