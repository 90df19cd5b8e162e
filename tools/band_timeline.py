#!/usr/bin/env python
"""Per-launch timeline (all streams) of a few steady-state band frames at 3840x2160, one file per rank.
usage: [torchrun ...] python tools/band_timeline.py [frames] [stripe_rows] -> gpurun_out/band_timeline_rank<r>.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from __graft_entry__ import load_package
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sr = int(sys.argv[2]) if len(sys.argv) > 2 else 64
import torch
dist = None
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
svo = load_package()
path, _ = bench.scene_path()
if rank == 0:
    bench.make_scene(path)
if dist is not None:
    dist.barrier()
octree, root, _ = svo.scene.octree_init(path)
rc, ocl = svo.raycast, svo.ocl
rc.S.mode = "fused"
RX, RY = 3840, 2160
db = svo.bands.DistributedBand(octree, root, RX, RY, lr, stripe_rows=sr, dist=dist)
P = []
for f in range(24 + nfr):
    rc.set_camera(*bench.flythrough_pose(f))
    P.append(rc.prepare_params(RX, RY, f))
for p in P[:24]:
    db.band.frame(p)
db.band.sync()
if dist is not None:
    dist.barrier()
ocl.profile_enable(True); ocl.profile_reset()
for p in P[24:]:
    db.band.frame(p)
db.band.sync()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"band_timeline_n{world}_rank{rank}.txt"), "w") as fh:
    for name, a, b in sorted(ocl.profile_timeline(), key=lambda t: t[1]):
        fh.write(f"{a:9.2f} {b:9.2f} {b - a:8.2f}  {name}\n")
ocl.profile_enable(False)
db.close()
if dist is not None:
    dist.destroy_process_group()
