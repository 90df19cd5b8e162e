#!/usr/bin/env python
"""Per-launch timeline (both streams) of a few steady-state frames of the bench workload, from CUDA events.
usage (GPU box): python tools/timeline.py [frames] [mode] > gpurun_out/timeline.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from __graft_entry__ import load_package
svo = load_package()
nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "fused"
path, _ = bench.scene_path()
bench.make_scene(path)
octree, root, _ = svo.scene.octree_init(path)
rc, ocl = svo.raycast, svo.ocl
rc.raycast_init(octree, root, max_w=1920, max_h=1024, mode=mode)
P = []
for f in range(40 + nfr):
    rc.set_camera(*bench.flythrough_pose(f))
    P.append(rc.prepare_params(1920, 1024, f))
for p in P[:40]:
    rc.draw_prepared(p, sync=False)
ocl.ocl_end_all_kernels()
ocl.profile_enable(True); ocl.profile_reset()
for p in P[40:]:
    rc.draw_prepared(p, sync=False)
ocl.ocl_end_all_kernels()
for name, a, b in sorted(ocl.profile_timeline(), key=lambda t: t[1]):
    print(f"{a:9.2f} {b:9.2f} {b - a:8.2f}  {name}")
rc.raycast_exit()
