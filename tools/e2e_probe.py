#!/usr/bin/env python
"""Where does the end-to-end frame time go?  Same 252 frames as bench.py: device loop, + alternating colorize targets,
+ present without waiting, + the full pipelined read-back (2..4 frames in flight, 32-bit and RGB24)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from __graft_entry__ import load_package
svo = load_package()
path, _ = bench.scene_path()
bench.make_scene(path)
octree, root, _ = svo.scene.octree_init(path)
rc, ocl = svo.raycast, svo.ocl
RX, RY = 1920, 1024
n = RX * RY
rc.raycast_init(octree, root, max_w=RX, max_h=RY, mode="fused")
P = []
for f in range(256):
    rc.set_camera(*bench.flythrough_pose(f))
    P.append(rc.prepare_params(RX, RY, f))

def warm():
    rc.reset_frames()
    for f in range(4):
        rc.draw_prepared(P[f], sync=True)
    ocl.ocl_end_all_kernels()

def report(name, t0):
    ocl.ocl_end_all_kernels()
    dt = time.perf_counter() - t0
    print(f"{name:48s} {252 / dt:8.1f} fps  {dt / 252 * 1e6:7.1f} us/frame", flush=True)

warm(); t0 = time.perf_counter()
for f in range(4, 256):
    rc.draw_prepared(P[f], sync=False)
report("device loop, one colorize target", t0)

for depth in (3,):
    for rgb in (False, True):
        for wait in (True,):
            host = [ocl.host_alloc(n * 4) for _ in range(depth)]
            warm(); t0 = time.perf_counter()
            for f in range(4, 256):
                k = rc.draw_present(P[f], host, rgb24=rgb)
                if wait and f - (depth - 1) >= 4:
                    ocl.present_wait((f - (depth - 1)) % depth)
            if wait:
                for f in range(256 - (depth - 1), 256):
                    ocl.present_wait(f % depth)
            report(f"present depth={depth} rgb24={rgb} wait={wait}", t0)
            for h in host:
                ocl.host_free(h)
rc.raycast_exit()
