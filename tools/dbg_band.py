import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from __graft_entry__ import load_package
import scenes
svo = load_package(); svo.Device.errors_return(); svo.ocl_init(0)
ocl = svo.ocl
octree, root, _ = svo.scene.build_octree(*scenes.small_world())
rx, ry = 320, 192
sr = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def pose(f): return (10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0)
bs = svo.bands.LocalBandSet([0, 0], octree, root, rx, ry, stripe_rows=sr)
rc = svo.raycast; rc.S.mode = "fused"
for ctx in bs.ctxs:
    ocl._svo_ctx_set_current(ctx); ocl.profile_enable(True); ocl.profile_reset()
rc.set_camera(*pose(0)); p = rc.prepare_params(rx, ry, 0)
t = time.time()
try:
    bs.frame(p)
    print("frame 0 ok", time.time() - t)
except RuntimeError as e:
    print("frame 0 FAILED", e, time.time() - t)
for i, ctx in enumerate(bs.ctxs):
    ocl._svo_ctx_set_current(ctx)
    print("band", i)
    for name, a, b in sorted(ocl.profile_timeline(), key=lambda t: t[1]):
        print(f"   {a:12.1f} {b:12.1f} {b - a:12.1f}  {name}")
