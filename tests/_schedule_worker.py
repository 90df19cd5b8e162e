"""Worker of test_schedule_switches: renders a short back-to-back fused sequence with the schedule switches named in
SVO_TEST_SWITCHES ("name=value,name=value", handed to svo_debug_set before the context exists) and prints a digest of every
buffer, of the id list and of the last two images (one process per combination: "main_lo" acts at context creation)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
import scenes  # noqa: E402

svo = load_package()
svo.Device.errors_return()
for item in filter(None, os.environ.get("SVO_TEST_SWITCHES", "").split(",")):
    name, _, value = item.partition("=")
    svo.ocl.debug_set(name, int(value or 1))
mode, rx, ry, nframes = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
octree, root, _ = svo.scene.build_octree(*scenes.small_world())
rc, ocl = svo.raycast, svo.ocl
rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode=mode)
n = rx * ry
host = [ocl.host_alloc(n * 4), ocl.host_alloc(n * 4)]
h = hashlib.sha1()
for f in range(nframes):
    rc.set_camera((10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0))
    if f >= 2:
        ocl.present_wait(f & 1)
        h.update(np.frombuffer(host[f & 1], dtype=np.uint32).tobytes())      # every frame's image
    rc.draw_present(rc.prepare_params(rx, ry, f), host)
for f in range(nframes - 2, nframes):
    ocl.present_wait(f & 1)
    h.update(np.frombuffer(host[f & 1], dtype=np.uint32).tobytes())
screen, back, idb = rc.read_buffers(rx, ry)
nb = (rx // 16) * (ry // 16)
size = rc.idbuf_size()
slot = rc.last_slot()
if mode == "pingpong":                                    # only the slot rendered into and the ids are defined
    h.update(screen[slot * n:(slot + 1) * n].tobytes())
else:
    h.update(screen.tobytes()); h.update(back.tobytes())
h.update(idb[:2 * nb + size].tobytes())
print("DIGEST", h.hexdigest(), ocl.frame_deferred_count())
rc.raycast_exit()
