"""The drop-in, proven mechanically: oracle/_ref/ref_raycast_dropin is the reference's OWN frame driver
src/raycast.h -- with its octree_init (RLE4::load -> convert_tree_blocks), its camera code and its 13 positional launches per
frame -- compiled unchanged against include/compat/ocl.h (in the place of src/ocl.h) and linked with libsvo_b200.so
(tests/dropin/ref_raycast_main.cpp; built by `make -C oracle dropin` where /root/reference is mounted).  It is driven like
src/main.cpp drives it (mouse position, W key) for 8 frames; every colour word, position and colorized pixel of every frame
must equal what the reference's kernel.cl (oracle/_ref, or the C restatement) produces for the same camera."""
import os
import subprocess

import numpy as np
import pytest

from oracle import binding, frame as ofr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_raycast_dropin")


def test_reference_raycast_h_runs_on_the_library(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/ref_raycast_dropin not built (needs /root/reference at build time)")
    from __graft_entry__ import load_package
    svo = load_package()
    # a small scene around the reference's start position pos = (1, 50, 1) (src/raycast.h:113), as a .rle4 file
    scene = str(tmp_path / "scene.rle4")
    vox = svo.scene.generate(kind=1, depth=11, size=640, nblobs=3, seed=7)
    vox.write_rle4(scene, 2048, 2048, 2048)
    vox.free()
    rx, ry, frames = 320, 192, 8
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([EXE, scene, str(rx), str(ry), str(frames), str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "through the reference's raycast.h" in r.stdout

    lib = binding.get("ref") if binding.have_ref() else binding.get("orc")
    octree, root = lib.build_octree_rle4(scene)
    O = ofr.OracleFrame(lib, octree, root, rx, ry, threads=os.cpu_count() or 4)
    n = rx * ry
    cams = [l.split() for l in open(out / "camera.txt")]
    assert len(cams) == frames
    seen = 0
    for f, c in enumerate(cams):
        pos = tuple(float(v) for v in c[1:4])
        rot = (float(c[4]), float(c[5]), 0.0001 * 0)           # rot.z stays 0 (raycast.h:114: vec3f rot(0.0001,0,0), x and y overwritten)
        O.draw(pos, rot)
        screen = np.fromfile(out / f"f{f:02d}_screen.bin", dtype=np.uint32)
        back = np.fromfile(out / f"f{f:02d}_back.bin", dtype=np.uint32)
        tex = np.fromfile(out / f"f{f:02d}_tex.bin", dtype=np.uint32)
        assert np.array_equal(screen, O.screen[:4 * n]), f"frame {f}: colour buffers"
        assert np.array_equal(back, O.back[:16 * n].view(np.uint32)), f"frame {f}: coordinate buffers"
        assert np.array_equal(tex, O.tex), f"frame {f}: colorized image"
        seen += int(np.count_nonzero((O.screen[:n] & 0xff) != 0))
    assert seen > frames * n // 4, "the camera does not see the scene"
