"""GPU tests of the C/C++ host side above the C ABI (include/svo_raycast.h, host/raycast_host.cpp, svo_headless): the
headless mirror of the reference's raycast_init / raycast_draw / raycast_exit driven from C types only, in all three
modes, against the CPU oracle; and the svo_headless program (the reference's main loop without its window)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import scenes
from oracle import frame as ofr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-voxel-octree-raycasting_b200")
MODES = {"reference": 0, "fused": 1, "pingpong": 2}


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(os.path.join(PKG, "libsvo_b200.so"))
    L.svo_octree_build.restype = C.c_void_p
    L.svo_octree_build.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.svo_octree_free.argtypes = [C.c_void_p]
    L.svo_raycast_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.svo_raycast_set_camera.argtypes = [C.c_void_p, C.c_void_p]
    L.svo_raycast_draw.argtypes = [C.c_int, C.c_int, C.c_int]
    L.svo_raycast_read_frame.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.svo_raycast_mem.restype = C.c_void_p
    L.svo_raycast_mem.argtypes = [C.c_char_p]
    L.svo_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.svo_raycast_write_ppm.argtypes = [C.c_char_p, C.c_int, C.c_int]
    return L


def read_mem(lib, name, dtype, count):
    out = np.empty(count, dtype=dtype)
    lib.svo_copy_to_host(out.ctypes.data, lib.svo_raycast_mem(name), out.nbytes, 0)
    return out


@pytest.mark.parametrize("mode", ["reference", "fused", "pingpong"])
def test_c_driver_matches_oracle(lib, orc, mode, tmp_path):
    x, y, z, c = (np.ascontiguousarray(a, dtype=np.uint32) for a in scenes.small_world())
    oct_ref, root = orc.build_octree(x, y, z, c)
    t = lib.svo_octree_build(len(x), x.ctypes.data, y.ctypes.data, z.ctypes.data, c.ctypes.data, 11)
    assert t
    rx, ry = 320, 192
    n = rx * ry
    assert lib.svo_raycast_init(t, rx, ry, 0, MODES[mode]) == 0
    lib.svo_octree_free(t)
    O = ofr.OracleFrame(orc, oct_ref, root, rx, ry, threads=os.cpu_count() or 4)
    try:
        for f in range(6):
            pos = np.array([10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f], dtype=np.float32)
            rot = np.array([0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0], dtype=np.float32)
            O.draw(tuple(pos), tuple(rot))
            lib.svo_raycast_set_camera(pos.ctypes.data, rot.ctypes.data)
            lib.svo_raycast_draw(rx, ry, 1)
            assert lib.svo_raycast_frame() == f
            assert lib.svo_raycast_idbuf_size() == O.idbuf_size
            tex = np.empty(n, dtype=np.uint32)
            lib.svo_raycast_read_frame(tex.ctypes.data, rx, ry)
            assert np.array_equal(tex, O.tex), f"frame {f} image"
            if mode != "pingpong":
                assert np.array_equal(read_mem(lib, b"screenbuffer", np.uint32, 4 * n), O.screen[:4 * n]), f"frame {f} colour"
                assert np.array_equal(read_mem(lib, b"backbuffer", np.uint32, 16 * n), O.back[:16 * n].view(np.uint32)), f"frame {f} xyz"
            ids = read_mem(lib, b"idbuffer", np.uint32, 2 * O.nblocks + O.idbuf_size)
            assert np.array_equal(ids, O.idbuf[:2 * O.nblocks + O.idbuf_size]), f"frame {f} ids"
        # headless writer: P6 PPM of the colorized frame (0x00RRGGBB -> R, G, B bytes)
        path = str(tmp_path / "frame.ppm")
        assert lib.svo_raycast_write_ppm(path.encode(), rx, ry) == 0
        raw = open(path, "rb").read()
        header = f"P6\n{rx} {ry}\n255\n".encode()
        assert raw.startswith(header) and len(raw) == len(header) + 3 * n
        rgb = np.frombuffer(raw[len(header):], dtype=np.uint8).reshape(ry, rx, 3)[::-1].reshape(n, 3)   # texture row 0 is the bottom row
        assert np.array_equal(rgb[:, 0], (O.tex >> 16) & 255) and np.array_equal(rgb[:, 1], (O.tex >> 8) & 255) and np.array_equal(rgb[:, 2], O.tex & 255)
    finally:
        lib.svo_raycast_exit()


def test_headless_program_runs(tmp_path):
    """svo_headless on a small .rle4 scene: loader -> builder -> 5 frames -> PPM, and the image equals the Python driver's."""
    from __graft_entry__ import load_package
    svo = load_package()
    vox = svo.scene.generate(kind=2, depth=11, size=256, nblobs=0, seed=3)
    path = str(tmp_path / "terrain.rle4")
    vox.write_rle4(path, 256, 2048, 256)
    vox.free()
    out = str(tmp_path / "f")
    exe = os.path.join(PKG, "svo_headless")
    r = subprocess.run([exe, "--rle4", path, "--res", "320x192", "--frames", "5", "--every", "4", "--out", out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "5 frames 320x192" in r.stdout
    raw = open(out + "_0004.ppm", "rb").read()
    rgb = np.frombuffer(raw[len(b"P6\n320 192\n255\n"):], dtype=np.uint8).reshape(192, 320, 3)[::-1].reshape(-1, 3).astype(np.uint32)
    # the same 5 frames through the Python mirror
    import math
    octree, root, _ = svo.scene.octree_init(path)
    rc = svo.raycast
    rc.raycast_init(octree, root, max_w=320, max_h=192, mode="fused")
    try:
        for f in range(5):
            # the program computes its pose in single precision (headless_main.cpp)
            f32 = np.float32
            p = f32(1.0) + f32(f) * f32(0.2357)
            rx_ = f32(0.6) + f32(0.1) * f32(math.sin(2.0 * 3.14159265358979 * f / 128.0))
            rc.set_camera((p, f32(50.0), p), (rx_, f32(0.8) + f32(0.005) * f32(f), f32(0.0)))
            rc.raycast_draw(320, 192)
        tex = rc.read_frame(320, 192).ravel()
    finally:
        rc.raycast_exit()
    assert np.array_equal((rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2], tex)
