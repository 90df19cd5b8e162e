"""Host-side scene contract (include/svo_host.h) bound for Python: .rle4 loader/writer, the direct
voxel -> compact-octree builder (bit-identical to set_voxel + convert_tree_blocks, src/octree/octree.h) and the
procedural stand-in scenes.  Mirrors ``octree_init()`` (src/raycast.h:13-46): ``octree_init(path)`` returns
(octree_array_compact, octree_root_normal)."""
import ctypes as C
import os

import numpy as np

from .ocl import lib

_vp, _sz, _u32, _i = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def _sig(name, res, *args):
    f = getattr(lib, name)
    f.restype = res
    f.argtypes = list(args)
    return f


_build = _sig("svo_octree_build", _vp, _sz, _u32p, _u32p, _u32p, _u32p, _i)
_build_voxels = _sig("svo_octree_build_voxels", _vp, _vp, _i)
_words = _sig("svo_octree_words", C.POINTER(C.c_uint32), _vp)
_num_words = _sig("svo_octree_num_words", _sz, _vp)
_root = _sig("svo_octree_root", _u32, _vp)
_num_voxels = _sig("svo_octree_num_voxels", C.c_uint64, _vp)
_num_unique = _sig("svo_octree_num_unique_voxels", C.c_uint64, _vp)
_octree_free = _sig("svo_octree_free", None, _vp)
_vox_count = _sig("svo_voxels_count", _sz, _vp)
_vox_x = _sig("svo_voxels_x", C.POINTER(C.c_uint32), _vp)
_vox_y = _sig("svo_voxels_y", C.POINTER(C.c_uint32), _vp)
_vox_z = _sig("svo_voxels_z", C.POINTER(C.c_uint32), _vp)
_vox_rgba = _sig("svo_voxels_rgba", C.POINTER(C.c_uint32), _vp)
_vox_free = _sig("svo_voxels_free", None, _vp)
_rle4_load = _sig("svo_rle4_load", _vp, C.c_char_p, _i, _i, _i, _i)
_rle4_write = _sig("svo_rle4_write", _i, C.c_char_p, _vp, _i, _i, _i)
_scene_generate = _sig("svo_scene_generate", _vp, _i, _i, _i, _i, _u32)


class Voxels:
    """A voxel stream in insertion order (host memory, owned by the library)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("svo_b200: could not create the voxel stream")
        self.handle = handle

    def __len__(self):
        return int(_vox_count(self.handle))

    def arrays(self):
        n = len(self)
        return tuple(np.ctypeslib.as_array(f(self.handle), shape=(n,)).copy() for f in (_vox_x, _vox_y, _vox_z, _vox_rgba))

    def write_rle4(self, path, sx, sy, sz):
        rc = _rle4_write(os.fsencode(path), self.handle, sx, sy, sz)
        if rc:
            raise RuntimeError(f"svo_rle4_write failed ({rc})")

    def free(self):
        if self.handle:
            _vox_free(self.handle)
            self.handle = None


def _finish(t):
    if not t:
        raise RuntimeError("svo_b200: octree build failed")
    n = int(_num_words(t))
    words = np.ctypeslib.as_array(_words(t), shape=(n,)).copy()
    root, nv, nu = int(_root(t)), int(_num_voxels(t)), int(_num_unique(t))
    _octree_free(t)
    return words, root, dict(num_voxels=nv, num_unique=nu)


def build_octree(x, y, z, rgba, depth=11):
    """set_voxel stream -> (octree_array_compact, octree_root_normal, stats)."""
    x, y, z, rgba = (np.ascontiguousarray(a, dtype=np.uint32) for a in (x, y, z, rgba))
    return _finish(_build(len(x), x, y, z, rgba, depth))


def build_octree_voxels(vox, depth=11):
    return _finish(_build_voxels(vox.handle, depth))


_build_device = _sig("svo_octree_build_device", _vp, _sz, _u32p, _u32p, _u32p, _u32p, _i, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64))


def build_octree_device(x, y, z, rgba, depth=11):
    """The same array as build_octree(), built on the GPU of the current context (ocl_init first) and left there:
    -> (Mem usable as mem_octree, octree_root_normal, stats)."""
    from . import ocl
    x, y, z, rgba = (np.ascontiguousarray(a, dtype=np.uint32) for a in (x, y, z, rgba))
    root, nu = C.c_uint32(), C.c_uint64()
    h = _build_device(len(x), x, y, z, rgba, depth, C.byref(root), C.byref(nu))
    ocl._check()
    if not h:
        raise RuntimeError("svo_octree_build_device failed")
    return ocl.Mem(h, int(ocl._svo_mem_size(h))), int(root.value), dict(num_voxels=len(x), num_unique=int(nu.value))


_rle4_load_mip = _sig("svo_rle4_load_mip", _vp, C.c_char_p, _i, _i, _i, _i, _i)
_load_rle4_device = _sig("svo_octree_load_rle4_device", _vp, C.c_char_p, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_uint32),
                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))


def rle4_load(path, palette=0, addx=0, addy=0, addz=0, mip=0):
    h = _rle4_load_mip(os.fsencode(path), mip, palette, addx, addy, addz)
    if not h:
        raise RuntimeError(f"svo_b200: cannot load mip {mip} of {path}")
    return Voxels(h)


def octree_init_device(path, depth=11, mip=0, palette=0, addx=0, addy=0, addz=0):
    """octree_init() with the .rle4 decoded and the octree built on the GPU of the current context (ocl_init first):
    -> (Mem usable as mem_octree, octree_root_normal, stats).  Nothing but the file's slab stream crosses PCIe."""
    from . import ocl
    root, nv, nu = C.c_uint32(), C.c_uint64(), C.c_uint64()
    h = _load_rle4_device(os.fsencode(path), mip, palette, addx, addy, addz, depth, C.byref(root), C.byref(nv), C.byref(nu))
    ocl._check()
    if not h:
        raise RuntimeError("svo_octree_load_rle4_device failed")
    return ocl.Mem(h, int(ocl._svo_mem_size(h))), int(root.value), dict(num_voxels=int(nv.value), num_unique=int(nu.value))


def octree_init(path, depth=11):
    """src/raycast.h:13-46 with the file name as a parameter (the reference hard-codes ../data/Imrodh.rle4)."""
    vox = rle4_load(path, 0, 0, 0, 0)
    try:
        return build_octree_voxels(vox, depth)
    finally:
        vox.free()


def generate(kind=1, depth=11, size=0, nblobs=6, seed=0x5EED):
    """kind 1 = S1 stand-in (floor plate + blobs), kind 2 = S2 fractal terrain."""
    return Voxels(_scene_generate(kind, depth, size, nblobs, seed))
