"""ctypes binding for the CPU oracles -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (sparse-voxel-octree-raycasting_b200/)
never does.

Two libraries share one C interface (functions differ only in prefix):

* ``oracle/_ref/libsvo_ref.so``  (prefix ``ref_``)  the reference's own
  kernel/kernel.cl + src/octree/octree.h + src/octree/Rle4.cpp compiled through
  oracle/ref_shim (only buildable where /root/reference is mounted);
* ``oracle/libsvo_oracle.so``    (prefix ``orc_``)  the plain-C restatement
  oracle/svo_oracle.c, buildable anywhere gcc exists.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsvo_ref.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libsvo_ref_fast.so")     # -O3 -ffast-math twin (src/ocl.h:47 -cl-fast-relaxed-math): timing only
ORC_SO = os.path.join(HERE, "libsvo_oracle.so")

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
HOLE = 0xFFFFFF00


def build(target="oracle"):
    """make -C oracle <target>; 'ref' needs /root/reference."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def have_ref():
    return os.path.exists(REF_SO)


def have_ref_fast():
    return os.path.exists(REF_FAST_SO)


import contextlib


@contextlib.contextmanager
def quiet_stdout():
    """The reference's loader prints its progress with printf (src/octree/Rle4.cpp:22-89): keep it off this process' stdout
    (bench.py prints one JSON line there)."""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        try:
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def _vec4(v):
    a = np.zeros(4, dtype=np.float32)
    v = np.asarray(v, dtype=np.float32).ravel()
    a[: len(v)] = v
    return a


class CpuOracle:
    """One loaded oracle library (prefix 'ref' or 'orc')."""

    def __init__(self, prefix):
        fast = prefix == "ref_fast"
        prefix = "ref" if fast else prefix
        self.prefix = prefix
        path = REF_FAST_SO if fast else REF_SO if prefix == "ref" else ORC_SO
        if prefix == "orc" and (not os.path.exists(path)
                                or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "svo_oracle.c"))):
            build("oracle")
        self.lib = C.CDLL(path)
        L, p = self.lib, prefix + "_"
        i, u, f, vp = C.c_int, C.c_uint, C.c_float, C.c_void_p

        def fn(name, res, *args):
            h = getattr(L, p + name)
            h.restype = res
            h.argtypes = list(args)
            return h

        self._reset = fn("reset", None)
        self._set_voxels = fn("set_voxels", None, C.c_size_t, u32p, u32p, u32p, u32p)
        self._load_rle4 = fn("load_rle4", None, C.c_char_p, i, i, i, i)
        self._convert = fn("convert", u)
        self._compact_words = fn("compact_words", C.c_size_t)
        self._compact_data = fn("compact_data", C.POINTER(C.c_uint32))
        self._num_voxels = fn("num_voxels", u)
        self._octree_root = fn("octree_root", u)
        self._num_nodes = fn("num_nodes", C.c_size_t)
        self._memset = fn("memset", None, i, u32p, u, u)
        self._memcpy = fn("memcpy", None, i, u32p, u, u32p, u)
        self._proj = fn("raycast_proj", None, i, i, i, i, u32p, f32p, vp, vp, vp, i, i, i, i, f32p, f32p, f32p, f32p)
        self._counthole = fn("raycast_counthole", None, i, i, i, i, i, u32p, vp, u32p, i, i, i)
        self._sumids = fn("raycast_sumids", None, i, i, i, i, u32p, vp, u32p, i, i, i)
        self._writeids = fn("raycast_writeids", None, i, i, i, i, i, u32p, vp, u32p, i, i, i)
        self._holes = fn("raycast_holes", None, i, i, i, i, i, u32p, f32p, u32p, vp, u32p, u, i, i, i, i,
                         f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f, f)
        self._fine2 = fn("raycast_fine_2", None, i, i, i, i, i, u32p, f32p, u32p, u, i, i, i, i, i,
                         f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f, f)
        self._fine = fn("raycast_fine", None, i, i, i, i, u32p, f32p, u32p, u, i, i, i, i, i,
                        f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f, f)
        i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        self._fillhole = fn("raycast_fillhole", None, i, i, i, i, i, u32p, f32p, i32p, i32p, vp, i, i, i)
        self._fillhole2 = fn("raycast_fillhole2", None, i, i, i, i, u32p, vp, i, i, i)
        self._colorize = fn("raycast_colorize", None, i, i, i, i, i, u32p, u32p, i, i)
        self._max_threads = fn("max_threads", i)
        # multi-threaded forms for the timed CPU baseline (bench.py); raycast_proj_mt is the racy kernel run work-group-parallel
        self._memset_mt = fn("memset_mt", None, i, u32p, u, u, i)
        self._memcpy_mt = fn("memcpy_mt", None, i, u32p, u, u32p, u, i)
        self._proj_mt = fn("raycast_proj_mt", None, i, i, i, i, i, u32p, f32p, vp, vp, vp, i, i, i, i, f32p, f32p, f32p, f32p)
        self._fillhole2_mt = fn("raycast_fillhole2_mt", None, i, i, i, i, i, u32p, vp, i, i, i)

    # ---- octree build -------------------------------------------------------
    def build_octree(self, x, y, z, rgba):
        """set_voxel stream (in the given order) -> (compact uint32 array, octree_root_normal)."""
        x, y, z, rgba = (np.ascontiguousarray(a, dtype=np.uint32) for a in (x, y, z, rgba))
        self._reset()
        self._set_voxels(len(x), x, y, z, rgba)
        return self._finish()

    def build_octree_rle4(self, path, palette=0):
        """RLE4::load -> set_voxel -> convert_tree_blocks (src/raycast.h:13-46)."""
        with quiet_stdout():
            self._reset()
            self._load_rle4(os.fsencode(path), palette, 0, 0, 0)
            return self._finish()

    def _finish(self):
        root = self._convert()
        n = self._compact_words()
        arr = np.ctypeslib.as_array(self._compact_data(), shape=(n,)).copy()
        return arr, int(root)

    def num_voxels(self):
        return int(self._num_voxels())

    def num_nodes(self):
        return int(self._num_nodes())

    def max_threads(self):
        return int(self._max_threads())

    # ---- kernels; buffers are numpy arrays modified in place -------------------
    def memset(self, dst, dstofs_words, val, nwords, threads=1):
        if threads > 1:
            self._memset_mt(nwords, dst, dstofs_words, val, threads)
        else:
            self._memset(nwords, dst, dstofs_words, val)

    def memcpy(self, dst, dstofs_words, src, srcofs_words, nwords, threads=1):
        if threads > 1:
            self._memcpy_mt(nwords, dst, dstofs_words, src, srcofs_words, threads)
        else:
            self._memcpy(nwords, dst, dstofs_words, src, srcofs_words)

    def raycast_proj(self, screen, back, res_x, res_y, frame, ofs_add, m0, mx, my, mz, racy_threads=1, xbuf=None, ybuf=None):
        """racy_threads > 1: the kernel work-group-parallel with its payload race, as an OpenCL CPU runtime runs it -- for
        timing only; the serial form (default) is the defined outcome the parity checks use."""
        xb = xbuf.ctypes.data if xbuf is not None else None      # int32 motion-vector buffers (C restatement only: the reference
        yb = ybuf.ctypes.data if ybuf is not None else None      # keeps their producer commented out, kernel.cl:587-588)
        if racy_threads > 1:
            self._proj_mt(res_x, res_y, 16, 16, racy_threads, screen, back, xb, yb, None, res_x, res_y, frame, ofs_add,
                          _vec4(m0), _vec4(mx), _vec4(my), _vec4(mz))
        else:
            self._proj(res_x, res_y, 16, 16, screen, back, xb, yb, None, res_x, res_y, frame, ofs_add,
                       _vec4(m0), _vec4(mx), _vec4(my), _vec4(mz))

    def raycast_counthole(self, screen, idbuf, res_x, res_y, frame=0, threads=1):
        self._counthole(res_x // 16, res_y // 16, 16, 16, threads, screen, None, idbuf, res_x, res_y, frame)

    def raycast_sumids(self, screen, idbuf, res_x, res_y, frame=0):
        self._sumids(1, 1, 1, 1, screen, None, idbuf, res_x, res_y, frame)

    def raycast_writeids(self, screen, idbuf, res_x, res_y, frame=0, threads=1):
        self._writeids(res_x // 16, res_y // 16, 16, 16, threads, screen, None, idbuf, res_x, res_y, frame)

    def raycast_holes(self, screen, back, octree, idbuf, root, res_x, res_y, frame, idbuf_size,
                      m0, mx, my, mz, fovx=1.0, fovy=1.0, threads=1):
        z = _vec4([0])
        self._holes(idbuf_size, 1, 256, 1, threads, screen, back, octree, None, idbuf, root, res_x, res_y, frame,
                    idbuf_size, z, z, z, z, _vec4(m0), _vec4(mx), _vec4(my), _vec4(mz), fovx, fovy)

    def raycast_fine_2(self, screen, back, octree, root, res_x, res_y, frame, add_x, add_y,
                       m0, mx, my, mz, fovx=1.0, fovy=1.0, threads=1, gx=None, gy=None):
        z = _vec4([0])
        gx = res_x // 8 if gx is None else gx
        gy = res_y // 4 if gy is None else gy
        self._fine2(gx, gy, 16, 16, threads, screen, back, octree, root, res_x, res_y, frame, add_x, add_y,
                    z, z, z, z, _vec4(m0), _vec4(mx), _vec4(my), _vec4(mz), fovx, fovy)

    def raycast_fine(self, screen, back, octree, root, res_x, res_y, frame, add_x, add_y, m0, mx, my, mz,
                     fovx=1.0, fovy=1.0, gx=None, gy=None):
        """kernel.cl:696-843 with the launch geometry of its (disabled) call site, src/raycast.h:238."""
        z = _vec4([0])
        gx = res_x // 4 if gx is None else gx
        gy = res_y // 4 if gy is None else gy
        self._fine(gx, gy, 16, 16, screen, back, octree, root, res_x, res_y, frame, add_x, add_y,
                   z, z, z, z, _vec4(m0), _vec4(mx), _vec4(my), _vec4(mz), fovx, fovy)

    def raycast_fillhole(self, screen, back, xbuf, ybuf, res_x, res_y, frame=0, threads=1):
        """kernel.cl:342-401 with the launch geometry of its (disabled) call site, src/raycast.h:209."""
        self._fillhole(res_x, res_y, 16, 16, threads, screen, back, xbuf, ybuf, None, res_x, res_y, frame)

    def raycast_fillhole2(self, screen, res_x, res_y, frame=0, threads=1):
        if threads > 1:
            self._fillhole2_mt(res_x, res_y, 16, 16, threads, screen, None, res_x, res_y, frame)
        else:
            self._fillhole2(res_x, res_y, 16, 16, screen, None, res_x, res_y, frame)

    def raycast_colorize(self, screen, tex, w, h, threads=1):
        self._colorize(w, h, 16, 16, threads, screen, tex, w, h)


_cache = {}


def get(prefix):
    if prefix not in _cache:
        _cache[prefix] = CpuOracle(prefix)
    return _cache[prefix]
