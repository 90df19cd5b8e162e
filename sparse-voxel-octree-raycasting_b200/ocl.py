"""Python mirror of the reference's device-runtime shim src/ocl.h, bound to libsvo_b200.so.

Same function names, argument order and meaning as src/ocl.h (ocl_init :57, ocl_get_kernel :148,
ocl_malloc :200, ocl_copy_to_host :215, ocl_begin/ocl_param/ocl_end :229/:238/:268,
ocl_begin_all_kernels/ocl_end_all_kernels :246/:253, ocl_memcpy :285, ocl_memset :299,
ocl_round_up :188); underneath is the C ABI of include/svo_b200.h.  Errors abort the process like the
reference's CL_CHECK (src/ocl.h:29-41) unless ``Device.errors_return()`` was selected, in which case they
raise ``RuntimeError``.
"""
import ctypes as C
import os

import numpy as np

# SVO_B200_LIB: another build of the same library (kernel tuning experiments, tools/variants.sh)
LIB_PATH = os.environ.get("SVO_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsvo_b200.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no CPU fallback.")
lib = C.CDLL(LIB_PATH)


class FrameParams(C.Structure):
    """struct svo_frame_params (include/svo_b200.h)."""
    _fields_ = [("res_x", C.c_int), ("res_y", C.c_int), ("frame", C.c_int),
                ("v0", C.c_float * 4), ("rows", (C.c_float * 4) * 3), ("cols", (C.c_float * 4) * 3),
                ("fovx", C.c_float), ("fovy", C.c_float), ("flags", C.c_int)]


FRAME_PINGPONG = 1
FRAME_CACHE_ROTATION = 2
FRAME_TEX_RGB24 = 4      # colorized frame as packed R,G,B bytes, stored by the producers (svo_b200.h)


def _sig(name, res, *args):
    f = getattr(lib, name)
    f.restype = res
    f.argtypes = list(args)
    return f


_vp, _sz, _u32, _i = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
_svo_init = _sig("svo_init", _i, _i)
_svo_exit = _sig("svo_exit", None)
_svo_set_error_mode = _sig("svo_set_error_mode", None, _i)
_svo_last_error = _sig("svo_last_error", _i)
_svo_last_error_string = _sig("svo_last_error_string", C.c_char_p)
_svo_clear_error = _sig("svo_clear_error", None)
_svo_ctx_create = _sig("svo_ctx_create", _vp, _i)
_svo_ctx_destroy = _sig("svo_ctx_destroy", None, _vp)
_svo_ctx_set_current = _sig("svo_ctx_set_current", None, _vp)
_svo_ctx_get_current = _sig("svo_ctx_get_current", _vp)
_svo_ctx_stream = _sig("svo_ctx_stream", _vp, _vp)
_svo_device_count = _sig("svo_device_count", _i)
_svo_set_octree_depth = _sig("svo_set_octree_depth", None, _i)
_svo_get_octree_depth = _sig("svo_get_octree_depth", _i)
_svo_malloc = _sig("svo_malloc", _vp, _sz, _vp)
_svo_free = _sig("svo_free", None, _vp)
_svo_copy_to_host = _sig("svo_copy_to_host", None, _vp, _vp, _sz, _sz)
_svo_copy_to_device = _sig("svo_copy_to_device", None, _vp, _sz, _vp, _sz)
_svo_memcpy = _sig("svo_memcpy", None, _vp, _u32, _vp, _u32, _u32)
_svo_memset = _sig("svo_memset", None, _vp, _u32, _u32, _u32)
_svo_mem_device_ptr = _sig("svo_mem_device_ptr", _vp, _vp)
_svo_mem_size = _sig("svo_mem_size", _sz, _vp)
_svo_get_kernel = _sig("svo_get_kernel", _vp, C.c_char_p)
_svo_begin = _sig("svo_begin", None, C.POINTER(_vp), _i, _i, _i, _i)
_svo_param = _sig("svo_param", None, _sz, _vp)
_svo_end = _sig("svo_end", None)
_svo_begin_all = _sig("svo_begin_all_kernels", None)
_svo_end_all = _sig("svo_end_all_kernels", None)
_svo_round_up = _sig("svo_round_up", _sz, _i, _i)
_svo_launch_count = _sig("svo_launch_count", C.c_uint64)
_svo_frame_fused = _sig("svo_frame_fused", None, _vp, _vp, _vp, _vp, _u32, _vp, C.POINTER(FrameParams))
_svo_raycast_batch = _sig("svo_raycast_batch", None, _vp, _vp, _vp, _u32, _i, _i, _i, C.POINTER(FrameParams))
_svo_frame_idbuf_size = _sig("svo_frame_idbuf_size", _i)
_svo_frame_last_slot = _sig("svo_frame_last_slot", _i)
_svo_frame_deferred_count = _sig("svo_frame_deferred_count", C.c_uint64)
_svo_frame_early_count = _sig("svo_frame_early_count", C.c_uint64)

_svo_event_record = _sig("svo_event_record", None, _i)
_svo_event_elapsed_ms = _sig("svo_event_elapsed_ms", C.c_float, _i, _i)
_svo_profile_enable = _sig("svo_profile_enable", None, _i)
_svo_profile_reset = _sig("svo_profile_reset", None)
_svo_profile_get = _sig("svo_profile_get", _i, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64))
_svo_profile_names = _sig("svo_profile_names", _i, C.c_char_p, _sz)
_svo_profile_timeline = _sig("svo_profile_timeline", _i, C.c_char_p, _sz)
_svo_host_alloc = _sig("svo_host_alloc", _vp, _sz)
_svo_host_free = _sig("svo_host_free", None, _vp)
_svo_copy_to_host_async = _sig("svo_copy_to_host_async", None, _vp, _vp, _sz, _sz)

_svo_present_async = _sig("svo_present_async", None, _vp, _vp, _sz, _i)
_svo_present_rgb24_async = _sig("svo_present_rgb24_async", None, _vp, _vp, _sz, _i)
_svo_present_wait = _sig("svo_present_wait", None, _i)

_raise_errors = False


def _check():
    if _raise_errors and _svo_last_error():
        msg = _svo_last_error_string().decode()
        _svo_clear_error()
        raise RuntimeError("svo_b200: " + msg)


class Device:
    """Process-wide settings of the library."""

    @staticmethod
    def errors_return():
        """Turn fatal errors into RuntimeError (tests); the default aborts like the reference."""
        global _raise_errors
        _raise_errors = True
        _svo_set_error_mode(1)

    @staticmethod
    def count():
        return _svo_device_count()


class Mem:
    """cl_mem analogue: an opaque device buffer handle."""

    def __init__(self, handle, size):
        self.handle, self.size = handle, size

    @property
    def _as_parameter_(self):
        return C.c_void_p(self.handle)

    def device_ptr(self):
        return _svo_mem_device_ptr(self.handle)

    def free(self):
        if self.handle:
            _svo_free(self.handle)
            self.handle = None

    def to_numpy(self, dtype=np.uint32, count=None, offset_bytes=0):
        itemsize = np.dtype(dtype).itemsize
        count = (self.size - offset_bytes) // itemsize if count is None else count
        out = np.empty(count, dtype=dtype)
        ocl_copy_to_host(out, self, count * itemsize, offset_bytes)
        return out


def ocl_init(device=0):
    rc = _svo_init(device)
    _check()
    if rc:
        raise RuntimeError("svo_init failed: " + _svo_last_error_string().decode())


def ocl_exit():
    _svo_exit()


def set_octree_depth(depth):
    _svo_set_octree_depth(depth)
    _check()


def launch_count():
    return int(_svo_launch_count())


def ocl_get_kernel(name):
    h = _svo_get_kernel(name.encode())
    _check()
    return C.c_void_p(h)


def ocl_malloc(size, ptr=None):
    """size in bytes; ptr: optional numpy array copied at creation (CL_MEM_COPY_HOST_PTR, src/ocl.h:202)."""
    if ptr is not None:
        ptr = np.ascontiguousarray(ptr)
        assert ptr.nbytes >= size
        h = _svo_malloc(size, ptr.ctypes.data)
    else:
        h = _svo_malloc(size, None)
    _check()
    return Mem(h, size) if h else None


def ocl_copy_to_host(dst, src, size, srcofs=0):
    assert dst.flags["C_CONTIGUOUS"] and dst.nbytes >= size
    _svo_copy_to_host(dst.ctypes.data, src.handle, size, srcofs)
    _check()


def ocl_copy_to_device(dst, dstofs, src, size=None):
    src = np.ascontiguousarray(src)
    size = src.nbytes if size is None else size
    _svo_copy_to_device(dst.handle, dstofs, src.ctypes.data, size)
    _check()


_cur = None


def ocl_begin(kernel, globalx, globaly, localx, localy):
    global _cur
    _cur = C.c_void_p(kernel.value)
    _svo_begin(C.byref(_cur), globalx, globaly, localx, localy)
    _check()


def ocl_param(value):
    """One positional kernel argument: a Mem (or None for an unallocated cl_mem), a ctypes scalar, or a
    sequence of 4 floats (float4)."""
    if value is None or isinstance(value, Mem):
        v = C.c_void_p(value.handle if value is not None else None)
    elif isinstance(value, (C._SimpleCData, C.Array)):
        v = value
    else:
        a = np.zeros(4, dtype=np.float32)
        vv = np.asarray(value, dtype=np.float32).ravel()
        a[:len(vv)] = vv
        v = (C.c_float * 4)(*a.tolist())
    _svo_param(C.sizeof(v), C.byref(v))
    _check()


def ocl_end():
    _svo_end()
    _check()


def ocl_begin_all_kernels():
    _svo_begin_all()


def ocl_end_all_kernels():
    _svo_end_all()
    _check()


def ocl_memcpy(dst, dstofs, src, srcofs, size):
    _svo_memcpy(dst.handle, dstofs, src.handle, srcofs, size)
    _check()


def ocl_memset(dst, dstofs, val, size):
    _svo_memset(dst.handle, dstofs, val, size)
    _check()


_svo_debug_set = _sig("svo_debug_set", _i, C.c_char_p, _i)


def debug_set(name, value=1):
    """svo_debug_set: schedule A/B switches of the fused frame (development aid; results never change)."""
    if _svo_debug_set(name.encode(), int(value)):
        raise ValueError(f"unknown debug switch {name!r}")


def ocl_round_up(group_size, global_size):
    return int(_svo_round_up(group_size, global_size))


def frame_fused(screen, back, idbuf, octree, root, tex, params):
    _svo_frame_fused(screen.handle, back.handle, idbuf.handle, octree.handle, root, tex.handle if tex else None,
                     C.byref(params))
    _check()


def raycast_batch(screen, back, octree, root, res_x, res_y, params_list):
    """Full raycasts of many cameras (svo_raycast_batch): camera i -> buffer 0 / 2 for even / odd i, consecutive cameras overlap."""
    arr = (FrameParams * len(params_list))(*params_list)
    _svo_raycast_batch(screen.handle, back.handle, octree.handle, root, res_x, res_y, len(params_list), arr)
    _check()


def frame_last_slot():
    return int(_svo_frame_last_slot())


def frame_deferred_count():
    """Fused frames whose reprojection pass carried the previous frame's cache copy (see svo_frame_deferred_count)."""
    return int(_svo_frame_deferred_count())


def frame_early_count():
    """Of those, the frames whose reprojection ran as early pass + list pass (see svo_frame_early_count)."""
    return int(_svo_frame_early_count())


def frame_idbuf_size():
    return int(_svo_frame_idbuf_size())


# ---- timing / profiling / pinned host memory (extensions of include/svo_b200.h) ----------------------------------
def event_record(slot):
    _svo_event_record(slot)


def event_elapsed_ms(a, b):
    return float(_svo_event_elapsed_ms(a, b))


def profile_enable(on):
    _svo_profile_enable(1 if on else 0)


def profile_reset():
    _svo_profile_reset()


def profile_all():
    """{cuda kernel name: (total ms, launches)} since the last reset."""
    buf = C.create_string_buffer(4096)
    _svo_profile_names(buf, 4096)
    out = {}
    for name in buf.value.decode().split():
        ms, cnt = C.c_double(), C.c_uint64()
        _svo_profile_get(name.encode(), C.byref(ms), C.byref(cnt))
        out[name] = (ms.value, cnt.value)
    return out


def profile_timeline():
    """[(cuda kernel name, start_us, end_us)] of the launches since the last flush, all streams, relative to the first."""
    buf = C.create_string_buffer(1 << 18)
    _svo_profile_timeline(buf, 1 << 18)
    out = []
    for line in buf.value.decode().splitlines():
        name, a, b = line.split()
        out.append((name, float(a), float(b)))
    return out


def host_alloc(nbytes):
    """Page-locked host buffer as a writable memoryview-compatible ctypes array."""
    p = _svo_host_alloc(nbytes)
    _check()
    if not p:
        raise MemoryError("svo_host_alloc failed")
    return (C.c_uint8 * nbytes).from_address(p)


def host_free(buf):
    """Release a buffer returned by host_alloc (the ctypes array must not be used afterwards)."""
    _svo_host_free(C.addressof(buf))


def copy_to_host_async(dst, src, size, srcofs=0):
    _svo_copy_to_host_async(C.addressof(dst), src.handle, size, srcofs)
    _check()


def present_async(dst, src, size, slot):
    """Frame read-back on the copy stream (overlaps the next frame); present_wait(slot) blocks until it has landed."""
    _svo_present_async(C.addressof(dst), src.handle, size, slot)
    _check()


def present_rgb24_async(dst, src, npixels, slot):
    """present_async with the frame packed to R,G,B bytes on the device first (npixels*3 bytes land in dst)."""
    _svo_present_rgb24_async(C.addressof(dst), src.handle, npixels, slot)
    _check()


def present_wait(slot):
    _svo_present_wait(slot)
    _check()
