"""Python face of the frame driver (include/svo_raycast.h, host/raycast_host.cpp -- the C mirror of the reference's
src/raycast.h) on top of ocl.py (-> libsvo_b200.so).

Same entry points as the reference: ``raycast_init()`` (:61-91), ``raycast_draw(res_x, res_y)`` (:93-510),
``raycast_exit()`` (:511-517).  What the window supplied in the reference (MOUSE_X/Y -> rot, WASD -> pos, :113-133) is set
explicitly with ``set_camera(pos, rot)``; the GL blit (:449-472) is replaced by ``read_frame()`` (headless framebuffer).

There is ONE restatement of the reference's launch sequence, the C driver: ``raycast_draw`` calls ``svo_raycast_draw``, which
in ``mode="reference"`` issues the reference's 13 launches one by one through svo_begin / svo_param / svo_end with the
argument lists of the call sites (including the blocking 4-byte readback of idbuf_size, :298), in ``mode="fused"`` the same
frame through ``svo_frame_fused`` (no host readback; results identical) and in ``mode="pingpong"`` the fused frame with
SVO_FRAME_PINGPONG (no cache copy, frames alternate between buffers 0 and 2; ``last_slot()`` says where the frame is).
This module adds what a Python host needs around it: the buffers as ``ocl.Mem`` objects, camera blocks prepared ahead of
a frame loop (``prepare_params`` / ``draw_prepared`` / ``draw_present``), and read-backs for parity checks.
"""
import ctypes as C
import math

import numpy as np

from . import ocl

HOLE = 0xFFFFFF00
f32 = np.float32

# globals of src/main.cpp:17-18,49-50 / src/raycast.h
WINDOW_WIDTH_MAX = 2048
WINDOW_HEIGHT_MAX = 1080
OCTREE_DEPTH = 11


class State:
    pass


S = State()
S.ready = False


def rotation_matrix(rot):
    """matrix44 m; m.rotate_z(rot.z); m.rotate_x(rot.x); m.rotate_y(rot.y)  (src/raycast.h:121-124,
    ext/mathlib/_matrix44.h:529-577), single precision."""
    m = np.eye(4, dtype=np.float32)
    for axis, (a, b) in ((2, (0, 1)), (0, (1, 2)), (1, (0, 2))):
        ang = float(f32(rot[axis]))
        c, s = f32(math.cos(ang)), f32(math.sin(ang))
        for i in range(4):
            va, vb = m[i, a], m[i, b]
            if axis == 1:       # rotate_y: m[i][0] = mi0*c + mi2*s ; m[i][2] = mi0*-s + mi2*c
                m[i, a] = f32(va * c) + f32(vb * s)
                m[i, b] = f32(va * -s) + f32(vb * c)
            else:               # rotate_x / rotate_z: first = a*c + b*-s ; second = a*s + b*c
                m[i, a] = f32(va * c) + f32(vb * -s)
                m[i, b] = f32(va * s) + f32(vb * c)
    return m


def wrap_pos(pos, depth=None):
    """src/raycast.h:136-145."""
    depth = OCTREE_DEPTH if depth is None else depth
    dim2 = f32((1 << depth) * 2)
    p = np.asarray(pos, dtype=np.float32) * f32(16.0)
    for k in range(3):
        while p[k] < 0:
            p[k] = p[k] + dim2
        while p[k] >= dim2:
            p[k] = p[k] - dim2
    return (p / f32(16.0)).astype(np.float32)


_MODES = {"reference": 0, "fused": 1, "pingpong": 2}          # SVO_MODE_* of include/svo_raycast.h
_sig = ocl._sig
_rc_init_words = _sig("svo_raycast_init_words", C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_rc_exit = _sig("svo_raycast_exit", None)
_rc_set_camera = _sig("svo_raycast_set_camera", None, C.c_void_p, C.c_void_p)
_rc_draw = _sig("svo_raycast_draw", None, C.c_int, C.c_int, C.c_int)
_rc_set_frame = _sig("svo_raycast_set_frame", None, C.c_int)
_rc_set_mode = _sig("svo_raycast_set_mode", None, C.c_int)
_rc_set_cache_rotation = _sig("svo_raycast_set_cache_rotation", C.c_int, C.c_int)
_rc_set_quality_passes = _sig("svo_raycast_set_quality_passes", C.c_int, C.c_int)
_rc_idbuf_size = _sig("svo_raycast_idbuf_size", C.c_int)
_rc_last_camera = _sig("svo_raycast_last_camera", None, C.c_void_p)
_rc_mem = _sig("svo_raycast_mem", C.c_void_p, C.c_char_p)


def raycast_init(octree_words, octree_root_normal, max_w=None, max_h=None, depth=11, device=0, mode="fused", cache_rotation=False):
    """src/raycast.h:61-91 (octree_init is replaced by the caller handing in the compact octree): svo_raycast_init_words.
    cache_rotation ("reference" and "fused" modes): copy target `((frame>>4)%2)+1`, the variant the reference keeps in a
    comment at src/raycast.h:395, instead of the hard-wired 2 -- cache buffers 1 and 2 then both hold real frames."""
    if cache_rotation and mode == "pingpong":
        raise ValueError("cache_rotation: ping-pong mode has no cache copy to rotate")
    global WINDOW_WIDTH_MAX, WINDOW_HEIGHT_MAX, OCTREE_DEPTH
    if max_w:
        WINDOW_WIDTH_MAX = max_w
    if max_h:
        WINDOW_HEIGHT_MAX = max_h
    OCTREE_DEPTH = depth
    octree_words = np.ascontiguousarray(octree_words, dtype=np.uint32)
    rc_ = _rc_init_words(octree_words.ctypes.data, len(octree_words), int(octree_root_normal), depth, WINDOW_WIDTH_MAX, WINDOW_HEIGHT_MAX,
                         device, _MODES[mode])
    ocl._check()
    if rc_:
        raise RuntimeError(f"svo_raycast_init_words failed ({rc_})")
    size = WINDOW_WIDTH_MAX * WINDOW_HEIGHT_MAX

    def mem(name, nbytes):
        return ocl.Mem(_rc_mem(name.encode()), nbytes)                             # owned by the C driver (raycast_exit frees them)

    S.mem_octree = mem("octree", octree_words.nbytes)                              # :68
    S.mem_bvh_nodes = S.mem_bvh_childs = S.mem_stack = None                        # never allocated / never touched (:71-77)
    S.mem_backbuffer = mem("backbuffer", size * 16 * 4)                            # :82
    S.mem_screenbuffer = mem("screenbuffer", size * 4 * 4)                         # :85
    S.mem_screenbuffer_tex = mem("screenbuffer_tex", size * 4)                     # PBO stand-in (:88-89)
    S.mem_screenbuffer_tex2 = None                                                 # second colorize target (draw_present)
    S.present_tex = []                                                             # colorize targets of the frames in flight
    S.mem_x = S.mem_y = None                                                       # dead kernel arguments (:170-171)
    S.mem_z = mem("z", 4 * size)                                                   # :172 (only ever memset)
    S.mem_idbuffer = mem("idbuffer", (size + 2 * (WINDOW_WIDTH_MAX // 16) * (WINDOW_HEIGHT_MAX // 16)) * 4)
    S.octree_root_normal = int(octree_root_normal)
    S.cache_rotation = bool(cache_rotation)
    if cache_rotation and _rc_set_cache_rotation(1):
        raise RuntimeError("svo_raycast_set_cache_rotation failed")
    S.frame = -1
    S.pos = np.array([1, 50, 1], dtype=np.float32)                                 # :113
    S.rot = np.array([0.0001, 0, 0], dtype=np.float32)                             # :114
    S.mode = mode
    S.idbuf_size = 0
    S.k = {}
    S.ready = True


def set_quality_passes(on=True):
    """mode="reference": enable raycast_fillhole every 4th frame (src/raycast.h:205-219, if(0) in the reference) together with
    the motion-vector producer in raycast_proj it needs (kernel.cl:587-588, commented out there)."""
    if _rc_set_quality_passes(1 if on else 0):
        raise ValueError("quality passes need mode='reference'")
    size = WINDOW_WIDTH_MAX * WINDOW_HEIGHT_MAX
    S.mem_x = ocl.Mem(_rc_mem(b"x"), 4 * size) if on else None
    S.mem_y = ocl.Mem(_rc_mem(b"y"), 4 * size) if on else None


def _kernel(name):
    if name not in S.k:
        S.k[name] = ocl.ocl_get_kernel(name)       # function-local statics in the reference
    return S.k[name]


def set_camera(pos, rot):
    S.pos = np.asarray(pos, dtype=np.float32).copy()
    S.rot = np.asarray(rot, dtype=np.float32).copy()


def tile_origin(frame, res_x, res_y):
    """src/raycast.h:363-364: rotating 8x4 tile refresh schedule."""
    return (res_x // 8) * (frame & 7), (res_y // 4) * ((frame >> 3) & 3)


def raycast_draw(res_x, res_y):
    """One frame through the C driver (svo_raycast_draw, synchronous); returns the camera actually used (after the wrap of
    :136-145)."""
    _rc_set_mode(_MODES[S.mode])
    _rc_set_frame(S.frame)                          # this module's counter is the master (draw_prepared issues frames itself)
    pos = np.ascontiguousarray(S.pos, dtype=np.float32)
    rot = np.ascontiguousarray(S.rot, dtype=np.float32)
    _rc_set_camera(pos.ctypes.data, rot.ctypes.data)
    _rc_draw(res_x, res_y, 1)
    ocl._check()
    S.frame += 1
    cam = np.zeros(28, dtype=np.float32)
    _rc_last_camera(cam.ctypes.data)
    S.pos = cam[:3].copy()                          # wrapped (:136-145)
    S.idbuf_size = _rc_idbuf_size() if S.mode == "reference" else 0
    S.last_camera = dict(pos=cam[:3].copy(), v0=cam[:4].copy(), rows=[cam[4 + 4 * i:8 + 4 * i].copy() for i in range(3)],
                         cols=[cam[16 + 4 * i:20 + 4 * i].copy() for i in range(3)])
    return S.last_camera


def prepare_params(res_x, res_y, frame):
    """The per-frame camera block of the fused path (struct svo_frame_params) for the camera set with set_camera()."""
    m = rotation_matrix(S.rot)
    pos = wrap_pos(S.pos)
    p = ocl.FrameParams()
    p.res_x, p.res_y, p.frame = res_x, res_y, frame
    p.v0[:] = [float(pos[0]), float(pos[1]), float(pos[2]), 1.0]
    for i in range(3):
        p.rows[i][:] = m[i, :].tolist()
        p.cols[i][:] = m[:, i].tolist()
    p.fovx = p.fovy = 1.0
    p.flags = ocl.FRAME_PINGPONG if S.mode == "pingpong" else ocl.FRAME_CACHE_ROTATION if getattr(S, "cache_rotation", False) else 0
    return p


def draw_prepared(p, sync=True):
    """raycast_draw() of the fused path with a camera block prepared in advance (keeps Python out of the frame loop)."""
    S.frame = p.frame
    ocl.frame_fused(S.mem_screenbuffer, S.mem_backbuffer, S.mem_idbuffer, S.mem_octree, S.octree_root_normal,
                    S.mem_screenbuffer_tex, p)
    if sync:
        ocl.ocl_end_all_kernels()


def prepare_present(depth):
    """Allocate the colorize targets draw_present() rotates through (one per frame in flight) ahead of time: a cudaMalloc
    inside a frame loop stalls it for anything between 0.1 and 100 ms."""
    while len(S.present_tex) < depth:
        S.present_tex.append(S.mem_screenbuffer_tex if not S.present_tex else ocl.ocl_malloc(S.mem_screenbuffer_tex.size))
    S.mem_screenbuffer_tex2 = S.present_tex[1] if depth > 1 else None


def draw_present(p, host_frames, rgb24=False):
    """Pipelined headless frame: with D = len(host_frames) (2..4) buffers, frame p.frame is rendered into colorize target
    k = p.frame % D and its read-back into host_frames[k] (page-locked, ocl.host_alloc) is queued on the copy stream, so
    it overlaps the following frames.  The caller owns the image after ocl.present_wait(k) and must have consumed it
    before frame p.frame + D is issued.  rgb24: the frame arrives as R,G,B bytes (PPM payload, 3 bytes per pixel) instead
    of the PBO's 0x00RRGGBB words."""
    depth = len(host_frames)
    prepare_present(depth)
    k = p.frame % depth
    tex = S.present_tex[k]
    S.frame = p.frame
    if rgb24 == "pack":                                  # word image + k_pack_rgb24 on the copy stream (A/B of the path below)
        ocl.frame_fused(S.mem_screenbuffer, S.mem_backbuffer, S.mem_idbuffer, S.mem_octree, S.octree_root_normal, tex, p)
        ocl.present_rgb24_async(host_frames[k], tex, p.res_x * p.res_y, k)
        return k
    if rgb24:
        # the kernels that produce the pixels store R,G,B bytes themselves (SVO_FRAME_TEX_RGB24): the read-back is a plain copy
        q = type(p).from_buffer_copy(p)
        q.flags |= ocl.FRAME_TEX_RGB24
        ocl.frame_fused(S.mem_screenbuffer, S.mem_backbuffer, S.mem_idbuffer, S.mem_octree, S.octree_root_normal, tex, q)
        ocl.present_async(host_frames[k], tex, p.res_x * p.res_y * 3, k)
        return k
    ocl.frame_fused(S.mem_screenbuffer, S.mem_backbuffer, S.mem_idbuffer, S.mem_octree, S.octree_root_normal, tex, p)
    ocl.present_async(host_frames[k], tex, p.res_x * p.res_y * 4, k)
    return k


def reset_frames():
    S.frame = -1


def full_raycast_ms(res_x, res_y, repeats=5):
    """Device time (ms, median) of raycast_fine_2 over the whole screen for the camera set with set_camera()."""
    m = rotation_matrix(S.rot)
    pos = wrap_pos(S.pos)
    v0 = np.array([pos[0], pos[1], pos[2], 1.0], dtype=np.float32)
    cols = [m[:, i].copy() for i in range(3)]
    dead = (0.0, 0.0, 0.0, 0.0)
    times = []
    for r in range(repeats + 2):
        ocl.event_record(2)
        ocl.ocl_begin(_kernel("raycast_fine_2"), res_x, res_y, 16, 16)
        for a in (S.mem_screenbuffer, S.mem_backbuffer, S.mem_octree):
            ocl.ocl_param(a)
        ocl.ocl_param(C.c_uint32(S.octree_root_normal))
        for a in (res_x, res_y, 0, 0, 0):
            ocl.ocl_param(C.c_int(a))
        for a in (dead, dead, dead, dead, v0, cols[0], cols[1], cols[2]):
            ocl.ocl_param(a)
        ocl.ocl_param(C.c_float(1.0)); ocl.ocl_param(C.c_float(1.0))
        ocl.ocl_end()
        ocl.event_record(3)
        ms = ocl.event_elapsed_ms(2, 3)
        if r >= 2:
            times.append(ms)
    return float(np.median(times))


def full_raycast_enqueue(res_x, res_y):
    """raycast_fine_2 over the whole screen for the camera set with set_camera(); asynchronous (view-parallel batches)."""
    m = rotation_matrix(S.rot)
    pos = wrap_pos(S.pos)
    v0 = np.array([pos[0], pos[1], pos[2], 1.0], dtype=np.float32)
    dead = (0.0, 0.0, 0.0, 0.0)
    ocl.ocl_begin(_kernel("raycast_fine_2"), res_x, res_y, 16, 16)
    for a in (S.mem_screenbuffer, S.mem_backbuffer, S.mem_octree):
        ocl.ocl_param(a)
    ocl.ocl_param(C.c_uint32(S.octree_root_normal))
    for a in (res_x, res_y, 0, 0, 0):
        ocl.ocl_param(C.c_int(a))
    for a in (dead, dead, dead, dead, v0, m[:, 0].copy(), m[:, 1].copy(), m[:, 2].copy()):
        ocl.ocl_param(a)
    ocl.ocl_param(C.c_float(1.0)); ocl.ocl_param(C.c_float(1.0))
    ocl.ocl_end()


def full_raycast_batch(res_x, res_y, poses):
    """Full raycasts for a list of (pos, rot) cameras through svo_raycast_batch; asynchronous (view-parallel batches)."""
    plist = []
    for pos, rot in poses:
        set_camera(pos, rot)
        plist.append(prepare_params(res_x, res_y, 0))
    ocl.raycast_batch(S.mem_screenbuffer, S.mem_backbuffer, S.mem_octree, S.octree_root_normal, res_x, res_y, plist)


def idbuf_size():
    return ocl.frame_idbuf_size() if S.mode in ("fused", "pingpong") else S.idbuf_size


def last_slot():
    """Buffer index the last frame was rendered into (0, or 0/2 alternating in ping-pong mode)."""
    return ocl.frame_last_slot() if S.mode == "pingpong" else 0


def read_frame(res_x, res_y):
    """Headless framebuffer: the colorized 0x00RRGGBB image the reference blits through the PBO (:449-455)."""
    return S.mem_screenbuffer_tex.to_numpy(np.uint32, res_x * res_y).reshape(res_y, res_x)


def read_buffers(res_x, res_y):
    """(screen[4N], back[16N] float32, idbuf) of the current resolution, for parity checks."""
    n = res_x * res_y
    nb = (res_x // 16) * (res_y // 16)
    screen = S.mem_screenbuffer.to_numpy(np.uint32, 4 * n)
    back = S.mem_backbuffer.to_numpy(np.float32, 16 * n)
    idb = S.mem_idbuffer.to_numpy(np.uint32, 2 * nb + n)
    return screen, back, idb


def raycast_exit():
    """src/raycast.h:511-517."""
    if not S.ready:
        return
    for m in getattr(S, "present_tex", [])[1:]:
        m.free()
    S.present_tex, S.mem_screenbuffer_tex2 = [], None
    for name in ("mem_octree", "mem_backbuffer", "mem_screenbuffer", "mem_screenbuffer_tex", "mem_z", "mem_idbuffer"):
        setattr(S, name, None)                      # the C driver owns and frees them
    _rc_exit()
    S.ready = False
