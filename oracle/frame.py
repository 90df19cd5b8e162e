"""Headless restatement of the reference's per-frame driver -- TEST INFRASTRUCTURE ONLY.

Follows src/raycast.h:93-438 (``raycast_draw``) launch by launch on a CpuOracle
(oracle/binding.py): same kernels, same argument values, same order.  The camera
math follows ext/mathlib/_matrix44.h:529-577 (rotate_x/y/z), :334-345 (transpose)
and :863-883 (row-vector * matrix).  GL/SDL calls are dropped; mouse/keyboard are
replaced by an explicit (pos, rot) per frame.
"""
import math
import numpy as np

HOLE = 0xFFFFFF00
f32 = np.float32


def rotation_matrix(rot):
    """m = I.rotate_z(rot.z).rotate_x(rot.x).rotate_y(rot.y)   (src/raycast.h:121-124)."""
    m = np.eye(4, dtype=np.float32)

    def cs(a):  # nmath.h:34-35: float(cos(x))
        return f32(math.cos(float(f32(a)))), f32(math.sin(float(f32(a))))

    c, s = cs(rot[2])                      # _matrix44.h:565-577
    for i in range(4):
        a0, a1 = m[i, 0], m[i, 1]
        m[i, 0] = f32(a0 * c) + f32(a1 * -s)
        m[i, 1] = f32(a0 * s) + f32(a1 * c)
    c, s = cs(rot[0])                      # _matrix44.h:529-541
    for i in range(4):
        a1, a2 = m[i, 1], m[i, 2]
        m[i, 1] = f32(a1 * c) + f32(a2 * -s)
        m[i, 2] = f32(a1 * s) + f32(a2 * c)
    c, s = cs(rot[1])                      # _matrix44.h:547-559
    for i in range(4):
        a0, a2 = m[i, 0], m[i, 2]
        m[i, 0] = f32(a0 * c) + f32(a2 * s)
        m[i, 2] = f32(a0 * -s) + f32(a2 * c)
    return m


def wrap_pos(pos, depth=11):
    """src/raycast.h:136-145: wrap pos*16 into [0, 2*octree_dim)."""
    dim2 = f32((1 << depth) * 2)
    p = np.asarray(pos, dtype=np.float32) * f32(16.0)
    for k in range(3):
        while p[k] < 0:
            p[k] = p[k] + dim2
        while p[k] >= dim2:
            p[k] = p[k] - dim2
    return (p / f32(16.0)).astype(np.float32)


def camera_args(pos, rot, depth=11):
    """-> dict(v0, rows=(vx,vy,vz) for raycast_proj, cols=(vx,vy,vz) for the ray kernels)."""
    m = rotation_matrix(rot)
    pos = wrap_pos(pos, depth)
    v0 = np.array([pos[0], pos[1], pos[2], 1.0], dtype=np.float32)       # vec4f v0=pos (w=1)
    rows = tuple(m[i, :].copy() for i in range(3))    # m*vec4f(1,0,0,0) = row 0 ... (:160-162)
    cols = tuple(m[:, i].copy() for i in range(3))    # after m.transpose()         (:322-325)
    return dict(pos=pos, v0=v0, rows=rows, cols=cols, m=m)


class OracleFrame:
    """State of the reference's frame loop (4 colour + 4 coordinate buffers, id buffer, frame counter)."""

    def __init__(self, orc, octree, root, res_x, res_y, threads=1, depth=11, cache_rotation=False, timed=False, quality_passes=False):
        """cache_rotation: the copy target the reference has in a comment, `((frame>>4)%2)+1` (src/raycast.h:395), instead of
        the hard-wired 2 -- the variant that makes the triple buffer real."""
        self.cache_rotation = cache_rotation
        # timed=True (bench.py's CPU baseline): every kernel work-group-parallel on `threads` host threads, the way an OpenCL
        # CPU runtime would run the frame -- raycast_proj with its payload race.  The result is then not the defined
        # (serial) outcome; parity checks use timed=False.
        self.timed = timed
        # quality_passes=True: the depth-discontinuity hole punch the reference keeps behind `if(0)` (raycast_fillhole,
        # src/raycast.h:205-219, every 4th frame) together with the motion-vector producer it needs, which the reference keeps
        # commented out in raycast_proj (kernel.cl:587-588) -- SURVEY.md 8(f) rank 4.  C restatement only.
        self.quality_passes = quality_passes
        self.xbuf = np.zeros(res_x * res_y, dtype=np.int32) if quality_passes else None   # mem_x / mem_y (:170-171)
        self.ybuf = np.zeros(res_x * res_y, dtype=np.int32) if quality_passes else None
        self.o, self.octree, self.root = orc, octree, root
        self.res_x, self.res_y, self.threads, self.depth = res_x, res_y, threads, depth
        n = res_x * res_y
        self.n = n
        self.nblocks = (res_x // 16) * (res_y // 16)
        self.screen = np.zeros(4 * n + 64, dtype=np.uint32)            # mem_screenbuffer  :85
        self.back = np.zeros(4 * n * 4 + 64, dtype=np.float32)         # mem_backbuffer    :82
        self.idbuf = np.zeros(n + 2 * self.nblocks + 64, dtype=np.uint32)  # mem_idbuffer  :268-270
        self.tex = np.zeros(n, dtype=np.uint32)                        # mem_screenbuffer_tex
        self.frame = -1
        self.idbuf_size = 0

    def tile(self):
        """src/raycast.h:363-364."""
        return (self.res_x // 8) * (self.frame & 7), (self.res_y // 4) * ((self.frame >> 3) & 3)

    def draw(self, pos, rot, stop_after=None):
        """One raycast_draw(); returns the wrapped camera position.  stop_after names a stage to stop at."""
        o, rx, ry, n, t = self.o, self.res_x, self.res_y, self.n, self.threads
        self.frame += 1
        frame = self.frame
        cam = camera_args(pos, rot, self.depth)
        v0 = cam["v0"]
        mt = t if self.timed else 1
        if frame < 2:                                                   # :150-154
            o.memset(self.screen, 0, HOLE, n * 4, threads=mt)
        o.memset(self.screen, 0, HOLE, n, threads=mt)                   # :157
        for i in range(2):                                              # :177-198
            o.raycast_proj(self.screen, self.back, rx, ry, frame, (i + 1) * n, v0, *cam["rows"], racy_threads=mt,
                           xbuf=self.xbuf, ybuf=self.ybuf)
        if self.quality_passes and (frame & 3) == 0:                    # :205-219 (if(0) in the reference)
            o.raycast_fillhole(self.screen, self.back, self.xbuf, self.ybuf, rx, ry, frame, t)
        if stop_after == "proj":
            return cam
        o.raycast_counthole(self.screen, self.idbuf, rx, ry, frame, t)  # :272-282
        o.raycast_sumids(self.screen, self.idbuf, rx, ry, frame)        # :287-296
        self.idbuf_size = int(self.idbuf[0])                            # :298
        o.raycast_writeids(self.screen, self.idbuf, rx, ry, frame, t)   # :305-315
        if stop_after == "ids":
            return cam
        if self.idbuf_size > 0:                                         # :332-359
            o.raycast_holes(self.screen, self.back, self.octree, self.idbuf, self.root, rx, ry, frame,
                            self.idbuf_size, v0, *cam["cols"], threads=t)
        add_x, add_y = self.tile()                                      # :361-387
        o.raycast_fine_2(self.screen, self.back, self.octree, self.root, rx, ry, frame, add_x, add_y,
                         v0, *cam["cols"], threads=t)
        if stop_after == "rays":
            return cam
        target = ((frame >> 4) % 2) + 1 if self.cache_rotation else 2   # :395
        o.memcpy(self.screen, target * n, self.screen, 0, n, threads=mt)    # :394-405
        back_u = self.back.view(np.uint32)
        o.memcpy(back_u, target * n * 4, back_u, 0, n * 4, threads=mt)
        if stop_after == "copy":
            return cam
        o.raycast_fillhole2(self.screen, rx, ry, frame, threads=t)      # :411-422 (snapshot semantics: race free)
        o.raycast_colorize(self.screen, self.tex, rx, ry, t)            # :429-437
        return cam


def flythrough_pose(f):
    """Scripted camera of SURVEY.md 8(d) config 2 (a definition of this repo, not of the reference).  rot.x is +0.6, not the
    survey's -0.6: with the reference's matrix conventions (ext/mathlib/_matrix44.h:529-541) a negative pitch points the camera
    at the sky -- 98.5 % of the primary rays leave the world at once -- while +0.6 looks down at the scene the way the
    reference's screenshot does.  bench.py uses the same path."""
    pos = (1.0 + f * 0.2357, 50.0, 1.0 + f * 0.2357)
    rot = (0.6 + 0.1 * math.sin(2.0 * math.pi * f / 128.0), 0.8 + 0.005 * f, 0.0)
    return pos, rot
