"""Synthetic voxel scenes for the parity tests (definitions of this repo; the reference ships none).

Every generator returns (x, y, z, rgba) uint32 arrays in INSERTION ORDER (the order matters:
an interior node's colour is the colour of the first voxel inserted below it,
src/octree/octree.h:48,80).  The low colour byte is always non-zero and has its two low bits
= 01 so a voxel can never alias the hole marker 0xffffff00.
"""
import numpy as np


def _colour(x, y, z, seed=0):
    h = (x.astype(np.uint64) * 73856093) ^ (y.astype(np.uint64) * 19349663) ^ (z.astype(np.uint64) * 83492791)
    h = (h + np.uint64(seed) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    h = (h ^ (h >> np.uint64(13))) * np.uint64(0x5BD1E995) & np.uint64(0xFFFFFFFF)
    lo = (1 + ((h >> np.uint64(8)) & np.uint64(0xFC))).astype(np.uint32)
    g = ((h >> np.uint64(16)) & np.uint64(0xF8)).astype(np.uint32)
    b = ((h >> np.uint64(24)) & np.uint64(0xF8)).astype(np.uint32)
    return lo | (g << 8) | (b << 16)


def _pack(x, y, z, seed=0):
    x, y, z = (np.asarray(a, dtype=np.uint32).ravel() for a in (x, y, z))
    return x, y, z, _colour(x, y, z, seed)


def value_noise(nx, nz, octaves=6, seed=0x5EED, base=4):
    """fBm of bilinear value noise on an nx*nz grid in [0,1) (lacunarity 2, gain 0.5)."""
    rng = np.random.RandomState(seed)
    out = np.zeros((nx, nz), dtype=np.float64)
    amp, tot, cells = 1.0, 0.0, base
    for _ in range(octaves):
        g = rng.rand(cells + 1, cells + 1)
        u = np.linspace(0, cells, nx, endpoint=False)
        v = np.linspace(0, cells, nz, endpoint=False)
        iu, iv = u.astype(int), v.astype(int)
        fu, fv = (u - iu)[:, None], (v - iv)[None, :]
        fu, fv = fu * fu * (3 - 2 * fu), fv * fv * (3 - 2 * fv)
        a = g[iu][:, iv]; b = g[iu + 1][:, iv]; c = g[iu][:, iv + 1]; d = g[iu + 1][:, iv + 1]
        out += amp * ((a * (1 - fu) + b * fu) * (1 - fv) + (c * (1 - fu) + d * fu) * fv)
        tot += amp; amp *= 0.5; cells *= 2
    return out / tot


def terrain(n=512, x0=0, z0=0, height=160, base=60, thickness=3, seed=0x5EED, order="zx"):
    """n*n heightfield shell of `thickness` voxels; insertion order z outer, x, then y descending
    (the order RLE4::load produces: slice, x, y1 ascending = y descending, Rle4.cpp:95-96,116-123)."""
    h = (base + height * value_noise(n, n, seed=seed)).astype(np.int64)
    xs, zs = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    if order == "zx":
        xs, zs, h = xs.T, zs.T, h.T
    X = np.repeat(xs.ravel(), thickness) + x0
    Z = np.repeat(zs.ravel(), thickness) + z0
    Y = (np.repeat(h.ravel(), thickness) - np.tile(np.arange(thickness), n * n))
    return _pack(X, Y, Z, seed)


def blob(cx, cy, cz, radius, amp=0.25, seed=7):
    """Closed surface shell: voxels with |p-c| within [r(dir), r(dir)+1.5) where r is a noisy radius."""
    r = int(radius * (1 + amp)) + 2
    ax = np.arange(-r, r + 1)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    d = np.sqrt(X * X + Y * Y + Z * Z) + 1e-9
    bump = (np.sin(X / d * 5.0 + seed) * np.cos(Y / d * 4.0 - seed) + np.sin(Z / d * 7.0)) * 0.5
    rr = radius * (1 + amp * bump)
    m = (d >= rr) & (d < rr + 1.5)
    # insertion order: z outer, x, y descending
    idx = np.argwhere(m)
    key = np.lexsort((-idx[:, 1], idx[:, 0], idx[:, 2]))
    idx = idx[key]
    return _pack(idx[:, 0] + cx, idx[:, 1] + cy, idx[:, 2] + cz, seed)


def concat(*scenes):
    return tuple(np.concatenate([s[k] for s in scenes]) for k in range(4))


def single_voxel(x=100, y=100, z=100):
    return _pack([x], [y], [z])


def cube(x0, y0, z0, n):
    ax = np.arange(n)
    X, Y, Z = np.meshgrid(ax + x0, ax + y0, ax + z0, indexing="ij")
    return _pack(X, Y, Z)


def random_cloud(n, lo=0, hi=2048, seed=1):
    rng = np.random.RandomState(seed)
    p = rng.randint(lo, hi, size=(n, 3))
    return _pack(p[:, 0], p[:, 1], p[:, 2], seed)


def duplicates():
    """The same voxel inserted twice with different colours + neighbours (leaf takes the later colour,
    interior nodes keep the first, src/octree/octree.h:67-71,75-82)."""
    x = np.array([5, 5, 4, 5, 1029], dtype=np.uint32)
    y = np.array([9, 9, 9, 8, 77], dtype=np.uint32)
    z = np.array([3, 3, 3, 3, 2000], dtype=np.uint32)
    c = np.array([0x000011, 0x0000F5, 0x000021, 0x000031, 0x0000FD], dtype=np.uint32)
    return x, y, z, c


def small_world(seed=0x5EED):
    """Terrain patch + blob + edge cases: the standard small scene of the per-kernel parity tests."""
    return concat(terrain(384, 0, 0, height=120, base=40, seed=seed),
                  blob(200, 260, 220, 48, seed=seed & 0xFF),
                  cube(60, 200, 60, 9),              # full 8x8x8-aligned cube + overhang: byte-packed k=8 records
                  single_voxel(2047, 2047, 2047),    # far corner: a block with a single voxel (k=1)
                  single_voxel(63, 300, 64),         # voxel next to a block boundary
                  random_cloud(500, 0, 2048, seed=3))
