"""Screen bands across the GPUs of one box: host side of the ``svo_band_*`` C ABI (include/svo_b200.h).

The reference is single-device (src/ocl.h:89,127,140); this is the multi-GPU row of SURVEY.md 8(e).  The screen is cut
into stripes of ``stripe_rows`` rows, stripe s belongs to rank s % G; the octree is replicated; the reprojection's depth
test is resolved by 64-bit atomicMin straight into the owner's key buffer over NVLink, so every rank's rows are
bit-identical to the rows one GPU renders.

Three ways to drive it:
  * ``LocalBandSet``  one process, one band per listed device (the same device may be listed several times: "virtual"
                      bands on one GPU, which is how the single-GPU test box exercises the exchange code);
  * ``DistributedBand`` one process per GPU (torchrun): CUDA IPC handles are exchanged once through
                      ``torch.distributed.all_gather`` (any backend); no collective on the data path afterwards;
  * the pure layout helpers (no device): ``layout``, ``owned_rows``, ``owned_blocks``.
"""
import ctypes as C

import numpy as np

from . import ocl

HOLE = 0xFFFFFF00
SCREEN, BACK, IDBUF, TEX, HALO = 0, 1, 2, 3, 4
HANDLE_BYTES = 6 * 64

lib = ocl.lib
_vp, _i, _sz, _u32 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint32
_band_layout = ocl._sig("svo_band_layout", _i, _i, _i, _i, _i, _i, C.POINTER(_i * 4))
_band_owner = ocl._sig("svo_band_owner", _i, _i, _i, _i)
_band_create = ocl._sig("svo_band_create", _vp, _i, _i, _i, _i, _i, _vp, _u32)
_band_destroy = ocl._sig("svo_band_destroy", None, _vp)
_band_get_handles = ocl._sig("svo_band_get_handles", None, _vp, _vp)
_band_connect_ipc = ocl._sig("svo_band_connect_ipc", None, _vp, _vp)
_band_connect_local = ocl._sig("svo_band_connect_local", None, _vp, C.POINTER(_vp))
_band_frame = ocl._sig("svo_band_frame", None, _vp, C.POINTER(ocl.FrameParams))
_band_raycast = ocl._sig("svo_band_raycast", None, _vp, C.POINTER(ocl.FrameParams))
_band_sync = ocl._sig("svo_band_sync", None, _vp)
_band_join = ocl._sig("svo_band_join", None, _vp)
_band_read = ocl._sig("svo_band_read", _i, _vp, _i, _vp, _sz, _sz)
_band_write = ocl._sig("svo_band_write", _i, _vp, _i, _vp, _sz, _sz)
_band_ctx = ocl._sig("svo_band_ctx", _vp, _vp)
_band_last_slot = ocl._sig("svo_band_last_slot", _i, _vp)
_band_deferred_count = ocl._sig("svo_band_deferred_count", C.c_ulonglong, _vp)


# ---- layout (mirrors svo_band_layout; pure arithmetic) -----------------------------------------------------------
def effective_stripe_rows(nranks, res_y, stripe_rows=0):
    if stripe_rows > 0:
        return (stripe_rows + 15) // 16 * 16
    brows = (res_y + 15) // 16
    return -(-brows // nranks) * 16


def layout(rank, nranks, res_x, res_y, stripe_rows=0):
    """{'SR', 'rows', 'brows', 'stripes'} of one rank, computed by the library (no device needed)."""
    out = (_i * 4)()
    if _band_layout(rank, nranks, res_x, res_y, stripe_rows, C.byref(out)):
        raise ValueError("bad band layout arguments")
    return dict(SR=out[0], rows=out[1], brows=out[2], stripes=out[3])


def owned_rows(rank, nranks, res_y, stripe_rows=0):
    """Global row numbers a rank owns, in its local order."""
    SR = effective_stripe_rows(nranks, res_y, stripe_rows)
    y = np.arange(res_y)
    return y[(y // SR) % nranks == rank]


def owned_blocks(rank, nranks, res_x, res_y, stripe_rows=0):
    """Global 16x16 block indices (row-major over the whole screen) a rank owns, in its local order."""
    SR = effective_stripe_rows(nranks, res_y, stripe_rows)
    nbx, nby = res_x // 16, res_y // 16
    by = np.arange(nby)
    mine = by[((by * 16) // SR) % nranks == rank]
    return (mine[:, None] * nbx + np.arange(nbx)[None, :]).ravel()


# ---- one rank ------------------------------------------------------------------------------------------------------
class Band:
    """svo_band_t of the CURRENT context (ocl.ocl_init / Context.make_current first); the octree must live in it."""

    def __init__(self, rank, nranks, res_x, res_y, octree_mem, octree_root, stripe_rows=0):
        self.rank, self.nranks, self.res_x, self.res_y = rank, nranks, res_x, res_y
        self.lay = layout(rank, nranks, res_x, res_y, stripe_rows)
        self.stripe_rows = stripe_rows
        self.handle = _band_create(rank, nranks, res_x, res_y, stripe_rows, octree_mem.handle, octree_root)
        ocl._check()
        if not self.handle:
            raise RuntimeError("svo_band_create failed")

    def handles(self):
        buf = np.zeros(HANDLE_BYTES, dtype=np.uint8)
        _band_get_handles(self.handle, buf.ctypes.data)
        ocl._check()
        return buf

    def connect_ipc(self, all_handles):
        """all_handles: uint8 array [nranks, HANDLE_BYTES] (own row ignored)."""
        a = np.ascontiguousarray(all_handles, dtype=np.uint8).reshape(self.nranks, HANDLE_BYTES)
        _band_connect_ipc(self.handle, a.ctypes.data)
        ocl._check()

    def connect_local(self, bands):
        arr = (_vp * self.nranks)(*[b.handle for b in bands])
        _band_connect_local(self.handle, arr)
        ocl._check()

    def frame(self, params):
        _band_frame(self.handle, C.byref(params))
        ocl._check()

    def raycast(self, params):
        _band_raycast(self.handle, C.byref(params))
        ocl._check()

    def sync(self):
        _band_sync(self.handle)
        ocl._check()

    def join(self):
        """Order the stream behind the end-of-frame barrier of the last raycast() (svo_band_join)."""
        _band_join(self.handle)
        ocl._check()

    def last_slot(self):
        return int(_band_last_slot(self.handle))

    def ctx(self):
        return _band_ctx(self.handle)

    def deferred_count(self):
        """Frames of this rank whose scatter carried the previous frame's cache copy (svo_band_deferred_count)."""
        return int(_band_deferred_count(self.handle))

    def read(self, which, dtype, count, offset_bytes=0):
        out = np.empty(count, dtype=dtype)
        if _band_read(self.handle, which, out.ctypes.data, out.nbytes, offset_bytes):
            ocl._check()
            raise RuntimeError("svo_band_read failed")
        return out

    def write(self, which, array, offset_bytes=0):
        a = np.ascontiguousarray(array)
        if _band_write(self.handle, which, a.ctypes.data, a.nbytes, offset_bytes):
            ocl._check()
            raise RuntimeError("svo_band_write failed")

    def idbuf(self):
        """(counts, offsets, ids) of this rank's blocks; counts[0] is the block's count restored from the offsets."""
        nbl = (self.res_x // 16) * self.lay["brows"]
        if nbl == 0:
            return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32)
        head = self.read(IDBUF, np.uint32, 2 * nbl)
        total = int(head[0])
        ids = self.read(IDBUF, np.uint32, total, 2 * nbl * 4) if total else np.zeros(0, np.uint32)
        counts, offsets = head[:nbl].copy(), head[nbl:]
        counts[0] = (offsets[1] if nbl > 1 else total) - offsets[0]
        return counts, offsets, ids

    def destroy(self):
        if self.handle:
            _band_destroy(self.handle)
            self.handle = None


# ---- one process, several bands -------------------------------------------------------------------------------------
class LocalBandSet:
    """G bands driven by this process: ``devices[r]`` is the CUDA device of rank r (repeat a device for virtual bands)."""

    def __init__(self, devices, octree_words, octree_root, res_x, res_y, stripe_rows=0, depth=11):
        self.G, self.res_x, self.res_y, self.stripe_rows = len(devices), res_x, res_y, stripe_rows
        octree_words = np.ascontiguousarray(octree_words, dtype=np.uint32)
        self.ctxs, self.octrees, self.bands = [], [], []
        self.prev_ctx = ocl._svo_ctx_get_current()
        for r, dev in enumerate(devices):
            ctx = ocl._svo_ctx_create(dev)
            ocl._check()
            ocl._svo_ctx_set_current(ctx)
            ocl.set_octree_depth(depth)
            self.ctxs.append(ctx)
            self.octrees.append(ocl.ocl_malloc(octree_words.nbytes, octree_words))
            self.bands.append(Band(r, self.G, res_x, res_y, self.octrees[-1], octree_root, stripe_rows))
        for b in self.bands:
            b.connect_local(self.bands)

    def frame(self, params, sync=True):
        for b in self.bands:
            b.frame(params)
        if sync:
            self.sync()

    def raycast(self, params, sync=True):
        for b in self.bands:
            b.raycast(params)
        if sync:
            self.sync()

    def sync(self):
        for b in self.bands:
            b.sync()

    def assemble(self, slots=(0, 1, 2, 3)):
        """The global (screen[4N], back[16N]) image: every rank contributes the rows it owns."""
        n, rx = self.res_x * self.res_y, self.res_x
        screen = np.zeros(4 * n, dtype=np.uint32)
        back = np.zeros(16 * n, dtype=np.float32)
        for b in self.bands:
            s = b.read(SCREEN, np.uint32, 4 * n).reshape(4, self.res_y, rx)
            k = b.read(BACK, np.float32, 16 * n).reshape(4, self.res_y, rx * 4)
            rows = owned_rows(b.rank, self.G, self.res_y, self.stripe_rows)
            for slot in slots:
                screen.reshape(4, self.res_y, rx)[slot, rows] = s[slot, rows]
                back.reshape(4, self.res_y, rx * 4)[slot, rows] = k[slot, rows]
        return screen, back

    def frame_image(self):
        """The writer's (rank 0) colorized frame."""
        return self.bands[0].read(TEX, np.uint32, self.res_x * self.res_y)

    def merged_ids(self):
        """(counts[B], ids) in the 1-GPU order: per global block the owner's count, and the concatenated id lists."""
        nb = (self.res_x // 16) * (self.res_y // 16)
        counts = np.zeros(nb, dtype=np.uint32)
        per_block = [None] * nb
        for b in self.bands:
            blocks = owned_blocks(b.rank, self.G, self.res_x, self.res_y, self.stripe_rows)
            c, o, ids = b.idbuf()
            assert len(c) == len(blocks)
            for lb, gb in enumerate(blocks):
                counts[gb] = c[lb]
                per_block[gb] = ids[o[lb]:o[lb] + c[lb]]
        ids = np.concatenate([p for p in per_block if p is not None and len(p)]) if counts.sum() else np.zeros(0, np.uint32)
        return counts, ids

    def close(self):
        for b in self.bands:
            b.destroy()
        for ctx, oc in zip(self.ctxs, self.octrees):
            ocl._svo_ctx_set_current(ctx)
            oc.free()
            ocl._svo_ctx_destroy(ctx)
        ocl._svo_ctx_set_current(self.prev_ctx)


# ---- one process per GPU ------------------------------------------------------------------------------------------------
def exchange_handles(my_handles, dist=None):
    """all_gather of the IPC handle blobs over torch.distributed (gloo or nccl); returns uint8 [world, HANDLE_BYTES]."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(np.ascontiguousarray(my_handles, dtype=np.uint8)).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return np.stack([t.cpu().numpy() for t in out])


class DistributedBand:
    """This process' band of a torchrun job: rank r drives GPU ``device`` and owns the stripes r, r+G, ..."""

    def __init__(self, octree_words, octree_root, res_x, res_y, device, stripe_rows=0, depth=11, dist=None):
        """dist: an initialised torch.distributed module, or None for a single-process (world 1) run."""
        self.dist = dist
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        ocl.ocl_init(device)
        ocl.set_octree_depth(depth)
        octree_words = np.ascontiguousarray(octree_words, dtype=np.uint32)
        self.octree = ocl.ocl_malloc(octree_words.nbytes, octree_words)
        self.band = Band(self.rank, self.world, res_x, res_y, self.octree, octree_root, stripe_rows)
        if self.world > 1:
            self.band.connect_ipc(exchange_handles(self.band.handles(), dist))
            dist.barrier()

    def close(self):
        if self.world > 1:
            self.dist.barrier()
        self.band.destroy()
        self.octree.free()
        ocl.ocl_exit()
