"""GPU parity tests (run with -m gpu on the B200 box): every CUDA kernel is driven through the C ABI
(libsvo_b200.so, the ocl_* mirror) with the argument lists of the reference's call sites and compared with
the CPU oracle on identical input buffers.

Bars: colour / depth-key / index words, hole index buffer, tile schedule and colorized output are BIT-EXACT;
hit positions are compared bit-exact as well (the kernels are compiled without FMA contraction) and, as the
stated tolerance of BASELINE.json, to 1e-4 relative.
"""
import ctypes as C
import os

import numpy as np
import pytest

import scenes
from oracle import frame as ofr

pytestmark = pytest.mark.gpu
HOLE = 0xFFFFFF00
RTOL = 1e-4          # north_star: hit positions and depth agree within 1e-4 relative error


@pytest.fixture(scope="module")
def svo():
    from __graft_entry__ import load_package
    m = load_package()
    m.Device.errors_return()
    m.ocl_init(0)
    yield m
    m.ocl_exit()


@pytest.fixture(scope="module")
def world(orc):
    octree, root = orc.build_octree(*scenes.small_world())
    return octree, root


def launch(svo, name, gx, gy, lx, ly, args):
    k = svo.ocl_get_kernel(name)
    svo.ocl_begin(k, gx, gy, lx, ly)
    for a in args:
        svo.ocl_param(a)
    svo.ocl_end()
    svo.ocl_end_all_kernels()


def i32(v):
    return C.c_int(int(v))


def random_screen(rng, n, hole_frac, blocks=True):
    img = rng.randint(0, 2 ** 32 - 1, size=n, dtype=np.uint64).astype(np.uint32)
    img[rng.rand(n) < hole_frac] = HOLE
    return img


# ------------------------------------------------------------------------------------------------------
def test_memset_memcpy(svo, orc):
    n = 256 * 40
    a = np.arange(3 * n, dtype=np.uint32)
    m = svo.ocl_malloc(a.nbytes, a)
    svo.ocl_memset(m, 256 * 4, 0xFFFFFF00, 256 * 8 * 4)
    svo.ocl_memcpy(m, 2 * n * 4, m, 512 * 4, 256 * 16 * 4)
    svo.ocl_end_all_kernels()
    orc.memset(a, 256, 0xFFFFFF00, 256 * 8)
    orc.memcpy(a, 2 * n, a, 512, 256 * 16)
    assert np.array_equal(m.to_numpy(), a)
    m.free()


@pytest.mark.parametrize("res", [(64, 48), (320, 192), (1920, 1024), (200, 120)])
@pytest.mark.parametrize("hole_frac", [0.0, 0.3, 0.9, 1.0])
def test_hole_gather(svo, orc, res, hole_frac):
    """raycast_counthole / raycast_sumids / raycast_writeids: the whole id buffer bit-exact."""
    rx, ry = res
    n, nb = rx * ry, (rx // 16) * (ry // 16)
    rng = np.random.RandomState(rx + int(hole_frac * 10))
    img = random_screen(rng, n, hole_frac)
    # holes come in 2x2 cells in practice: clear random cells entirely so that cells do qualify
    cells = rng.rand(ry // 2, rx // 2) < hole_frac * 0.5
    full = np.kron(cells, np.ones((2, 2), dtype=bool))
    img2 = img.reshape(ry, rx).copy()
    img2[:full.shape[0], :full.shape[1]][full] = HOLE
    img = img2.ravel()
    idb = np.zeros(n + 2 * nb + 64, dtype=np.uint32)
    orc.raycast_counthole(img, idb, rx, ry)
    orc.raycast_sumids(img, idb, rx, ry)
    total = int(idb[0])
    orc.raycast_writeids(img, idb, rx, ry)
    ms, mi = svo.ocl_malloc(img.nbytes, img), svo.ocl_malloc(idb.nbytes, np.zeros_like(idb))
    for name in ("raycast_counthole", "raycast_sumids", "raycast_writeids"):
        g = (1, 1, 1, 1) if name == "raycast_sumids" else (rx // 16, ry // 16, 16, 16)
        launch(svo, name, *g, [ms, None, mi, i32(rx), i32(ry), i32(0)])
    got = mi.to_numpy()
    assert int(got[0]) == total
    assert np.array_equal(got[:2 * nb + total], idb[:2 * nb + total])
    ms.free(); mi.free()


@pytest.mark.parametrize("res", [(80, 48), (320, 192), (1920, 1024)])
@pytest.mark.parametrize("hole_frac", [0.02, 0.3, 0.8, 0.97])
def test_fillhole2(svo, orc, res, hole_frac):
    rx, ry = res
    n = rx * ry
    rng = np.random.RandomState(7)
    img = random_screen(rng, 4 * n, hole_frac)
    exp = img.copy()
    orc.raycast_fillhole2(exp, rx, ry)
    m = svo.ocl_malloc(img.nbytes, img)
    launch(svo, "raycast_fillhole2", rx, ry, 16, 16, [m, None, i32(rx), i32(ry), i32(0)])
    assert np.array_equal(m.to_numpy(), exp)
    m.free()


def test_colorize(svo, orc):
    rx, ry = 512, 256
    rng = np.random.RandomState(3)
    img = random_screen(rng, rx * ry, 0.1)
    img[:1024] = np.arange(1024)          # every (palette, intensity) combination
    exp = np.zeros(rx * ry, dtype=np.uint32)
    orc.raycast_colorize(img, exp, rx, ry)
    ms, mt = svo.ocl_malloc(img.nbytes, img), svo.ocl_malloc(img.nbytes)
    launch(svo, "raycast_colorize", rx, ry, 16, 16, [ms, mt, i32(rx), i32(ry)])
    assert np.array_equal(mt.to_numpy(), exp)
    ms.free(); mt.free()


@pytest.mark.parametrize("pose", [((10, 22, 9), (0.4, 0.7, 0.0)), ((1, 50, 1), (0.6, 0.8, 0.0)),
                                  ((30, 12, 30), (0.2, 3.9, 0.05)), ((20, 3, 20), (1.2, -0.7, 0.0)),
                                  ((100, 200, 100), (-0.5, 1.0, 0.0))])
def test_raycast_fine_2_full_screen(svo, orc, world, pose):
    """Full-screen primary rays: colour words bit-exact, positions within RTOL (and reported if not bit-exact)."""
    octree, root = world
    rx, ry = 320, 192
    n = rx * ry
    cam = ofr.camera_args(*pose)
    screen = np.full(4 * n, HOLE, dtype=np.uint32)
    back = np.zeros(16 * n, dtype=np.float32)
    orc.raycast_fine_2(screen, back, octree, root, rx, ry, 0, 0, 0, cam["v0"], *cam["cols"], gx=rx, gy=ry, threads=4)
    mo = svo.ocl_malloc(octree.nbytes, octree)
    ms = svo.ocl_malloc(4 * n * 4, np.full(4 * n, HOLE, dtype=np.uint32))
    mb = svo.ocl_malloc(16 * n * 4, np.zeros(16 * n, dtype=np.float32))
    dead = (0, 0, 0, 0)
    launch(svo, "raycast_fine_2", rx, ry, 16, 16,
           [ms, mb, mo, C.c_uint32(root), i32(rx), i32(ry), i32(0), i32(0), i32(0), dead, dead, dead, dead,
            cam["v0"], *cam["cols"], C.c_float(1.0), C.c_float(1.0)])
    got_s, got_b = ms.to_numpy(), mb.to_numpy(np.float32)
    same = got_s == screen
    assert same.mean() >= 0.9999, f"only {same.mean():.6f} of pixels agree on the hit word"
    assert same.all(), f"{(~same).sum()} hit words differ"
    assert np.allclose(got_b, back, rtol=RTOL, atol=1e-4)
    assert np.array_equal(got_b.view(np.uint32), back.view(np.uint32)), "positions not bit-exact"
    for m in (mo, ms, mb):
        m.free()


def test_raycast_proj(svo, orc, world):
    """Reprojection of a raycast frame into a moved camera: key words, payload and source invalidation."""
    octree, root = world
    rx, ry = 320, 192
    n = rx * ry
    cam0 = ofr.camera_args((10, 22, 9), (0.4, 0.7, 0.0))
    screen = np.full(4 * n, HOLE, dtype=np.uint32)
    back = np.zeros(16 * n, dtype=np.float32)
    # cache buffer 2 <- full raycast from camera 0 (oracle), buffer 1 <- a shifted second view to exercise both launches
    tmp_s, tmp_b = screen.copy(), back.copy()
    orc.raycast_fine_2(tmp_s, tmp_b, octree, root, rx, ry, 0, 0, 0, cam0["v0"], *cam0["cols"], gx=rx, gy=ry, threads=4)
    screen[2 * n:3 * n], back[8 * n:12 * n] = tmp_s[:n], tmp_b[:4 * n]
    cam1 = ofr.camera_args((10.6, 22.2, 9.5), (0.42, 0.75, 0.0))
    orc.raycast_fine_2(tmp_s, tmp_b, octree, root, rx, ry, 0, 0, 0, cam1["v0"], *cam1["cols"], gx=rx, gy=ry, threads=4)
    screen[n:2 * n], back[4 * n:8 * n] = tmp_s[:n], tmp_b[:4 * n]
    cam2 = ofr.camera_args((11.0, 22.0, 10.0), (0.45, 0.8, 0.01))
    ms, mb = svo.ocl_malloc(screen.nbytes, screen), svo.ocl_malloc(back.nbytes, back)
    for i in range(2):
        orc.raycast_proj(screen, back, rx, ry, 5, (i + 1) * n, cam2["v0"], *cam2["rows"])
        launch(svo, "raycast_proj", rx, ry, 16, 16,
               [ms, mb, None, None, None, i32(rx), i32(ry), i32(5), i32((i + 1) * n), cam2["v0"], *cam2["rows"]])
        assert np.array_equal(ms.to_numpy(), screen), f"launch {i}: colour/depth words differ"
        assert np.array_equal(mb.to_numpy(np.uint32), back.view(np.uint32)), f"launch {i}: payload differs"
    assert (screen[:n] != HOLE).mean() > 0.5
    ms.free(); mb.free()


@pytest.mark.parametrize("mode,res,nframes", [("reference", (320, 192), 40), ("fused", (320, 192), 40),
                                              ("fused", (200, 120), 8), ("reference", (200, 120), 4),
                                              ("fused", (1920, 1024), 5)])
def test_frame_sequence(svo, orc, world, mode, res, nframes):
    """Frames of the full pipeline (40 covers a whole 32-tile refresh period): every buffer after every frame.
    200x120 is not a multiple of the 16x16 hole-block size (right/bottom strips); 1920x1024 is the bench size."""
    octree, root = world
    rx, ry = res
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4)
    rc = svo.raycast
    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode=mode)
    try:
        for f in range(nframes):
            pos, rot = (10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0)
            O.draw(pos, rot)
            rc.set_camera(pos, rot)
            rc.raycast_draw(rx, ry)
            assert rc.tile_origin(f, rx, ry) == O.tile()
            screen, back, idb = rc.read_buffers(rx, ry)
            assert rc.idbuf_size() == O.idbuf_size, f"frame {f}"
            assert np.array_equal(idb[:2 * O.nblocks + O.idbuf_size], O.idbuf[:2 * O.nblocks + O.idbuf_size]), f"frame {f} ids"
            assert np.array_equal(screen, O.screen[:4 * n]), f"frame {f} colour"
            assert np.array_equal(back.view(np.uint32), O.back[:16 * n].view(np.uint32)), f"frame {f} xyz"
            assert np.array_equal(rc.read_frame(rx, ry).ravel(), O.tex), f"frame {f} tex"
    finally:
        rc.raycast_exit()
        svo.ocl_init(0)


@pytest.mark.parametrize("res,nframes", [((320, 192), 36), ((200, 120), 6), ((1920, 1024), 4)])
def test_frame_sequence_pingpong(svo, orc, world, res, nframes):
    """SVO_FRAME_PINGPONG: no cache copy.  The id buffer and the colorized image are the reference's; the slot rendered
    into holds what the reference's cache buffer 2 holds (colour words and xyzw)."""
    octree, root = world
    rx, ry = res
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4)
    rc = svo.raycast
    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode="pingpong")
    try:
        for f in range(nframes):
            pos, rot = (10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0)
            O.draw(pos, rot)
            rc.set_camera(pos, rot)
            rc.raycast_draw(rx, ry)
            slot = rc.last_slot()
            assert slot == (0 if f % 2 == 0 else 2)
            screen, back, idb = rc.read_buffers(rx, ry)
            assert rc.idbuf_size() == O.idbuf_size, f"frame {f}"
            assert np.array_equal(idb[:2 * O.nblocks + O.idbuf_size], O.idbuf[:2 * O.nblocks + O.idbuf_size]), f"frame {f} ids"
            assert np.array_equal(screen[slot * n:(slot + 1) * n], O.screen[2 * n:3 * n]), f"frame {f} colour"
            got = back[slot * 4 * n:(slot + 1) * 4 * n].view(np.uint32).reshape(n, 4)
            exp = O.back[8 * n:12 * n].view(np.uint32).reshape(n, 4)
            # xyz of a pixel that is still a hole is whatever an older frame left there (never read: the reprojection
            # skips holes, kernel.cl:499); compare the positions of all non-hole pixels
            live = O.screen[2 * n:3 * n] != HOLE
            assert np.array_equal(got[live, :3], exp[live, :3]), f"frame {f} xyz"
            # w (camera z) is written by the reprojection only; for ray-filled pixels the reference keeps whatever an
            # older frame left in buffer 0 (dead data no enabled kernel reads) and the alternating slots keep another
            # stale value.  Compare w where this frame's reprojection wrote it.
            traced = np.zeros(n, dtype=bool)
            ids = O.idbuf[2 * O.nblocks:2 * O.nblocks + O.idbuf_size]
            traced[(ids & 0xffff).astype(np.int64) + (ids >> 16).astype(np.int64) * rx] = True
            ax, ay = O.tile()
            tmask = np.zeros((ry, rx), dtype=bool)
            tmask[ay:ay + -(-(ry // 4) // 16) * 16, ax:ax + -(-(rx // 8) // 16) * 16] = True
            proj = live & ~traced & ~tmask.ravel()
            assert np.array_equal(got[proj, 3], exp[proj, 3]), f"frame {f} w"
            assert np.array_equal(rc.read_frame(rx, ry).ravel(), O.tex), f"frame {f} tex"
    finally:
        rc.raycast_exit()
        svo.ocl_init(0)


@pytest.mark.parametrize("res,nframes,observe", [((320, 192), 3, ()), ((320, 192), 7, ()), ((320, 192), 40, ()),
                                                 ((320, 192), 12, (4, 5, 9)), ((200, 120), 9, ()),
                                                 ((1920, 1024), 8, ()), ((1920, 1024), 7, (3,))])
def test_frame_sequence_back_to_back(svo, orc, world, res, nframes, observe):
    """The schedule the bench runs: fused frames enqueued back to back with no observation in between, so every frame from
    the third on reads the previous frame out of buffer 0 and carries its cache copy in the reprojection pass, while the
    colorize pass and the gap filter run on the side stream (svo_frame_deferred_count proves it).  The colorized image of EVERY frame comes back through the
    headless present path (no flush) and is compared with the oracle; after the last frame every buffer is compared.
    `observe` lists frames after which the host reads the buffers (flushes the pipeline): the next frame must then fall
    back to the cache copy as its source."""
    octree, root = world
    rx, ry = res
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4)
    rc, ocl = svo.raycast, svo.ocl
    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode="fused")
    host = [ocl.host_alloc(n * 4), ocl.host_alloc(n * 4)]
    try:
        tex_ref, params, obs_ref = [], [], {}
        for f in range(nframes):
            pos, rot = (10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0)
            O.draw(pos, rot)
            tex_ref.append(O.tex.copy())
            if f in observe:
                obs_ref[f] = O.screen[:n].copy()
            rc.set_camera(pos, rot)
            params.append(rc.prepare_params(rx, ry, f))
        base, base_early = ocl.frame_deferred_count(), ocl.frame_early_count()
        expect_deferred = 0
        for f in range(nframes):
            if f >= 2:                                   # the caller owns host[f & 1] again: frame f-2 has landed
                ocl.present_wait(f & 1)
                got = np.frombuffer(host[f & 1], dtype=np.uint32).copy()
                assert np.array_equal(got, tex_ref[f - 2]), f"frame {f - 2} tex"
            k = rc.draw_present(params[f], host)
            assert k == (f & 1)
            if f >= 2 and (f - 1) not in observe:
                expect_deferred += 1
            if f in observe:
                screen, back, idb = rc.read_buffers(rx, ry)      # flushes: buffer 0 gets the filter's in-place words
                assert np.array_equal(screen[:n], obs_ref[f]), f"frame {f} colour (observed)"
        for f in range(max(0, nframes - 2), nframes):
            ocl.present_wait(f & 1)
            got = np.frombuffer(host[f & 1], dtype=np.uint32).copy()
            assert np.array_equal(got, tex_ref[f]), f"frame {f} tex"
        assert ocl.frame_deferred_count() - base == expect_deferred
        # ... and every one of them ran its reprojection as early pass (beside the previous frame's hole rays) + list pass
        assert ocl.frame_early_count() - base_early == expect_deferred
        screen, back, idb = rc.read_buffers(rx, ry)
        assert rc.idbuf_size() == O.idbuf_size
        assert np.array_equal(idb[:2 * O.nblocks + O.idbuf_size], O.idbuf[:2 * O.nblocks + O.idbuf_size]), "ids"
        assert np.array_equal(screen, O.screen[:4 * n]), "colour"
        assert np.array_equal(back.view(np.uint32), O.back[:16 * n].view(np.uint32)), "xyz"
    finally:
        for h in host:
            ocl.host_free(h)
        rc.raycast_exit()
        svo.ocl_init(0)


def test_resolution_change_back_to_back(svo, orc, world):
    """A host that changes the resolution inside one context (the reference's window resize): 2 frames at 320x192, then the
    frame counter restarts and 7 frames at 200x120 follow, all back to back.  200x120 has right / bottom strips outside the
    whole 16x16 blocks, which the early reprojection pass must not skip: its cell mask is laid out per resolution and the
    strip cells are never written by the id pass, so bytes of the 320x192 layout must not survive the change -- and after
    frames 0 and 1 (every cell a hole cell) that layout is all ones."""
    octree, root = world
    rc, ocl = svo.raycast, svo.ocl
    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=320, max_h=192, mode="fused")
    host = [ocl.host_alloc(320 * 192 * 4), ocl.host_alloc(320 * 192 * 4)]
    try:
        for (rx, ry), nframes in (((320, 192), 2), ((200, 120), 7)):
            n = rx * ry
            O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4)
            params = []
            for f in range(nframes):
                pos, rot = (10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0)
                O.draw(pos, rot)
                rc.set_camera(pos, rot)
                params.append(rc.prepare_params(rx, ry, f))
            base = ocl.frame_early_count()
            for f in range(nframes):
                if f >= 2:
                    ocl.present_wait(f & 1)
                rc.draw_present(params[f], host)
            for f in range(max(0, nframes - 2), nframes):
                ocl.present_wait(f & 1)
            assert ocl.frame_early_count() - base == nframes - 2
            got = np.frombuffer(host[(nframes - 1) & 1], dtype=np.uint32)[:n].copy()
            assert np.array_equal(got, O.tex), f"{rx}x{ry} tex"
            screen, back, idb = rc.read_buffers(rx, ry)
            assert rc.idbuf_size() == O.idbuf_size
            assert np.array_equal(idb[:2 * O.nblocks + O.idbuf_size], O.idbuf[:2 * O.nblocks + O.idbuf_size]), f"{rx}x{ry} ids"
            assert np.array_equal(screen, O.screen[:4 * n]), f"{rx}x{ry} colour"
            # positions: the words a frame defines -- xyz of the non-hole pixels of the cache copy (buffer 2, the pre-filter
            # frame; in buffer 0 the gap filter has coloured pixels that have no position).  The rest of the coordinate buffers
            # still holds the other resolution's data here and zeros in the freshly started oracle.
            got_b, exp_b = back.view(np.uint32).reshape(4, n, 4), O.back[:16 * n].view(np.uint32).reshape(4, n, 4)
            live = screen[2 * n:3 * n] != HOLE
            assert np.array_equal(got_b[2][live][:, :3], exp_b[2][live][:, :3]), f"{rx}x{ry} xyz of the cache copy"
    finally:
        for h in host:
            ocl.host_free(h)
        rc.raycast_exit()
        svo.ocl_init(0)


@pytest.mark.parametrize("res", [(320, 192), (201, 121)])
def test_present_rgb24(svo, orc, world, res):
    """Headless present as R,G,B bytes, both ways: (a) SVO_FRAME_TEX_RGB24 -- the kernels that produce the pixels store the
    packed bytes themselves and the read-back is a plain copy (draw_present(rgb24=True), what bench.py's e2e leg runs);
    (b) svo_present_rgb24_async -- a 0x00RRGGBB word image packed by k_pack_rgb24 on the copy stream.  Both must equal the
    word image of the same frame byte for byte, for observed and for back-to-back frames; 201x121 has an odd row length
    (scalar store paths) and a pixel count that is not a multiple of four (tail path of the pack kernel)."""
    octree, root = world
    rx, ry = res
    n = rx * ry
    rc, ocl = svo.raycast, svo.ocl
    nframes = 7
    cams = [((10 + 0.25 * f, 22, 9 + 0.2 * f), (0.4, 0.7 + 0.01 * f, 0.0)) for f in range(nframes)]

    def to_rgb(tex):
        return np.stack([(tex >> 16) & 255, (tex >> 8) & 255, tex & 255], axis=1).astype(np.uint8)

    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode="fused")
    host = [ocl.host_alloc(n * 4), ocl.host_alloc(n * 4)]
    try:
        words = []                                       # word images, frames observed one by one; (b) on each of them
        for f, (pos, rot) in enumerate(cams):
            rc.set_camera(pos, rot)
            k = rc.draw_present(rc.prepare_params(rx, ry, f), host)
            ocl.present_wait(k)
            tex = np.frombuffer(host[k], dtype=np.uint32)[:n].copy()
            assert tex.max() < (1 << 24)
            words.append(tex)
            src = svo.raycast.S.mem_screenbuffer_tex2 if k else svo.raycast.S.mem_screenbuffer_tex
            ocl.present_rgb24_async(host[k], src, n, k)
            ocl.present_wait(k)
            assert np.array_equal(np.frombuffer(host[k], dtype=np.uint8)[:3 * n].reshape(n, 3), to_rgb(tex)), f"pack, frame {f}"
    finally:
        rc.raycast_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode="fused")
    try:
        for f, (pos, rot) in enumerate(cams):            # (a), back to back: frames 2.. run the early / lazy schedule
            rc.set_camera(pos, rot)
            if f >= 2:
                ocl.present_wait(f & 1)
                got = np.frombuffer(host[f & 1], dtype=np.uint8)[:3 * n].reshape(n, 3).copy()
                assert np.array_equal(got, to_rgb(words[f - 2])), f"rgb24 producers, frame {f - 2}"
            assert rc.draw_present(rc.prepare_params(rx, ry, f), host, rgb24=True) == (f & 1)
        for f in range(nframes - 2, nframes):
            ocl.present_wait(f & 1)
            got = np.frombuffer(host[f & 1], dtype=np.uint8)[:3 * n].reshape(n, 3).copy()
            assert np.array_equal(got, to_rgb(words[f])), f"rgb24 producers, frame {f}"
    finally:
        for h in host:
            ocl.host_free(h)
        rc.raycast_exit()
        svo.ocl_init(0)


@pytest.mark.parametrize("mode", ["fused", "pingpong"])
def test_schedule_switches(svo, mode):
    """The A/B switches of the fused frame's schedule only move work between streams and launches:
    every combination must leave bit-identical buffers, ids and images.  The default schedule is the one the other tests
    pin against the oracle; here each switch (svo_debug_set) runs the same 12 back-to-back frames in a process of its
    own and its digest is compared with the default's."""
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_schedule_worker.py")

    def run(extra):
        env = {k: v for k, v in os.environ.items() if not k.startswith("SVO_")}
        env.update(extra)
        out = subprocess.run([sys.executable, worker, mode, "320", "192", "12"], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        line = [l for l in out.stdout.splitlines() if l.startswith("DIGEST")][-1].split()
        return line[1], int(line[2])

    ref, deferred = run({})
    assert deferred == (10 if mode == "fused" else 0)          # frames 2..11 carried the previous frame's cache copy
    for switches in ("no_split_resolve", "no_tile_staging", "no_lazy_copy", "no_overlap", "no_lazy_copy,no_split_resolve", "frame_l2_pin",
                     "main_lo", "holes_smax=1", "holes_smax=32", "no_early_scatter", "no_early_scatter,holes_smax=32", "no_tile_early"):
        got, _ = run({"SVO_TEST_SWITCHES": switches})
        assert got == ref, f"{switches} changes the result"


@pytest.mark.parametrize("mode,observe", [("reference", True), ("fused", True), ("fused", False)], ids=["reference", "fused", "fused-back-to-back"])
def test_cache_rotation(svo, orc, world, mode, observe):
    """The copy target the reference keeps in a comment (`((frame>>4)%2)+1`, src/raycast.h:395): cache
    buffers 1 and 2 alternate every 16 frames, so both reprojection launches see real frames, depth ties between them occur
    (buffer 1 wins: the earlier launch) and the source-side store of raycast_proj (a cached pixel that leaves the view
    becomes a hole, kernel.cl:559-562) is visible in later frames.  Launch-by-launch and fused (SVO_FRAME_CACHE_ROTATION)
    against the oracle over 40 frames (two target switches), every buffer; the fused frame also back to back, where the
    next frame's reprojection carries the copy into the rotating target."""
    octree, root = world
    rx, ry = 320, 192
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4, cache_rotation=True)
    rc = svo.raycast
    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode=mode, cache_rotation=True)
    try:
        d0 = svo.ocl.frame_deferred_count()
        for f in range(40):
            pos, rot = (10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0)
            O.draw(pos, rot)
            rc.set_camera(pos, rot)
            if observe:
                rc.raycast_draw(rx, ry)
            else:
                rc.draw_prepared(rc.prepare_params(rx, ry, f), sync=False)
                if f != 39:
                    continue
                assert svo.ocl.frame_deferred_count() - d0 == 38            # every frame from 2 on carried its predecessor's copy
            screen, back, idb = rc.read_buffers(rx, ry)
            assert rc.idbuf_size() == O.idbuf_size, f"frame {f}"
            assert np.array_equal(idb[:2 * O.nblocks + O.idbuf_size], O.idbuf[:2 * O.nblocks + O.idbuf_size]), f"frame {f} ids"
            assert np.array_equal(screen, O.screen[:4 * n]), f"frame {f} colour"
            assert np.array_equal(back.view(np.uint32), O.back[:16 * n].view(np.uint32)), f"frame {f} xyz"
            assert np.array_equal(rc.read_frame(rx, ry).ravel(), O.tex), f"frame {f} tex"
        assert np.any(O.screen[n:2 * n] != HOLE) and np.any(O.screen[2 * n:3 * n] != HOLE)      # both caches are live
        with pytest.raises(ValueError):
            rc.raycast_exit()
            rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode="pingpong", cache_rotation=True)
    finally:
        rc.raycast_exit()
        svo.ocl_init(0)


def test_quality_passes_reference_mode(svo, orc, world):
    """the depth-discontinuity hole punch the reference keeps behind if(0) (raycast_fillhole every 4th frame,
    src/raycast.h:205-219) made usable: raycast_proj writes the winners' motion vectors into mem_x / mem_y (the producer the
    reference keeps commented out, kernel.cl:587-588) and the pass runs at its call site.  Launch-by-launch mode against the
    C restatement with the same two lines enabled: every buffer and both motion-vector buffers over 24 frames."""
    octree, root = world
    rx, ry = 320, 192
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4, quality_passes=True)
    rc = svo.raycast
    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode="reference")
    try:
        rc.set_quality_passes(True)
        punched = 0
        for f in range(24):
            pos, rot = (10 + 0.4 * f, 22 + 0.05 * f, 9 + 0.3 * f), (0.4 + 0.002 * f, 0.7 + 0.015 * f, 0.0)
            before = None
            O.draw(pos, rot)
            rc.set_camera(pos, rot)
            rc.raycast_draw(rx, ry)
            screen, back, idb = rc.read_buffers(rx, ry)
            assert rc.idbuf_size() == O.idbuf_size, f"frame {f}"
            assert np.array_equal(rc.S.mem_x.to_numpy(np.int32, n), O.xbuf), f"frame {f} mem_x"
            assert np.array_equal(rc.S.mem_y.to_numpy(np.int32, n), O.ybuf), f"frame {f} mem_y"
            assert np.array_equal(idb[:2 * O.nblocks + O.idbuf_size], O.idbuf[:2 * O.nblocks + O.idbuf_size]), f"frame {f} ids"
            assert np.array_equal(screen, O.screen[:4 * n]), f"frame {f} colour"
            assert np.array_equal(back.view(np.uint32), O.back[:16 * n].view(np.uint32)), f"frame {f} xyz"
            assert np.array_equal(rc.read_frame(rx, ry).ravel(), O.tex), f"frame {f} tex"
        assert np.any(O.xbuf != 0) and np.any(O.ybuf != 0), "no motion vectors were produced"
    finally:
        rc.raycast_exit()
        svo.ocl_init(0)


def test_raycast_batch(svo, orc, world):
    """svo_raycast_batch (BASELINE config 5): five cameras traced back to back into buffers 0 / 2 on two alternating streams;
    afterwards the buffers hold the last even / odd camera's full raycast, bit for bit the oracle's raycast_fine_2."""
    octree, root = world
    rx, ry = 320, 192
    n = rx * ry
    rc = svo.raycast
    svo.ocl_exit()
    rc.raycast_init(octree, root, max_w=rx, max_h=ry, mode="fused")
    try:
        poses = [((10 + 0.9 * f, 22 + 0.3 * f, 9 + 0.7 * f), (0.4 + 0.01 * f, 0.7 + 0.05 * f, 0.0)) for f in range(5)]
        rc.full_raycast_batch(rx, ry, poses)
        screen, back, _ = rc.read_buffers(rx, ry)
        for slot, f in ((0, 4), (2, 3)):
            cam = ofr.camera_args(*poses[f])
            es = np.full(4 * n, HOLE, dtype=np.uint32)
            eb = np.zeros(16 * n, dtype=np.float32)
            orc.raycast_fine_2(es, eb, octree, root, rx, ry, 0, 0, 0, cam["v0"], *cam["cols"], gx=rx, gy=ry, threads=4)
            assert np.array_equal(screen[slot * n:(slot + 1) * n], es[:n]), f"slot {slot}"
            got = back[slot * 4 * n:(slot + 1) * 4 * n].view(np.uint32).reshape(n, 4)[:, :3]
            assert np.array_equal(got, eb[:4 * n].view(np.uint32).reshape(n, 4)[:, :3]), f"slot {slot} xyz"
    finally:
        rc.raycast_exit()
        svo.ocl_init(0)


def test_golden_frames(svo):
    """The CUDA path against the committed vectors produced by the reference's own source (tests/golden)."""
    from golden.make_golden import golden_pose, NFRAMES
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "frames_small.npz"))
    octree = np.zeros(int(gold["nwords"]), dtype=np.uint32)
    octree[:46810] = gold["normal_region"]
    octree[2097152:] = gold["blocks"]
    rx, ry = (int(v) for v in gold["res"])
    n = rx * ry
    rc = svo.raycast
    svo.ocl_exit()
    rc.raycast_init(octree, int(gold["root"]), max_w=rx, max_h=ry, mode="fused")
    try:
        for f in range(NFRAMES):
            rc.set_camera(*golden_pose(f))
            rc.raycast_draw(rx, ry)
            screen, back, idb = rc.read_buffers(rx, ry)
            ids = gold[f"f{f}_ids"]
            assert np.array_equal(screen, gold[f"f{f}_screen"])
            assert np.array_equal(idb[:len(ids)], ids)
            assert np.allclose(back, gold[f"f{f}_back"], rtol=RTOL, atol=1e-4)
            assert np.array_equal(rc.read_frame(rx, ry).ravel(), gold[f"f{f}_tex"])
    finally:
        rc.raycast_exit()
        svo.ocl_init(0)


def test_errors_are_reported(svo):
    with pytest.raises(RuntimeError):
        svo.ocl_get_kernel("no_such_kernel")
    k = svo.ocl_get_kernel("raycast_fine")          # launched without its 19 arguments
    svo.ocl_begin(k, 16, 16, 16, 16)
    with pytest.raises(RuntimeError):
        svo.ocl_end()
    k = svo.ocl_get_kernel("raycast_colorize")
    svo.ocl_begin(k, 16, 16, 16, 16)
    svo.ocl_param(C.c_int(1))
    with pytest.raises(RuntimeError):
        svo.ocl_end()                                # wrong argument count / sizes


def test_frame_sequence_depth_14(svo, orc):
    """OCTREE_DEPTH 14 (BASELINE.json config 4; a compile-time constant of the reference, a context setting here):
    full pipeline on a small fractal-terrain patch of a 16384^3 world, every buffer bit-exact against the oracle built
    with the same depth."""
    t = scenes.terrain(192, 6000, 9000, height=260, base=8000, seed=5)
    orc.lib.orc_set_depth(14)
    try:
        octree, root = orc.build_octree(*t)
        rx, ry = 320, 192
        n = rx * ry
        O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4, depth=14)
        rc = svo.raycast
        svo.ocl_exit()
        rc.raycast_init(octree, root, max_w=rx, max_h=ry, depth=14, mode="fused")
        try:
            for f in range(8):
                # world units = voxels / 8: look down at the patch from above, moving fast (high hole fraction)
                pos, rot = (6000 / 8 + 6 + 0.6 * f, 8000 / 8 + 60, 9000 / 8 + 4 + 0.4 * f), (0.9, 0.6 + 0.03 * f, 0.0)
                O.draw(pos, rot)
                rc.set_camera(pos, rot)
                rc.raycast_draw(rx, ry)
                screen, back, idb = rc.read_buffers(rx, ry)
                assert rc.idbuf_size() == O.idbuf_size, f"frame {f}"
                assert np.array_equal(idb[:2 * O.nblocks + O.idbuf_size], O.idbuf[:2 * O.nblocks + O.idbuf_size]), f"frame {f} ids"
                assert np.array_equal(screen, O.screen[:4 * n]), f"frame {f} colour"
                assert np.array_equal(back.view(np.uint32), O.back[:16 * n].view(np.uint32)), f"frame {f} xyz"
                assert np.array_equal(rc.read_frame(rx, ry).ravel(), O.tex), f"frame {f} tex"
            hit = (O.screen[:n] & 0xff) != 0
            assert 0.05 < hit.mean(), "the camera does not see the terrain patch"
        finally:
            rc.raycast_exit()
            svo.ocl_init(0)
    finally:
        orc.lib.orc_set_depth(11)


EDGE_SCENES = {
    "dense_cube": lambda: scenes.cube(1000, 1000, 1000, 40),                      # every byte-packed record full (k = 8)
    "sparse_cloud": lambda: scenes.random_cloud(4000, 900, 1200, 5),              # mostly single-voxel records (k = 1)
    "single_voxel": lambda: scenes.single_voxel(1024, 1024, 1024),
    "block_edges": lambda: scenes.concat(scenes.cube(1023 - 4, 1023 - 4, 1023 - 4, 10), scenes.single_voxel(1087, 1024, 1088)),
    "duplicates": scenes.duplicates,
}
EDGE_POSES = [((127.0, 128.5, 120.0), (0.05, 0.02, 0.0)),         # in front of the centre, nearly axis-aligned rays
              ((128.0, 128.0, 128.0), (1.5707963, 0.0, 0.0)),     # integer camera position, looking straight along an axis
              ((125.3, 135.2, 126.1), (0.9, 2.4, 0.3)),           # oblique, rolled
              ((-3.0, 300.0, 260.5), (0.7, 0.8, 0.0))]            # outside the world: position wraps (src/raycast.h:136-145)


@pytest.mark.parametrize("scene", sorted(EDGE_SCENES))
def test_raycast_edge_scenes(svo, orc, scene):
    """Octree edge cases (full / single-entry byte-packed records, voxels on block boundaries, duplicates) under
    axis-aligned, integer-position, rolled and wrapped cameras: raycast_fine_2 over the screen and raycast_holes over a
    ragged id list (length not a multiple of the work-group size), hit words and positions bit-exact."""
    octree, root = orc.build_octree(*EDGE_SCENES[scene]())
    rx, ry = 160, 96
    n, nb = rx * ry, (rx // 16) * (ry // 16)
    mo = svo.ocl_malloc(octree.nbytes, octree)
    dead = (0, 0, 0, 0)
    rng = np.random.RandomState(11)
    for pose in EDGE_POSES:
        cam = ofr.camera_args(*pose)
        screen = np.full(4 * n, HOLE, dtype=np.uint32)
        back = np.zeros(16 * n, dtype=np.float32)
        orc.raycast_fine_2(screen, back, octree, root, rx, ry, 0, 0, 0, cam["v0"], *cam["cols"], gx=rx, gy=ry, threads=4)
        ms = svo.ocl_malloc(4 * n * 4, np.full(4 * n, HOLE, dtype=np.uint32))
        mb = svo.ocl_malloc(16 * n * 4, np.zeros(16 * n, dtype=np.float32))
        launch(svo, "raycast_fine_2", rx, ry, 16, 16,
               [ms, mb, mo, C.c_uint32(root), i32(rx), i32(ry), i32(0), i32(0), i32(0), dead, dead, dead, dead,
                cam["v0"], *cam["cols"], C.c_float(1.0), C.c_float(1.0)])
        assert np.array_equal(ms.to_numpy(), screen), (scene, pose)
        assert np.array_equal(mb.to_numpy(np.uint32), back.view(np.uint32)), (scene, pose)
        # raycast_holes over a ragged list of 1003 random pixels (and an empty list: no launch, src/raycast.h:330)
        for count in (1003, 0):
            ids = np.zeros(2 * nb + n, dtype=np.uint32)
            px = rng.randint(0, rx, size=count).astype(np.uint32)
            py = rng.randint(0, ry, size=count).astype(np.uint32)
            ids[2 * nb:2 * nb + count] = px | (py << 16)
            ids[0] = count
            s2 = np.full(4 * n, HOLE, dtype=np.uint32)
            b2 = np.zeros(16 * n, dtype=np.float32)
            mi = svo.ocl_malloc(ids.nbytes, ids)
            svo.ocl_copy_to_device(ms, 0, s2)
            svo.ocl_copy_to_device(mb, 0, b2)
            if count:
                orc.raycast_holes(s2, b2, octree, ids, root, rx, ry, 0, count, cam["v0"], *cam["cols"], threads=4)
                launch(svo, "raycast_holes", count, 1, 256, 1,
                       [ms, mb, mo, None, None, None, mi, C.c_uint32(root), i32(rx), i32(ry), i32(0), i32(count),
                        dead, dead, dead, dead, cam["v0"], *cam["cols"], C.c_float(1.0), C.c_float(1.0)])
            assert np.array_equal(ms.to_numpy(), s2), (scene, pose, count)
            assert np.array_equal(mb.to_numpy(np.uint32), b2.view(np.uint32)), (scene, pose, count)
            mi.free()
        ms.free(); mb.free()
    mo.free()


@pytest.mark.parametrize("res", [(80, 48), (200, 120), (1920, 1024)])
def test_disabled_kernel_fillhole(svo, orc, res):
    """raycast_fillhole (kernel.cl:342-401; if(0) at its call site, src/raycast.h:205) through the kernel API with the
    argument list of that call site: depth-discontinuity hole punch, bit-exact."""
    rx, ry = res
    n = rx * ry
    rng = np.random.RandomState(rx)
    screen = random_screen(rng, 4 * n, 0.2)
    back = rng.rand(16 * n).astype(np.float32)
    back[3:4 * n:4] = np.where(rng.rand(n) < 0.5, 10.0, 10.0 + 30.0 * rng.rand(n)).astype(np.float32)
    xb = (rng.randint(0, 3, size=n) - 1).astype(np.int32)
    yb = rng.randint(0, 2, size=n).astype(np.int32)
    exp = screen.copy()
    orc.raycast_fillhole(exp, back, xb, yb, rx, ry, threads=4)
    ms, mb = svo.ocl_malloc(screen.nbytes, screen), svo.ocl_malloc(back.nbytes, back)
    mx, my, mz = svo.ocl_malloc(xb.nbytes, xb), svo.ocl_malloc(yb.nbytes, yb), svo.ocl_malloc(4 * n)
    launch(svo, "raycast_fillhole", rx, ry, 16, 16, [ms, mb, mx, my, mz, i32(rx), i32(ry), i32(4)])
    got = ms.to_numpy()
    assert (exp != screen).sum() > 0
    assert np.array_equal(got, exp)
    for m in (ms, mb, mx, my, mz):
        m.free()


@pytest.mark.parametrize("frame", [0, 1, 2, 3, 4, 8, 12])
def test_disabled_kernel_fine(svo, orc, world, frame):
    """raycast_fine (kernel.cl:696-843; if(0) at its call site, src/raycast.h:234) with that call site's launch geometry
    and argument list: a quarter of the screen, one ray per 2x2 cell, holes first."""
    octree, root = world
    rx, ry = 320, 192
    n = rx * ry
    cam = ofr.camera_args((10, 22, 9), (0.4, 0.7, 0.0))
    rng = np.random.RandomState(frame)
    screen = rng.randint(0, 2 ** 24, size=4 * n).astype(np.uint32)
    screen[rng.rand(4 * n) < 0.4] = HOLE
    back = np.zeros(16 * n, dtype=np.float32)
    add_x, add_y = (rx // 2) * (frame & 1), (ry // 2) * ((frame >> 1) & 1)              # src/raycast.h:236-237
    exp_s, exp_b = screen.copy(), back.copy()
    orc.raycast_fine(exp_s, exp_b, octree, root, rx, ry, frame, add_x, add_y, cam["v0"], *cam["cols"])
    mo, ms, mb = svo.ocl_malloc(octree.nbytes, octree), svo.ocl_malloc(screen.nbytes, screen), svo.ocl_malloc(back.nbytes, back)
    dead = (0, 0, 0, 0)
    launch(svo, "raycast_fine", rx // 4, ry // 4, 16, 16,
           [ms, mb, mo, C.c_uint32(root), i32(rx), i32(ry), i32(frame), i32(add_x), i32(add_y), dead, dead, dead, dead,
            cam["v0"], *cam["cols"], C.c_float(1.0), C.c_float(1.0)])
    assert np.array_equal(ms.to_numpy(), exp_s)
    assert np.array_equal(mb.to_numpy(np.uint32), exp_b.view(np.uint32))
    for m in (mo, ms, mb):
        m.free()


@pytest.mark.parametrize("name,depth", [("small_world", 11), ("duplicates", 11), ("single", 11), ("cloud", 11), ("dense", 11),
                                        ("zero_colour", 11), ("terrain14", 14)])
def test_device_octree_builder(svo, orc, name, depth):
    """svo_octree_build_device: the node pool built on the GPU is, word for word, the array of the
    reference's set_voxel + convert_tree_blocks (via the oracle), including duplicate insertions, zero colour words, a
    single voxel and OCTREE_DEPTH 14."""
    rng = np.random.RandomState(4)
    sc = {"small_world": scenes.small_world, "duplicates": scenes.duplicates, "single": scenes.single_voxel,
          "cloud": lambda: scenes.random_cloud(20000, 0, 2048, 9), "dense": lambda: scenes.cube(1000, 1000, 1000, 70),
          "zero_colour": lambda: (np.array([7, 8, 9], np.uint32), np.array([1, 1, 1], np.uint32), np.array([5, 5, 5], np.uint32),
                                  np.array([0x11, 0, 0x21], np.uint32)),
          "terrain14": lambda: scenes.concat(scenes.terrain(256, 5000, 9000, height=300, base=8000, seed=2),
                                             scenes.random_cloud(30000, 0, 16384, 4))}[name]()
    orc.lib.orc_set_depth(depth)
    try:
        exp, root = orc.build_octree(*sc)
    finally:
        orc.lib.orc_set_depth(11)
    mem, got_root, st = svo.scene.build_octree_device(*sc, depth=depth)
    try:
        assert got_root == root
        assert mem.size == exp.nbytes
        assert np.array_equal(mem.to_numpy(), exp)
        assert st["num_unique"] <= len(sc[0])
    finally:
        mem.free()


@pytest.mark.parametrize("case", ["s1", "terrain_offsets_palette", "mip1", "truncated"])
def test_device_rle4_decode(svo, orc, tmp_path, case):
    """svo_octree_load_rle4_device: the .rle4 slab stream decoded on the GPU and fed to the device
    builder gives, word for word, the octree of the host loader + host builder (which the CPU tests pin against the
    reference's RLE4::load + convert_tree_blocks) -- with offsets and the palette colour mapping, for a higher mip volume
    of a multi-mip file, and for a stream cut off in the middle of a column."""
    from test_host_scene import stitch_mips
    sc = svo.scene
    kw = dict(mip=0, palette=0, addx=0, addy=0, addz=0)
    if case == "s1":
        v = sc.generate(kind=1, depth=11, size=320, nblobs=3, seed=11)
        path = str(tmp_path / "s1.rle4")
        v.write_rle4(path, 320, 2048, 320)
    elif case == "terrain_offsets_palette":
        v = sc.generate(kind=2, depth=11, size=200, nblobs=0, seed=12)
        path = str(tmp_path / "t.rle4")
        v.write_rle4(path, 200, 1024, 200)
        kw.update(palette=1, addx=100, addy=300, addz=777)
    elif case == "mip1":
        v0 = sc.generate(kind=2, depth=11, size=96, nblobs=0, seed=13)
        v = sc.generate(kind=1, depth=10, size=160, nblobs=2, seed=14)
        p0, p1, path = (str(tmp_path / n) for n in ("a.rle4", "b.rle4", "ab.rle4"))
        v0.write_rle4(p0, 96, 2048, 96)
        v.write_rle4(p1, 160, 1024, 160)
        stitch_mips([p0, p1], path)
        v0.free()
        kw.update(mip=1)
    else:
        v = sc.generate(kind=1, depth=11, size=128, nblobs=1, seed=15)
        full = str(tmp_path / "full.rle4")
        v.write_rle4(full, 128, 2048, 128)
        raw = open(full, "rb").read()
        cut = (len(raw) - 20) * 3 // 5 // 2 * 2                       # keep the header's slab count: the stream simply ends early
        path = str(tmp_path / "cut.rle4")
        import struct
        hdr = bytearray(raw[:20])
        hdr[16:20] = struct.pack("<i", cut // 2)
        open(path, "wb").write(bytes(hdr) + raw[20:20 + cut])
    v.free()
    host = sc.rle4_load(path, kw["palette"], kw["addx"], kw["addy"], kw["addz"], mip=kw["mip"])
    exp, exp_root, st = sc.build_octree_voxels(host, depth=11)
    nvox = len(host.arrays()[0])
    host.free()
    mem, root, dst = sc.octree_init_device(path, depth=11, **kw)
    try:
        assert dst["num_voxels"] == nvox and dst["num_unique"] == st["num_unique"]
        assert root == exp_root and mem.size == exp.nbytes
        assert np.array_equal(mem.to_numpy(), exp)
    finally:
        mem.free()
