"""GPU parity tests of the multi-GPU screen-band frame: N bands must reproduce the 1-GPU / reference
result bit for bit.  On the single-GPU test box the bands are "virtual" -- G contexts on device 0, exchanging through the
same peer-pointer tables, atomics and flag barriers as G GPUs do; with >= 2 devices the same tests also run across real
GPUs (NVLink peer access)."""
import os

import numpy as np
import pytest

import scenes
from oracle import frame as ofr

pytestmark = pytest.mark.gpu
HOLE = 0xFFFFFF00


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
    return orc.build_octree(*scenes.small_world())


def pose(f):
    return (10 + 0.25 * f, 22 + 0.05 * f, 9 + 0.2 * f), (0.4 + 0.002 * f, 0.7 + 0.01 * f, 0.0)


def frame_params(svo, rx, ry, f, pingpong=False):
    rc = svo.raycast
    rc.S.mode = "pingpong" if pingpong else "fused"
    rc.set_camera(*pose(f))
    return rc.prepare_params(rx, ry, f)


def devices_for(svo, G, real):
    n = svo.Device.count()
    if real:
        if n < 2:
            pytest.skip("needs >= 2 CUDA devices")
        return [r % n for r in range(G)]
    return [0] * G


def check_ids(bs, O):
    nb = O.nblocks
    counts, ids = bs.merged_ids()
    exp_off = O.idbuf[nb:2 * nb].astype(np.int64)
    exp_counts = np.diff(np.append(exp_off, O.idbuf_size)).astype(np.uint32)
    assert int(counts.sum()) == O.idbuf_size
    assert np.array_equal(counts, exp_counts)
    assert np.array_equal(ids, O.idbuf[2 * nb:2 * nb + O.idbuf_size])


@pytest.mark.parametrize("real", [False, True], ids=["virtual", "multi-gpu"])
@pytest.mark.parametrize("G,sr,res,nframes", [(1, 0, (320, 192), 6), (2, 0, (320, 192), 36), (3, 16, (320, 192), 12),
                                              (4, 32, (200, 120), 8), (8, 16, (320, 192), 6), (2, 64, (1920, 1024), 4)])
def test_band_frames_match_oracle(svo, orc, world, real, G, sr, res, nframes):
    """Exact mode: every buffer of every frame, assembled from the ranks' rows, equals the reference's."""
    octree, root = world
    rx, ry = res
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4)
    bs = svo.bands.LocalBandSet(devices_for(svo, G, real), octree, root, rx, ry, stripe_rows=sr)
    try:
        for f in range(nframes):
            O.draw(*pose(f))
            bs.frame(frame_params(svo, rx, ry, f))
            screen, back = bs.assemble()
            check_ids(bs, O)
            assert np.array_equal(screen, O.screen[:4 * n]), f"frame {f} colour"
            assert np.array_equal(back.view(np.uint32), O.back[:16 * n].view(np.uint32)), f"frame {f} xyz"
            assert np.array_equal(bs.frame_image(), O.tex), f"frame {f} tex"
    finally:
        bs.close()


@pytest.mark.parametrize("real", [False, True], ids=["virtual", "multi-gpu"])
@pytest.mark.parametrize("pingpong", [False, True], ids=["exact", "pingpong"])
@pytest.mark.parametrize("G,sr,res,nframes", [(1, 0, (320, 192), 40), (2, 16, (320, 192), 40), (4, 32, (200, 120), 9), (2, 64, (1920, 1024), 6)])
def test_band_frames_back_to_back(svo, orc, world, real, pingpong, G, sr, res, nframes):
    """The schedule the bench runs: frames enqueued back to back on every rank with nothing observed in between -- the scatter
    carries the previous frame's cache copy, the gap filter runs beside the next frame, the filter's in-place write and the
    last copy are issued only when the buffers are read at the end.  Everything must still be the reference's, bit for bit."""
    octree, root = world
    rx, ry = res
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4)
    bs = svo.bands.LocalBandSet(devices_for(svo, G, real), octree, root, rx, ry, stripe_rows=sr)
    try:
        for f in range(nframes):
            O.draw(*pose(f))
            bs.frame(frame_params(svo, rx, ry, f, pingpong=pingpong), sync=False)
        bs.sync()
        if not pingpong:
            assert all(b.deferred_count() == nframes - 2 for b in bs.bands), [b.deferred_count() for b in bs.bands]
        assert np.array_equal(bs.frame_image(), O.tex), "tex"
        check_ids(bs, O)
        screen, back = bs.assemble()
        if pingpong:
            slot = bs.bands[0].last_slot()
            assert np.array_equal(screen[slot * n:(slot + 1) * n], O.screen[2 * n:3 * n]), "colour"
            got = back[slot * 4 * n:(slot + 1) * 4 * n].view(np.uint32).reshape(n, 4)
            exp = O.back[8 * n:12 * n].view(np.uint32).reshape(n, 4)
            live = O.screen[2 * n:3 * n] != HOLE
            assert np.array_equal(got[live, :3], exp[live, :3]), "xyz"
        else:
            assert np.array_equal(screen, O.screen[:4 * n]), "colour"
            assert np.array_equal(back.view(np.uint32), O.back[:16 * n].view(np.uint32)), "xyz"
    finally:
        svo.raycast.S.mode = "fused"
        bs.close()


@pytest.mark.parametrize("real", [False, True], ids=["virtual", "multi-gpu"])
@pytest.mark.parametrize("G,sr,res,nframes", [(2, 0, (320, 192), 12), (4, 16, (320, 192), 34)])
def test_band_frames_pingpong(svo, orc, world, real, G, sr, res, nframes):
    """SVO_FRAME_PINGPONG across bands: ids and the colorized frame are the reference's, the slot rendered into holds the
    reference's cache buffer 2 (same comparison as the 1-GPU ping-pong test)."""
    octree, root = world
    rx, ry = res
    n = rx * ry
    O = ofr.OracleFrame(orc, octree, root, rx, ry, threads=os.cpu_count() or 4)
    bs = svo.bands.LocalBandSet(devices_for(svo, G, real), octree, root, rx, ry, stripe_rows=sr)
    try:
        for f in range(nframes):
            O.draw(*pose(f))
            bs.frame(frame_params(svo, rx, ry, f, pingpong=True))
            slot = bs.bands[0].last_slot()
            assert slot == (0 if f % 2 == 0 else 2)
            screen, back = bs.assemble()
            check_ids(bs, O)
            assert np.array_equal(screen[slot * n:(slot + 1) * n], O.screen[2 * n:3 * n]), f"frame {f} colour"
            got = back[slot * 4 * n:(slot + 1) * 4 * n].view(np.uint32).reshape(n, 4)
            exp = O.back[8 * n:12 * n].view(np.uint32).reshape(n, 4)
            live = O.screen[2 * n:3 * n] != HOLE
            assert np.array_equal(got[live, :3], exp[live, :3]), f"frame {f} xyz"
            assert np.array_equal(bs.frame_image(), O.tex), f"frame {f} tex"
    finally:
        svo.raycast.S.mode = "fused"
        bs.close()


@pytest.mark.parametrize("real", [False, True], ids=["virtual", "multi-gpu"])
@pytest.mark.parametrize("G,sr", [(1, 0), (2, 16), (8, 16)])
def test_band_full_raycast_pingpong(svo, orc, world, real, G, sr):
    """SVO_FRAME_PINGPONG raycasts: consecutive frames alternate between buffers 0 and 2 (and between two streams, so that
    they overlap) and between the writer's two frames.  Five frames are issued back to back; afterwards the last two
    frames' hit words and positions sit in their slots and the image read back is the last frame's."""
    octree, root = world
    rx, ry = 320, 200
    n = rx * ry
    bs = svo.bands.LocalBandSet(devices_for(svo, G, real), octree, root, rx, ry, stripe_rows=sr)
    try:
        ref = []
        for f in range(5):
            cam = ofr.camera_args(*pose(f))
            screen = np.full(4 * n, HOLE, dtype=np.uint32)
            back = np.zeros(16 * n, dtype=np.float32)
            orc.raycast_fine_2(screen, back, octree, root, rx, ry, 0, 0, 0, cam["v0"], *cam["cols"], gx=rx, gy=ry, threads=4)
            ref.append((screen[:n].copy(), back[:4 * n].copy()))
            svo.raycast.set_camera(*pose(f))
            p = svo.raycast.prepare_params(rx, ry, 0)
            p.flags |= svo.ocl.FRAME_PINGPONG
            bs.raycast(p, sync=False)
        bs.sync()
        assert bs.bands[0].last_slot() == 0                      # frames 0, 2, 4 -> buffer 0; 1, 3 -> buffer 2
        got_s, got_b = bs.assemble(slots=(0, 2))
        for slot, f in ((0, 4), (2, 3)):
            assert np.array_equal(got_s[slot * n:(slot + 1) * n], ref[f][0]), f"slot {slot}"
            assert np.array_equal(got_b[slot * 4 * n:(slot + 1) * 4 * n].view(np.uint32).reshape(n, 4)[:, :3],
                                  ref[f][1].view(np.uint32).reshape(n, 4)[:, :3]), f"slot {slot} xyz"
        tex = np.zeros(n, dtype=np.uint32)
        full = np.full(4 * n, HOLE, dtype=np.uint32)
        full[:n] = ref[4][0]
        orc.raycast_colorize(full, tex, rx, ry)
        assert np.array_equal(bs.frame_image(), tex)
    finally:
        bs.close()


@pytest.mark.parametrize("real", [False, True], ids=["virtual", "multi-gpu"])
@pytest.mark.parametrize("G,sr", [(1, 0), (2, 0), (4, 16), (8, 16)])
def test_band_full_raycast(svo, orc, world, real, G, sr):
    """Banded full raycast (BASELINE.json config 3): hit words, positions and the colorized frame of a full-screen
    raycast_fine_2, every rank tracing only its stripes."""
    octree, root = world
    rx, ry = 320, 200                       # 200 rows: the last stripe is partial
    n = rx * ry
    cam = ofr.camera_args(*pose(3))
    screen = np.full(4 * n, HOLE, dtype=np.uint32)
    back = np.zeros(16 * n, dtype=np.float32)
    orc.raycast_fine_2(screen, back, octree, root, rx, ry, 0, 0, 0, cam["v0"], *cam["cols"], gx=rx, gy=ry, threads=4)
    tex = np.zeros(n, dtype=np.uint32)
    orc.raycast_colorize(screen, tex, rx, ry)
    bs = svo.bands.LocalBandSet(devices_for(svo, G, real), octree, root, rx, ry, stripe_rows=sr)
    try:
        svo.raycast.set_camera(*pose(3))
        bs.raycast(svo.raycast.prepare_params(rx, ry, 0))
        got_s, got_b = bs.assemble(slots=(0,))
        assert np.array_equal(got_s[:n], screen[:n])
        assert np.array_equal(got_b[:4 * n].view(np.uint32).reshape(n, 4)[:, :3], back[:4 * n].view(np.uint32).reshape(n, 4)[:, :3])
        assert np.array_equal(bs.frame_image(), tex)
    finally:
        bs.close()
