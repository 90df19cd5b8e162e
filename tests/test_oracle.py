"""CPU tests (no GPU): the plain-C oracle (oracle/svo_oracle.c) against
 (1) the golden vectors the reference's own source produced (tests/golden/frames_small.npz), and
 (2) the reference's own source run live (oracle/_ref), where it is built."""
import hashlib
import os

import numpy as np
import pytest

import scenes
from oracle import frame as fr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "frames_small.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_octree_layout_matches_golden(orc, gold):
    octree, root = orc.build_octree(gold["x"], gold["y"], gold["z"], gold["rgba"])
    assert root == int(gold["root"])
    assert len(octree) == int(gold["nwords"])
    assert np.array_equal(octree[:46810], gold["normal_region"])
    assert np.array_equal(octree[2097152:], gold["blocks"])
    assert hashlib.sha256(octree.tobytes()).digest() == gold["sha256"].tobytes()


def test_frames_match_golden(orc, gold):
    from golden.make_golden import golden_pose, NFRAMES
    octree, root = orc.build_octree(gold["x"], gold["y"], gold["z"], gold["rgba"])
    rx, ry = (int(v) for v in gold["res"])
    F = fr.OracleFrame(orc, octree, root, rx, ry)
    for f in range(NFRAMES):
        F.draw(*golden_pose(f))
        n = F.n
        assert np.array_equal(F.screen[:4 * n], gold[f"f{f}_screen"]), f"screen frame {f}"
        assert np.array_equal(F.back[:16 * n].view(np.uint32), gold[f"f{f}_back"].view(np.uint32)), f"xyz frame {f}"
        assert np.array_equal(F.idbuf[:2 * F.nblocks + F.idbuf_size], gold[f"f{f}_ids"]), f"ids frame {f}"
        assert np.array_equal(F.tex, gold[f"f{f}_tex"]), f"tex frame {f}"


def test_frame0_is_full_raycast_in_block_order(orc, gold):
    """frame 0 gathers every pixel, ids run block by block, 256 consecutive ids per block."""
    octree, root = orc.build_octree(gold["x"], gold["y"], gold["z"], gold["rgba"])
    F = fr.OracleFrame(orc, octree, root, 64, 48)
    F.draw((6, 22, 5), (0.45, 0.75, 0), stop_after="ids")
    assert F.idbuf_size == 64 * 48
    ids = F.idbuf[2 * F.nblocks:2 * F.nblocks + 8]
    assert [(int(v) & 0xffff, int(v) >> 16) for v in ids[:5]] == [(0, 0), (1, 0), (1, 1), (0, 1), (2, 0)]
    assert int(F.idbuf[2 * F.nblocks + 255]) == (14 | (15 << 16))
    assert int(F.idbuf[2 * F.nblocks + 256]) == 16


@pytest.mark.parametrize("scene", ["small_world", "duplicates", "single", "cloud"])
def test_builder_matches_reference(orc, ref, scene):
    sc = {"small_world": scenes.small_world, "duplicates": scenes.duplicates,
          "single": scenes.single_voxel, "cloud": lambda: scenes.random_cloud(20000, 0, 2048, 9)}[scene]()
    a, ra = ref.build_octree(*sc)
    b, rb = orc.build_octree(*sc)
    assert ra == rb and np.array_equal(a, b)
    assert ref.num_voxels() == orc.num_voxels() and ref.num_nodes() == orc.num_nodes()


@pytest.mark.parametrize("pose", [((1, 50, 1), (0.6, 0.8, 0)), ((12, 20, 12), (0.3, 0.7, 0)),
                                  ((100, 120, 30), (-0.4, 2.5, 0.1)), ((20, 3, 20), (1.2, -0.7, 0))])
def test_frame_sequence_matches_reference(orc, ref, pose):
    octree, root = ref.build_octree(*scenes.small_world())
    rx, ry = 160, 96
    A, B = fr.OracleFrame(ref, octree, root, rx, ry, threads=4), fr.OracleFrame(orc, octree, root, rx, ry, threads=4)
    pos, rot = pose
    for f in range(5):
        p = (pos[0] + 0.4 * f, pos[1], pos[2] + 0.3 * f)
        r = (rot[0], rot[1] + 0.015 * f, rot[2])
        A.draw(p, r); B.draw(p, r)
        assert A.idbuf_size == B.idbuf_size
        assert np.array_equal(A.screen, B.screen)
        assert np.array_equal(A.back.view(np.uint32), B.back.view(np.uint32))
        assert np.array_equal(A.idbuf[:2 * A.nblocks + A.idbuf_size], B.idbuf[:2 * B.nblocks + B.idbuf_size])
        assert np.array_equal(A.tex, B.tex)


def test_fillhole2_snapshot_random_images(orc, ref):
    rng = np.random.RandomState(5)
    rx, ry = 80, 48
    n = rx * ry
    for density in (0.02, 0.3, 0.8, 0.97):
        img = rng.randint(0, 2 ** 32 - 1, size=4 * n, dtype=np.uint64).astype(np.uint32)
        img[rng.rand(4 * n) < density] = 0xFFFFFF00
        a, b = img.copy(), img.copy()
        ref.raycast_fillhole2(a, rx, ry)
        orc.raycast_fillhole2(b, rx, ry)
        assert np.array_equal(a, b)
        assert np.array_equal(a[n:], img[n:])      # only buffer 0 is written


def test_rle4_roundtrip(orc, ref, tmp_path):
    import rle4
    sc = scenes.terrain(96, 0, 0, height=40, base=10, seed=3)
    path = str(tmp_path / "t.rle4")
    vox = rle4.write_rle4(path, sc[0], sc[1], sc[2], sc[3], 128, 256, 128)
    a, ra = ref.build_octree_rle4(path)
    b, rb = orc.build_octree_rle4(path)
    assert ra == rb and np.array_equal(a, b)
    assert ref.num_voxels() == orc.num_voxels() == vox
