"""CPU tests of the host-side scene contract (include/svo_host.h): the direct compact-octree builder must produce,
word for word, the array of the reference's set_voxel + convert_tree_blocks (src/octree/octree.h), the .rle4 loader must
decode what RLE4::load decodes (src/octree/Rle4.cpp), and the writer must round-trip.  No GPU needed."""
import os

import numpy as np
import pytest

import scenes


@pytest.fixture(scope="module")
def svo():
    from __graft_entry__ import load_package
    return load_package()


SCENES = {"small_world": scenes.small_world, "duplicates": scenes.duplicates, "single": scenes.single_voxel,
          "cloud": lambda: scenes.random_cloud(20000, 0, 2048, 9), "cube_edge": lambda: scenes.cube(2040, 2040, 2040, 8),
          "zero_colour": lambda: (np.array([7, 8, 9], np.uint32), np.array([1, 1, 1], np.uint32), np.array([5, 5, 5], np.uint32),
                                  np.array([0x11, 0, 0x21], np.uint32))}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_builder_matches_oracle(svo, orc, name):
    sc = SCENES[name]()
    a, ra = orc.build_octree(*sc)
    b, rb, st = svo.scene.build_octree(*sc)
    assert ra == rb
    assert len(a) == len(b) and np.array_equal(a, b)
    assert st["num_voxels"] == len(sc[0])


@pytest.mark.parametrize("name", ["small_world", "duplicates", "cloud"])
def test_builder_matches_reference_source(svo, ref, name):
    sc = SCENES[name]()
    a, ra = ref.build_octree(*sc)
    b, rb, _ = svo.scene.build_octree(*sc)
    assert ra == rb and np.array_equal(a, b)


def test_builder_matches_golden(svo):
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "frames_small.npz"))
    b, rb, _ = svo.scene.build_octree(gold["x"], gold["y"], gold["z"], gold["rgba"])
    assert rb == int(gold["root"]) and len(b) == int(gold["nwords"])
    assert np.array_equal(b[:46810], gold["normal_region"]) and np.array_equal(b[2097152:], gold["blocks"])


def test_depth_14_builder_matches_oracle(svo, orc):
    """OCTREE_DEPTH is a parameter here (compile-time 11 in the reference): same layout rules at depth 14."""
    rng = np.random.RandomState(4)
    p = rng.randint(0, 16384, size=(30000, 3)).astype(np.uint32)
    t = scenes.terrain(256, 5000, 9000, height=300, base=8000, seed=2)
    x = np.concatenate([p[:, 0], t[0]]); y = np.concatenate([p[:, 1], t[1]]); z = np.concatenate([p[:, 2], t[2]])
    c = np.concatenate([1 + (p[:, 0] & 0xfc), t[3]]).astype(np.uint32)
    orc.lib.orc_set_depth(14)
    try:
        a, ra = orc.build_octree(x, y, z, c)
    finally:
        orc.lib.orc_set_depth(11)
    b, rb, _ = svo.scene.build_octree(x, y, z, c, depth=14)
    assert ra == rb and np.array_equal(a, b)


def test_rle4_loader_and_writer(svo, orc, tmp_path):
    vox = svo.scene.generate(kind=1, depth=11, size=256, nblobs=2, seed=7)
    x, y, z, c = vox.arrays()
    assert len(x) > 50000 and (c & 255).min() >= 1
    path = str(tmp_path / "standin.rle4")
    vox.write_rle4(path, 256, 2048, 256)
    # product loader == stream that was written
    v2 = svo.scene.rle4_load(path)
    x2, y2, z2, c2 = v2.arrays()
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and np.array_equal(z, z2) and np.array_equal(c, c2)
    # oracle loader (restatement of RLE4::load) builds the same octree from the file as the product does
    a, ra = orc.build_octree_rle4(path)
    b, rb, _ = svo.scene.octree_init(path)
    assert ra == rb and np.array_equal(a, b)
    vox.free(); v2.free()


def test_rle4_loader_matches_reference_source(svo, ref, tmp_path):
    vox = svo.scene.generate(kind=2, depth=11, size=192, nblobs=0, seed=3)
    path = str(tmp_path / "terrain.rle4")
    vox.write_rle4(path, 192, 2048, 192)
    a, ra = ref.build_octree_rle4(path)
    b, rb, _ = svo.scene.octree_init(path)
    assert ra == rb and np.array_equal(a, b)
    vox.free()


def stitch_mips(paths, out):
    """Concatenate single-mip .rle4 files into one file with len(paths) mip volumes (Rle4.cpp:18-43: int32 nummaps, then
    per map sx, sy, sz, slabs_size and the slab stream, back to back)."""
    import struct
    bodies = []
    for p in paths:
        b = open(p, "rb").read()
        assert struct.unpack("<i", b[:4])[0] == 1
        bodies.append(b[4:])
    with open(out, "wb") as f:
        f.write(struct.pack("<i", len(bodies)))
        for b in bodies:
            f.write(b)


def test_rle4_mip_volumes(svo, orc, tmp_path):
    """svo_rle4_load_mip: every mip volume of a multi-mip file decodes to the stream it was written from; mip 0 is what the
    reference voxelises (the oracle's loader reads the same file)."""
    v0 = svo.scene.generate(kind=1, depth=11, size=128, nblobs=1, seed=5)
    v1 = svo.scene.generate(kind=2, depth=10, size=64, nblobs=0, seed=6)
    p0, p1, pm = (str(tmp_path / n) for n in ("m0.rle4", "m1.rle4", "mips.rle4"))
    v0.write_rle4(p0, 128, 2048, 128)
    v1.write_rle4(p1, 64, 1024, 64)
    stitch_mips([p0, p1], pm)
    for mip, v in ((0, v0), (1, v1)):
        got = svo.scene.rle4_load(pm, mip=mip)
        for a, b in zip(v.arrays(), got.arrays()):
            assert np.array_equal(a, b)
        got.free()
    a, ra = orc.build_octree_rle4(pm)
    b, rb, _ = svo.scene.octree_init(pm)
    assert ra == rb and np.array_equal(a, b)
    with pytest.raises(RuntimeError):
        svo.scene.rle4_load(pm, mip=2)
    v0.free(); v1.free()


def test_missing_rle4_is_an_error(svo, tmp_path):
    with pytest.raises(RuntimeError):
        svo.scene.rle4_load(str(tmp_path / "nope.rle4"))


def test_png_writer_round_trips(tmp_path):
    """svo_write_png (display/encode stage): a standards-conforming 8-bit RGB PNG without a compression
    library -- decoded here with zlib + the PNG chunk rules and compared pixel for pixel, flipped and not."""
    import ctypes as C
    import struct
    import zlib
    from __graft_entry__ import load_package
    lib = load_package().ocl.lib
    lib.svo_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(3)
    for (w, h), flip in (((37, 11), 0), ((320, 192), 1), ((1920, 1024), 1)):        # the last one spans many 64 KiB stored blocks
        px = rng.integers(0, 1 << 24, size=(h, w), dtype=np.uint32)
        path = str(tmp_path / f"t_{w}x{h}.png")
        assert lib.svo_write_png(path.encode(), px.ctypes.data, w, h, flip) == 0
        blob = open(path, "rb").read()
        assert blob[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(blob):
            n, typ = struct.unpack(">I4s", blob[pos:pos + 8])
            data = blob[pos + 8:pos + 8 + n]
            assert struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + data) & 0xffffffff, typ
            chunks.append((typ, data))
            pos += 12 + n
        assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        assert struct.unpack(">IIBBBBB", chunks[0][1]) == (w, h, 8, 2, 0, 0, 0)
        raw = np.frombuffer(zlib.decompress(chunks[1][1]), dtype=np.uint8).reshape(h, 1 + 3 * w)
        assert not raw[:, 0].any()                                                       # filter type 0 on every scanline
        rgb = raw[:, 1:].reshape(h, w, 3).astype(np.uint32)
        got = (rgb[..., 0] << 16) | (rgb[..., 1] << 8) | rgb[..., 2]
        assert np.array_equal(got, px[::-1] if flip else px)
