"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libsvo_ref.so =
/root/reference/kernel/kernel.cl + src/octree/octree.h + src/octree/Rle4.cpp compiled through
oracle/ref_shim).  Run in the build container only (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

The vectors pin the C restatement (oracle/svo_oracle.c) and, on the GPU box, the CUDA kernels.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import binding, frame as fr  # noqa: E402
import scenes  # noqa: E402

RES = (96, 64)
NFRAMES = 6


def golden_scene():
    return scenes.concat(scenes.terrain(192, 0, 0, height=90, base=30, seed=11),
                         scenes.blob(90, 150, 100, 24, seed=5),
                         scenes.cube(40, 120, 40, 9),
                         scenes.single_voxel(2047, 2047, 2047),
                         scenes.duplicates())


def golden_pose(f):
    return (6.0 + 0.35 * f, 22.0, 5.0 + 0.3 * f), (0.45, 0.75 + 0.02 * f, 0.0)


def main():
    ref = binding.get("ref")
    sc = golden_scene()
    octree, root = ref.build_octree(*sc)
    used = int(np.flatnonzero(octree)[-1]) + 1
    out = dict(x=sc[0], y=sc[1], z=sc[2], rgba=sc[3], root=np.uint32(root), nwords=np.uint64(len(octree)),
               sha256=np.frombuffer(hashlib.sha256(octree.tobytes()).digest(), dtype=np.uint8),
               normal_region=octree[:4681 * 10].copy(), blocks=octree[2097152:].copy(), used=np.uint64(used),
               res=np.array(RES))
    F = fr.OracleFrame(ref, octree, root, RES[0], RES[1])
    for f in range(NFRAMES):
        pos, rot = golden_pose(f)
        F.draw(pos, rot)
        n = F.n
        out[f"f{f}_screen"] = F.screen[:4 * n].copy()
        out[f"f{f}_back"] = F.back[:16 * n].copy()
        out[f"f{f}_ids"] = F.idbuf[:2 * F.nblocks + F.idbuf_size].copy()
        out[f"f{f}_tex"] = F.tex.copy()
    np.savez_compressed(os.path.join(HERE, "frames_small.npz"), **out)
    print("wrote frames_small.npz", os.path.getsize(os.path.join(HERE, "frames_small.npz")), "bytes")


if __name__ == "__main__":
    main()
