"""Test-side .rle4 encoder (mip 0 only): the inverse of RLE4::load (src/octree/Rle4.cpp:18-73,91-161).

Column stream, x fastest then z: [count][numtex][count slabs][numtex colours]; slab = skip:10 | run:6;
voxel y = sy-1-y1 with y1 ascending.  colour16 is chosen so that the loader with palette==0 gives back
rgba.x = 1 + ((255-(tcol&255)) & 0xfc)  ->  tcol&255 = 255 - (rgba.x - 1).
"""
import struct

import numpy as np


def encode_columns(x, y, z, rgba, sx, sy, sz):
    """-> uint16 slab stream.  Duplicate voxels in a column collapse (last colour wins)."""
    cols = {}
    for xi, yi, zi, ci in zip(x.tolist(), y.tolist(), z.tolist(), rgba.tolist()):
        cols.setdefault((zi, xi), {})[sy - 1 - yi] = ci
    out = []
    nvox = 0
    for zi in range(sz):
        for xi in range(sx):
            col = cols.get((zi, xi))
            if not col:
                out += [0, 0]
                continue
            ys = sorted(col)
            slabs, tex, cur = [], [], 0
            i = 0
            while i < len(ys):
                j = i
                while j + 1 < len(ys) and ys[j + 1] == ys[j] + 1 and j + 1 - i < 63:
                    j += 1
                skip = ys[i] - cur
                while skip > 1023:                      # long gap: emit empty slabs
                    slabs.append(1023); skip -= 1023
                slabs.append(skip | ((j - i + 1) << 10))
                for k in range(i, j + 1):
                    c = col[ys[k]]
                    tex.append((255 - ((c & 255) - 1)) & 255)
                cur = ys[j] + 1
                i = j + 1
            nvox += len(tex)
            out += [len(slabs), len(tex)] + slabs + tex
    return np.array(out, dtype=np.uint16), nvox


def write_rle4(path, x, y, z, rgba, sx, sy, sz):
    slabs, nvox = encode_columns(np.asarray(x), np.asarray(y), np.asarray(z), np.asarray(rgba), sx, sy, sz)
    with open(path, "wb") as f:
        f.write(struct.pack("<5i", 1, sx, sy, sz, len(slabs)))
        f.write(slabs.tobytes())
    return nvox
