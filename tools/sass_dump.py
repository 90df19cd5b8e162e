#!/usr/bin/env python
"""SASS of one address range of a kernel in libsvo_b200.so: tools/sass_dump.py <name-substring> [lo hi] [so]"""
import re, subprocess, sys
so = sys.argv[4] if len(sys.argv) > 4 else "sparse-voxel-octree-raycasting_b200/libsvo_b200.so"
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur and sys.argv[1] in cur and lo <= int(m.group(1), 16) <= hi:
        print(m.group(1), m.group(2).strip())
