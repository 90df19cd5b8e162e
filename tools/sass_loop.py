#!/usr/bin/env python
"""Instruction count of the traversal loop of a kernel in libsvo_b200.so (SASS between the first backward-branch target and the
backward branch): a CPU-side proxy for the issue-bound ray kernels.  usage: tools/sass_loop.py <mangled-name-substring>"""
import re, subprocess, sys
so = "sparse-voxel-octree-raycasting_b200/libsvo_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if sys.argv[1] not in name:
        continue
    back = [(a, int(re.search(r"BRA (0x[0-9a-f]+)", t).group(1), 16)) for a, t in ins if re.search(r"BRA 0x", t) and int(re.search(r"BRA (0x[0-9a-f]+)", t).group(1), 16) < a]
    print(name, "total", len(ins))
    for a, tgt in back:
        body = [t for ad, t in ins if tgt <= ad <= a]
        ctl = sum(1 for t in body if re.match(r"(@!?U?P\d+ )?(BSSY|BSYNC|BREAK|BRA|WARPSYNC)", t))
        print(f"  loop 0x{tgt:x}..0x{a:x}: {len(body)} instr, {ctl} control, {sum(1 for t in body if 'LDG' in t)} LDG, {sum(1 for t in body if 'MOV' in t)} MOV")
