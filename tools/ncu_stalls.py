#!/usr/bin/env python
"""Per-instruction warp-stall samples of one kernel from an ncu report (--set full --import-source on):
tools/ncu_stalls.py report.ncu-rep [min_samples]  ->  address, SASS, samples, executed, dominant stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
print(lines[0][:200])
rd = csv.DictReader(io.StringIO("\n".join(lines[hdr_i:])))
rows = list(rd)
tot = sum(int(r["# Samples"] or 0) for r in rows)
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: 0 for c in stall_cols}
base = int(rows[0]["Address"], 16)
for r in rows:
    s = int(r["# Samples"] or 0)
    for c in stall_cols:
        agg[c] += int(r[c] or 0)
    if s >= mins:
        top = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f'{int(r["Address"],16)-base:5x} {r["Source"].strip():58s} smp {s:6d} {100.0*s/tot:5.1f}%  exec {r["Instructions Executed"]:>8s}  ' + " ".join(f"{n}:{v}" for v, n in top if v))
print("total samples", tot)
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
