#!/usr/bin/env python
"""Per-kernel averages of an ncu --csv metric log (gpu__time_duration, dram__bytes_read/write, lts__t_bytes).
usage: python tools/dram_summary.py gpurun_out/<tag>_dram_inpipeline.csv [out.json] > profiles/<name>.md
The JSON ({kernel: {"dram_bytes_per_launch", "l2_bytes_per_launch", "time_us", "launches"}}) is what bench.py reads for
roofline.traffic (profiles/ncu_traffic.json)."""
import collections
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}


def main():
    rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("svo::", "")
        k = k.split("<")[0]
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        acc[k][r["Metric Name"]].append(v)
    out = {}
    print(f"# per-kernel averages of `{sys.argv[1]}`\n")
    print("| kernel | launches | time us | DRAM read MB | DRAM write MB | L2 traffic MB |")
    print("|---|---|---|---|---|---|")
    for k, m in sorted(acc.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
        n = len(m.get("gpu__time_duration.sum", [])) or 1
        avg = lambda name: sum(m.get(name, [0])) / max(1, len(m.get(name, [0])))
        rd, wr, l2, t = avg("dram__bytes_read.sum"), avg("dram__bytes_write.sum"), avg("lts__t_bytes.sum"), avg("gpu__time_duration.sum")
        out[k] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "l2_bytes_per_launch": l2, "time_us": t, "launches": n}
        print(f"| {k} | {n} | {t:.1f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | {l2 / 1e6:.1f} |")
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
