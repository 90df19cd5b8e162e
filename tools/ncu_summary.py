#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep) into the per-kernel table kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [traffic.json] > profiles/<name>.md
With a second argument the per-kernel average of dram__bytes_read.sum + dram__bytes_write.sum per launch is merged into
that JSON file ({kernel: {"dram_bytes_per_launch": ...}}): bench.py reads profiles/ncu_traffic.json for roofline.traffic."""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_act%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act%"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit%"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_sel"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), s) for m, s in WANT if m in hdr]
    print(f"# ncu --set full summary of `{rep}` (one row per profiled launch; cold-cache replays, compare shares not absolutes)\n")
    print("| kernel | " + " | ".join(f"{s} [{units[i]}]" if units[i] else s for i, s in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        vals = []
        for i, _ in cols:
            v = r[i].replace(",", "")
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(v)
        print("| " + r[ki].split("(")[0].replace("void ", "") + " | " + " | ".join(vals) + " |")
    if len(sys.argv) > 2:
        import json, os
        UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        acc = {}
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        ii, il, ia = hdr.index("smsp__inst_executed.sum"), hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio"), hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")

        def num(x):
            return float(x.replace(",", ""))
        for r in rows[2:]:
            k = r[ki].split("(")[0].replace("void ", "").replace("svo::", "").split("<")[0]
            b = num(r[ir]) * UNIT.get(units[ir], 1.0) + num(r[iw]) * UNIT.get(units[iw], 1.0)
            acc.setdefault(k, []).append((b, num(r[it]), num(r[ii]), num(r[il]), num(r[ia])))
        out = json.load(open(sys.argv[2])) if os.path.exists(sys.argv[2]) else {}
        for k, v in acc.items():
            m = lambda j: sum(x[j] for x in v) / len(v)
            keep = {f: out[k][f] for f in ("rays_per_launch", "rays_note") if k in out and f in out[k]}   # annotations survive a refresh
            out[k] = {**keep, "dram_bytes_per_launch": m(0), "launches_profiled": len(v),
                      # issue roof of the traversal kernels (bench.py roofline): warp instructions issued, lanes active per instruction, issue slots busy
                      "warp_inst_per_launch": m(2), "thr_per_inst": m(3), "issue_active_pct": m(4),
                      "source": os.path.basename(rep) + " (ncu --set full, caches flushed before each replay)"}
        json.dump(out, open(sys.argv[2], "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
