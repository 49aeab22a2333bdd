#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt [units_per_launch unit_name]"""
import csv
import re
import subprocess
import sys

KEYS = [r"^gpu__time_duration\.sum$", r"^launch__(grid_size|block_size|registers_per_thread|occupancy_limit_\w+)$",
        r"^dram__bytes_(read|write)\.sum$", r"^lts__t_bytes\.sum$", r"^smsp__inst_executed\.sum$",
        r"^smsp__issue_active\.avg\.per_cycle_active$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$", r"^sm__pipe_(alu|fma)_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__thread_inst_executed_per_inst_executed\.ratio$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
        r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$", r"^sass__inst_executed_local_(loads|stores)$",
        r"^smsp__sass_average_branch_targets_threads_uniform\.pct$", r"^smsp__inst_executed_op_tma_ld\.sum$",
        r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^gpu__dram_throughput"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    uname = sys.argv[4] if len(sys.argv) > 4 else "unit"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, un = rows[0], rows[1]
    lines = [f"# ncu summary of {rep}", f"# command: ncu --set full --clock-control none --import-source on (see DESIGN.md §6)", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== kernel: {d.get('Kernel Name', '?')[:120]}")
        vals = {}
        for h, u, v in zip(hdr, un, r):
            if any(re.search(k, h) for k in KEYS):
                lines.append(f"{h:85s} {v:>18s} {u}")
                try:
                    vals[h] = float(v)
                except ValueError:
                    pass
        if units:
            inst = vals.get("smsp__inst_executed.sum")
            if inst:
                lines.append(f"derived: warp instructions per {uname}: {inst / units:.1f}")
            rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    cur, h2, agg = None, None, []
    for r in srows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r and r[0] == "Line No":
            h2 = r; continue
        if h2 is None or len(r) < 10:
            continue
        if r[2] == "-" and r[0].isdigit():
            d = dict(zip(h2[4:], r[4:]))
            try:
                agg.append((int(d["# Samples"]), int(d["Instructions Executed"]), cur, int(r[0]), r[1].strip()[:90]))
            except (KeyError, ValueError):
                pass
    ts, ti = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    lines.append("== top source lines by stall samples (pct samples, pct instructions)")
    for a in sorted(agg, reverse=True)[:25]:
        lines.append(f"{100 * a[0] / ts:5.1f}% smp {100 * a[1] / ti:5.1f}% inst  {a[2]}:{a[3]:<4d} {a[4]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
