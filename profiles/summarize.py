#!/usr/bin/env python
"""Turns an ncu report (.ncu-rep, brought back in gpurun_out/) into a small committed summary:
key raw metrics of the first profiled launch plus the hottest source lines.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/NAME.md ["title"]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_bytes.sum", "sm__tensor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    lines = ["# " + title, "", "source: `%s` (ncu --set full --clock-control none --import-source on)" % rep, ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines += ["## launch: `%s`" % name, "", "| metric | value | unit |", "|---|---|---|"]
        for i, h in enumerate(hdr):
            if h in KEYS:
                lines.append("| %s | %s | %s |" % (h, r[i], units[i]))
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur, agg = None, {}
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 60 or r[0] == "Line No" or r[2] != "-":
            continue
        try:
            agg[(cur, int(r[0]))] = (int(r[4]), int(r[7]), int(r[8]), r[1].strip()[:100])
        except ValueError:
            pass
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    lines += ["## hottest source lines (warp-stall samples, all profiled launches)", "",
              "| samples % | warp inst % | threads/inst | line | source |", "|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        lines.append("| %.1f | %.1f | %.1f | %s:%d | `%s` |" % (100 * a[0] / tot, 100 * a[1] / toti,
                                                             a[2] / max(a[1], 1), k[0], k[1], a[3].replace("|", "\\|")))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
