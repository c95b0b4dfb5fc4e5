#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full --import-source on) into the text summaries kept under profiles/:
headline counters from the raw page + SASS opcode mix and the top stall instructions from the source page.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__maximum_warps_per_active_cycle_pct",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    head, units, rows = raw[0], raw[1], raw[2:]
    for r in rows:
        d = dict(zip(head, r))
        print("kernel", d.get("Kernel Name", "?"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for m in METRICS:
            if m in d:
                print(f"{m:<78} {units[head.index(m)]:<10} {d[m]}")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    if src and src[0] and src[0][0] == "Kernel Name":
        src = src[1:]
    if len(src) < 2:
        return
    h = src[0]
    ci = {n: i for i, n in enumerate(h)}
    col_src = ci.get("Source")
    col_inst = ci.get("# Instructions Executed") or ci.get("Instructions Executed")
    col_samp = ci.get("# Warp Stall Sampling (All Samples)") or ci.get("Warp Stall Sampling (All Samples)")
    col_thr = ci.get("Avg. Threads Executed")
    col_tag = ci.get("L1 Tag Requests Global")
    if col_src is None or col_inst is None or col_samp is None:
        print("source page columns:", h)
        return
    tot_i = tot_s = 0
    mix_i, mix_s = collections.Counter(), collections.Counter()
    lines = []
    for r in src[1:]:
        try:
            ins, smp = int(r[col_inst] or 0), int(r[col_samp] or 0)
        except ValueError:
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[col_src].strip()).split(" ")[0].split(".")[0]
        tot_i += ins; tot_s += smp
        mix_i[op] += ins; mix_s[op] += smp
        lines.append((smp, ins, r[col_src].strip(), r[col_thr] if col_thr is not None else "", r[col_tag] if col_tag is not None else ""))
    print(f"total warp inst {tot_i} samples {tot_s} n sass {len(lines)}")
    for op, n in mix_i.most_common(24):
        print(f"{op:<10} inst {100*n/max(tot_i,1):5.1f}%  stall-samples {100*mix_s[op]/max(tot_s,1):5.1f}%")
    print("top stall instructions:")
    for smp, ins, text, thr, tag in sorted(lines, reverse=True)[:16]:
        print(f"  {100*smp/max(tot_s,1):5.2f}% inst={ins:>12} thr={thr:>5} {text[:90]}")
    print("global loads/stores (L1 tag requests):")
    for smp, ins, text, thr, tag in lines:
        if tag not in ("", "0"):
            print(f"  inst={ins:>12} thr={thr:>5} tags={tag:>12} {text[:90]}")


if __name__ == "__main__":
    main()
