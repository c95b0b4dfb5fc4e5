#!/usr/bin/env python
"""gpurun_out/ncu/*.ncu-rep (tools/ncu_capture.sh) -> profiles/r02_<workload>_summary.txt + profiles/ncu_counters.json.

ncu_counters.json carries, per bench workload, the counters bench.py quotes next to the SURVEY 8d roofline — DRAM bytes of
the captured launch, issue-slot utilisation, lanes per instruction, L1 data-pipe utilisation, long-scoreboard stalls —
stamped with the hash of the kernel sources the captured library was built from and the git commit; bench.py refuses
the file when that hash is not the hash of the sources it runs (bench.kernel_source_hash)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCU = os.path.join(ROOT, "gpurun_out", "ncu")
WANT = {"gpu__time_duration.sum": "kernel_ms", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1tex_data_pipe_pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "long_scoreboard",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "launch__registers_per_thread": "registers", "smsp__inst_executed.sum": "warp_instructions",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[head.index("Kernel Name")]}
    for k, name in WANT.items():
        if k in head:
            i = head.index(k)
            v = float(vals[i].replace(",", ""))
            d[name] = v * UNIT.get(units[i], 1.0) if name in ("dram_read", "dram_write", "kernel_ms") else v
    return d


def main():
    with open(os.path.join(NCU, "kernel_source_hash.txt")) as f:
        khash = f.read().strip()
    captured = {}
    with open(os.path.join(NCU, "captured.txt")) as f:
        for line in f:
            name, as_ = line.split()
            captured[name] = as_.split("=")[1]
    git = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"kernel_source_hash": khash, "git_head_when_summarised": git,
           "how": "tools/ncu_capture.sh on the B200 (ncu --set full --clock-control none --import-source on, 4th launch of the "
                  "kernel), summarised by tools/ncu_counters.py", "workloads": {}}
    for name, as_ in captured.items():
        rep = os.path.join(NCU, f"r02_{name}.ncu-rep")
        if not os.path.exists(rep):
            continue
        d = raw(rep)
        summ = f"profiles/r02_{name}_summary.txt"
        with open(os.path.join(ROOT, summ), "w") as f:
            f.write(f"# ncu --set full of the 4th launch in `bench.py --workload {as_} --only`; library built from kernel sources {khash} "
                    f"(git {git}); tools/ncu_capture.sh + tools/ncu_summary.py\n")
            f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
            f.write("\n# ---- warp-instruction share per source line (tools/ncu_phases.py)\n")
            f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), rep, "0.5"], capture_output=True, text=True).stdout)
        d["dram_bytes_per_launch"] = d.pop("dram_read", 0.0) + d.pop("dram_write", 0.0)
        d["captured_as"] = as_
        d["summary"] = summ
        out["workloads"][name] = d
        print(name, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})
    with open(os.path.join(ROOT, "profiles", "ncu_counters.json"), "w") as f:
        json.dump(out, f, indent=1)
    lst = os.path.join(NCU, "r02_launches_default_bench.csv")
    if os.path.exists(lst):
        import shutil
        shutil.copy(lst, os.path.join(ROOT, "profiles", "r02_launches_default_bench.csv"))


if __name__ == "__main__":
    main()
