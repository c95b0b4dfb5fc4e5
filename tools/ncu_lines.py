#!/usr/bin/env python
"""Per-source-line instruction counts from an .ncu-rep captured with --import-source on (-lineinfo build):
    python tools/ncu_lines.py x.ncu-rep [N]   -> the N hottest source lines (warp instructions, samples, lanes)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, acc = "", None, {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[2] == "-":  # a source line header row: aggregate rows follow with Address set
        cur_line = (cur_file, r[0], r[1].strip()); continue
    try:
        i_ins = hdr.index("Instructions Executed"); i_thr = hdr.index("Thread Instructions Executed"); i_s = hdr.index("# Samples")
        ins, thr, smp = int(r[i_ins] or 0), int(r[i_thr] or 0), int(r[i_s] or 0)
    except (ValueError, IndexError):
        continue
    a = acc.setdefault(cur_line, [0, 0, 0]); a[0] += ins; a[1] += thr; a[2] += smp
tot = sum(a[0] for a in acc.values()); tots = sum(a[2] for a in acc.values())
print(f"total warp instructions {tot}, samples {tots}")
byfile = {}
for (f, l, s), a in acc.items(): byfile[f] = byfile.get(f, 0) + a[0]
for f, n in sorted(byfile.items(), key=lambda x: -x[1]): print(f"  {f:<18} {100*n/max(tot,1):5.1f}% of instructions")
for (f, l, s), a in sorted(acc.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100*a[0]/max(tot,1):5.2f}% inst {100*a[2]/max(tots,1):5.2f}% smp lanes {a[1]/max(a[0],1):5.1f}  {f}:{l}  {s[:100]}")
