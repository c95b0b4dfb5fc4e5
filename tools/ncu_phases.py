#!/usr/bin/env python
"""Warp-instruction share per source line of an .ncu-rep (--import-source on, -lineinfo), in file/line order, so that
the phases of the persistent kernels (vote loop, node step, leaf, shade sub-phases, refill, begin) can be summed.
    python tools/ncu_phases.py x.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.08
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, acc = "", None, {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[2] == "-":
        cur_line = (cur_file, int(r[0]), r[1].strip()); continue
    try:
        i_ins = hdr.index("Instructions Executed"); i_thr = hdr.index("Thread Instructions Executed"); i_s = hdr.index("# Samples")
        ins, th, smp = int(r[i_ins] or 0), int(r[i_thr] or 0), int(r[i_s] or 0)
    except (ValueError, IndexError):
        continue
    a = acc.setdefault(cur_line, [0, 0, 0]); a[0] += ins; a[1] += th; a[2] += smp
tot = sum(a[0] for a in acc.values()); tots = sum(a[2] for a in acc.values())
print(f"total warp instructions {tot}, samples {tots}, lanes/inst {sum(a[1] for a in acc.values())/max(tot,1):.2f}")
byfile = {}
for (f, l, s), a in acc.items():
    b = byfile.setdefault(f, [0, 0, 0]); b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
for f, b in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"  {f:<26} {100*b[0]/max(tot,1):5.1f}% inst {100*b[2]/max(tots,1):5.1f}% smp lanes {b[1]/max(b[0],1):5.1f}")
for (f, l, s), a in sorted(acc.items()):
    if 100 * a[0] / tot >= thr or 100 * a[2] / tots >= 2 * thr:
        print(f"{100*a[0]/tot:5.2f}% {100*a[2]/tots:5.2f}%s L{a[1]/max(a[0],1):4.1f} {f}:{l} {s[:100]}")
