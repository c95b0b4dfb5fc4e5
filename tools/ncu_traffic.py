#!/usr/bin/env python
"""profiles/traffic.json: DRAM bytes per launch of the dominant kernel of each bench workload, from ncu --set full
captures (dram__bytes_read.sum + dram__bytes_write.sum).  bench.py reads it for roofline.traffic.

    python tools/ncu_traffic.py workload:kernel=report.ncu-rep [...]
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[m].replace(",", "")) * UNIT[units[head.index(m)]]
        out.append((d.get("Kernel Name", "?"), tot, float(d["gpu__time_duration.sum"].replace(",", "")), units[head.index("gpu__time_duration.sum")]))
    return out


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        launches = dram_bytes(rep)
        b = sum(x[1] for x in launches) / len(launches)
        data[key] = {"dram_bytes_per_launch": b, "launches_captured": len(launches), "kernel": launches[0][0][:80],
                     "duration_under_ncu": f"{launches[0][2]} {launches[0][3]}", "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_summary.txt")}
        print(key, f"{b/1e6:.1f} MB per launch")
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
