#!/usr/bin/env python
"""Turn `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` of one
core step into the per-kernel traffic table (markdown) and the JSON bench.py reads."""
import collections
import csv
import json
import re
import sys

src, grid, md_out, json_out, tag = sys.argv[1:6]
nx, ny, Nz = (int(x) for x in grid.split(","))
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("lg::", "").strip()
    m, u, v = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    a = agg.setdefault(name, {"n": 0, "rd": 0.0, "wr": 0.0, "ns": 0.0})
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u)
    if m == "dram__bytes_read.sum":
        a["rd"] += v * scale
    elif m == "dram__bytes_write.sum":
        a["wr"] += v * scale
    elif m == "gpu__time_duration.sum":
        a["ns"] += v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(u, 1.0)
        a["n"] += 1
tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
tot_ms = sum(a["ns"] for a in agg.values()) / 1e6
pts = nx * ny * Nz
with open(md_out, "w") as f:
    f.write(f"# DRAM traffic and launch list of ONE core step at {nx}x{ny}x{Nz} on one B200 ({tag})\n\n")
    f.write("`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
            f"--profile-from-start off python tools/profile_step.py --grid {grid}` (times under ncu are serialised/cold: use the shares).\n\n")
    f.write("| kernel | launches | ms (ncu) | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---|---|---|---|---|---|\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        ms = a["ns"] / 1e6
        f.write(f"| `{name}` | {a['n']} | {ms:.3f} | {100 * ms / tot_ms:.1f}% | {a['rd'] / 1e9:.2f} | {a['wr'] / 1e9:.2f} | "
                f"{(a['rd'] + a['wr']) / 1e9 / (ms / 1e3):.0f} |\n")
    f.write(f"\n**Total: {tot_ms:.2f} ms, {tot_b / 1e9:.1f} GB DRAM traffic per step = {tot_b / pts:.0f} B/point against the "
            f"algorithmic 296 B/point: overhead factor {tot_b / pts / 296:.2f}.**\n")
json.dump({"grid": [nx, ny, Nz], "dram_bytes_per_step": tot_b, "bytes_per_point": tot_b / pts, "ncu_ms": tot_ms,
           "source": f"{md_out} (ncu dram__bytes_read.sum + dram__bytes_write.sum over every launch of one lesgo_gpu_step)"},
          open(json_out, "w"), indent=1)
print(f"{tot_ms:.2f} ms, {tot_b / 1e9:.1f} GB, {tot_b / pts:.0f} B/point")
