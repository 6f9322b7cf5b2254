#!/usr/bin/env python
"""Writes profiles/traffic.json (DRAM read + write bytes per launch of the dominant kernels) from the raw-page CSV exports
of `ncu --set full` captures of `bench.py --steps 1 --warmup 1`.  usage: ncu_traffic.py <raw.csv> [<raw.csv> ...]"""
import csv, json, re, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
out = {"source": "ncu --set full --clock-control none, tools/one_chr1.py (chr1-sized synthetic local pair, second launch of each kernel); dram__bytes_read.sum + dram__bytes_write.sum", "kernels": {}}
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    kn, rd, wr, du = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    for r in rows[2:]:
        name = re.sub(r"^(void )?(sccg::)?", "", r[kn]).split("(")[0].split("<")[0]
        b = float(r[rd]) * unit[units[rd]] + float(r[wr]) * unit[units[wr]]
        e = out["kernels"].setdefault(name, {"dram_bytes_per_launch": 0, "launches": 0, "gpu_time_us_under_ncu": 0.0, "capture": Path(f).name})
        e["dram_bytes_per_launch"] += b; e["launches"] += 1; e["gpu_time_us_under_ncu"] += float(r[du])
for e in out["kernels"].values():
    e["dram_bytes_per_launch"] = int(e["dram_bytes_per_launch"] / e["launches"]); e["gpu_time_us_under_ncu"] = round(e["gpu_time_us_under_ncu"] / e["launches"], 1)
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
