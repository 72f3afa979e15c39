#!/usr/bin/env python3
"""Turn an ncu launch list (--csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over
`bench.py --steps S --warmup W`) into profiles/<tag>_launches.md (per-kernel shares) and profiles/<tag>_traffic.json
(DRAM bytes of the conv / fc launches of ONE step, what bench.py reports as roofline.traffic).

    python tools/ncu_launches.py gpurun_out/launches.csv --tag r01 --steps-in-capture N
"""
import argparse
import collections
import csv
import json
import os
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--tag", default="r01")
ap.add_argument("--forwards", type=int, default=0, help="network forward passes inside the capture (default: count the avgpool launches)")
ap.add_argument("--note", default="")
a = ap.parse_args()

rows = []
with open(a.csv, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
per = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rd:
    name = r.get("Kernel Name", "")
    m, v = r.get("Metric Name"), r.get("Metric Value", "0").replace(",", "")
    unit = r.get("Metric Unit", "")
    try:
        val = float(v)
    except ValueError:
        continue
    if m == "gpu__time_duration.sum":
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)        # -> us
    if m and m.startswith("dram__bytes"):
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per[name][m] += val
    if m == "gpu__time_duration.sum":
        per[name]["launches"] += 1
tot = sum(d["gpu__time_duration.sum"] for d in per.values()) or 1.0
if not a.forwards:
    a.forwards = int(sum(d["launches"] for n, d in per.items() if "avgpool_i8" in n)) or 1
ours = re.compile(r"conv_ws_kernel|stem_ws_kernel|gemm_ws_kernel|bsr_tcp_kernel|bsr_tc_kernel")
os.makedirs("profiles", exist_ok=True)
with open(f"profiles/{a.tag}_launches.md", "w") as f:
    f.write(f"# ncu launch list ({a.tag}): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
            f"--clock-control none` over `bench.py` ({a.forwards} forward passes in the capture){' - ' + a.note if a.note else ''}\n\n"
            "Cold-cache, serialised per-launch times: compare SHARES, not absolutes.\n\n"
            "| kernel | launches | total us | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|\n")
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        f.write(f"| `{name[:70]}` | {int(d['launches'])} | {d['gpu__time_duration.sum']:.1f} | "
                f"{100 * d['gpu__time_duration.sum'] / tot:.1f} % | {d['dram__bytes_read.sum'] / 1e6:.1f} | "
                f"{d['dram__bytes_write.sum'] / 1e6:.1f} |\n")
conv_bytes = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for n, d in per.items() if ours.search(n))
conv_us = sum(d["gpu__time_duration.sum"] for n, d in per.items() if ours.search(n))
out = {"conv_fc_dram_bytes_per_step": conv_bytes / a.forwards, "conv_fc_us_per_step_under_ncu": conv_us / a.forwards,
       "source": f"profiles/{a.tag}_launches.csv: sum of dram__bytes_read + dram__bytes_write over the conv_ws / stem_ws / bsr_tcp / bsr_tc "
                 f"launches, divided by the {a.forwards} forward passes of the capture"}
with open(f"profiles/{a.tag}_traffic.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
