#!/usr/bin/env python3
"""Summarise an .ncu-rep here on the CPU box: per-kernel headline metrics and the hottest SASS lines.

    python tools/ncu_top.py gpurun_out/prof.ncu-rep [--kernel regex] [--top 25] [--view sass|source]
"""
import argparse
import csv
import io
import subprocess
import sys

HEAD = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", default=None)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--view", default="sass")
    ap.add_argument("--no-source", action="store_true")
    a = ap.parse_args()
    raw = list(csv.reader(io.StringIO(run(["-i", a.rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    ki = hdr.index("Kernel Name")
    for r in raw[2:]:
        print("==", r[ki][:90])
        for m in HEAD:
            if m in hdr:
                print(f"   {m:70s} {r[hdr.index(m)]:>16s} {units[hdr.index(m)]}")
    if a.no_source:
        return
    cmd = ["-i", a.rep, "--page", "source", "--csv", "--print-source", a.view]
    if a.kernel:
        cmd += ["--kernel-name", "regex:" + a.kernel]
    out = run(cmd)
    # the page repeats (header, rows) per profiled launch: keep the first launch of each kernel
    seen = set()
    blocks = out.split('"Kernel Name",')
    for b in blocks[1:]:
        lines = b.splitlines()
        name = lines[0].strip('",')
        if name in seen:
            continue
        seen.add(name)
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        if not rows:
            continue
        h = rows[0]
        try:
            si, src, ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        except ValueError:
            continue
        stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        body = [r for r in rows[1:] if len(r) == len(h)]
        tot = sum(int(r[si] or 0) for r in body) or 1
        print(f"\n#### {name[:100]}   total samples {tot}, instructions {sum(int(r[ex] or 0) for r in body)}")
        agg = {h[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
        print("   stall mix:", ", ".join(f"{k[6:]}={v*100//tot}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        for r in sorted(body, key=lambda r: -int(r[si] or 0))[:a.top]:
            why = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
            print(f"   {int(r[si])*100/tot:5.1f}%  ex={r[ex]:>9s}  {r[src].strip()[:90]:90s} {why}")


if __name__ == "__main__":
    sys.exit(main())
