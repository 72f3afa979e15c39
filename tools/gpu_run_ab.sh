#!/bin/bash
# A/B on one box: per-layer probe, normal build vs the bulk-copy loader experiment (wrong results, timing only)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
for i in 1 2; do
WHICH=net timeout 300 python tools/perf_probe.py > $O/perf_base_$i.txt 2>&1
ACCEL_B200_LIB=$PWD/tools/probe/libaccel_bulk.so WHICH=net timeout 300 python tools/perf_probe.py > $O/perf_bulk_$i.txt 2>&1
done
paste <(cut -c1-36 $O/perf_base_1.txt) <(cut -c20-36 $O/perf_bulk_1.txt) <(cut -c20-36 $O/perf_base_2.txt) <(cut -c20-36 $O/perf_bulk_2.txt)
