#!/bin/bash
# A/B on one box: per-layer probe, current build vs tools/probe/libaccel_head.so
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_conv_ws.py tests/test_gpu_network.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
WHICH=net timeout 300 python tools/perf_probe.py > $O/perf_cur_$i.txt 2>&1
ACCEL_B200_LIB=$PWD/tools/probe/libaccel_head.so WHICH=net timeout 300 python tools/perf_probe.py > $O/perf_head_$i.txt 2>&1
done
paste <(cut -c1-36 $O/perf_cur_1.txt) <(cut -c20-36 $O/perf_head_1.txt) <(cut -c20-36 $O/perf_cur_2.txt) <(cut -c20-36 $O/perf_head_2.txt)
