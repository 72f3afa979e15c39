#!/bin/bash
# per-layer probe of the ResNet-18 network at batch 256 / 128 / 64 on one box (wave quantisation and per-launch fixed cost)
mkdir -p gpurun_out/r2
for b in 256 128 64; do BATCH=$b WHICH=net timeout 300 python tools/perf_probe.py > gpurun_out/r2/pp_batch_$b.txt 2>&1; done
paste <(cut -c1-32 gpurun_out/r2/pp_batch_256.txt) <(cut -c20-32 gpurun_out/r2/pp_batch_128.txt) <(cut -c20-32 gpurun_out/r2/pp_batch_64.txt)
