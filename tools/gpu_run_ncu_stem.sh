#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
ncu --set full --clock-control none --import-source on -k regex:stem_ws_kernel -s 4 -c 1 -f -o $O/prof_stem python tools/stem_probe.py > $O/ncu_stem.log 2>&1
ls -la $O/prof_stem*.ncu-rep
