#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export LAYER=layer1.0.conv1 BATCH=256 WHICH=conv REPS=2
ACCEL_DBG_FLAGS=3 python tools/ncu_target.py > $O/ncu_plain_l1_load.log 2>&1 && \
ACCEL_DBG_FLAGS=3 ncu --set full --clock-control none --import-source on -k regex:conv_ws_kernel -s 1 -c 1 -o $O/prof_l1_loadonly python tools/ncu_target.py > $O/ncu_l1_load.log 2>&1
python tools/ncu_target.py > $O/ncu_plain_l1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_ws_kernel -s 1 -c 1 -o $O/prof_l1_full python tools/ncu_target.py > $O/ncu_l1_full.log 2>&1
ls -la $O/*.ncu-rep
