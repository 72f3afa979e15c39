#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
LAYERS=layer1.0.conv1,layer2.1.conv1,layer3.1.conv1
for f in 3 35 67 99 0 32 96; do
  echo "== dbg=$f" | tee -a $O/ws_iso8.txt
  GRAPH=1 LAYER=$LAYERS ACCEL_DBG_FLAGS=$f timeout 300 python tools/ws_probe.py 2>&1 | tail -3 | tee -a $O/ws_iso8.txt
done
