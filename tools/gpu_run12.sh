#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
LAYERS=layer1.0.conv1,layer1.0.conv2,layer2.1.conv1,layer3.1.conv1,layer4.1.conv1
for f in 0 1 3; do
  echo "== dbg=$f" | tee -a $O/ws_iso5.txt
  GRAPH=1 LAYER=$LAYERS ACCEL_DBG_FLAGS=$f timeout 300 python tools/ws_probe.py 2>&1 | tail -5 | tee -a $O/ws_iso5.txt
done
