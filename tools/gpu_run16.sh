#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
LAYERS=layer1.0.conv1,layer1.0.conv2,layer2.1.conv2
for f in 2 258 514 770; do
  echo "== dbg=$f" | tee -a $O/ws_iso11.txt
  GRAPH=1 LAYER=$LAYERS ACCEL_DBG_FLAGS=$f timeout 300 python tools/ws_probe.py 2>&1 | tail -3 | tee -a $O/ws_iso11.txt
done
