#!/bin/bash
mkdir -p gpurun_out/r2
for f in 3 1 0; do ACCEL_DBG_FLAGS=$f timeout 200 python tools/ws_timeline.py 2>&1 | tee -a gpurun_out/r2/ws_timeline.txt; done
