#!/bin/bash
# stage isolation of the weight-stationary convolution kernels and the stem (needs the -DACCEL_DEV=1 build):
#   dbg 0 = everything, 1 = no epilogue work, 2 = no MMAs, 3 = loads only
mkdir -p gpurun_out/r2
O=gpurun_out/r2/isolation.txt
: > $O
export ACCEL_B200_LIB=$PWD/tools/probe/libaccel_dev.so
for d in 0 1 2 3; do
  echo "== dbg=$d" >> $O
  ACCEL_DBG_FLAGS=$d LAYER=layer1.0.conv1,layer1.0.conv2,layer2.0.conv1,layer2.1.conv1,layer2.1.conv2,layer3.1.conv1,layer4.1.conv1 timeout 300 python tools/ws_probe.py >> $O 2>&1
done
for d in 0 1 2 3; do
  echo "== stem dbg=$d" >> $O
  ACCEL_DBG_FLAGS=$d timeout 300 python tools/stem_probe.py >> $O 2>&1
done
cat $O
