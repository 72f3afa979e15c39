#!/bin/bash
# per-layer probe of library variants (ACCEL_B200_LIB) on ONE box, then the conv / stem / network parity tests on the default build
mkdir -p gpurun_out/r2
O=gpurun_out/r2
R=$PWD
VARS="${VARS:-nofence}"
for i in 1 2; do
  WHICH=net timeout 300 python tools/perf_probe.py > $O/pp_cur_$i.txt 2>&1
  for v in $VARS; do
    ACCEL_B200_LIB=$R/tools/probe/libaccel_$v.so WHICH=net timeout 300 python tools/perf_probe.py > $O/pp_${v}_$i.txt 2>&1
  done
done
for i in 1 2; do
echo "layer                    cur        $VARS"
eval "paste <(cut -c1-32 $O/pp_cur_$i.txt) $(for v in $VARS; do printf "<(cut -c20-32 $O/pp_%s_$i.txt) " $v; done)"
done

