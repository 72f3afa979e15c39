#!/bin/bash
# ncu --set full of every hot kernel variant on the current build (one GPU).  Each capture runs only after the same
# command has exited 0 without ncu.  Reports land in gpurun_out/r2/ncu_*.ncu-rep; summarise with tools/ncu_top.py.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
cap() {   # cap <tag> <kernel regex> <skip> -- env... cmd
  local tag=$1 rx=$2 skip=$3; shift 3
  env "$@" > $O/plain_$tag.log 2>&1 || { echo "$tag: plain run failed"; tail -3 $O/plain_$tag.log; return; }
  env "$@" ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/ncu_$tag > $O/ncu_$tag.log 2>&1
  echo "$tag rc=$?"
}
ncu_py() { :; }
for L in layer1.0.conv1 layer1.0.conv2 layer2.0.conv1 layer2.1.conv2 layer3.1.conv1 layer4.1.conv1; do
  env LAYER=$L BATCH=256 WHICH=conv REPS=2 python tools/ncu_target.py > $O/plain_$L.log 2>&1 || { echo "$L plain failed"; continue; }
  env LAYER=$L BATCH=256 WHICH=conv REPS=2 ncu --set full --clock-control none --import-source on -k regex:conv_ws_kernel -s 1 -c 1 -f -o $O/ncu_$L python tools/ncu_target.py > $O/ncu_$L.log 2>&1; echo "$L rc=$?"
done
python tools/stem_probe.py > $O/plain_stem.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stem_ws_kernel -s 4 -c 1 -f -o $O/ncu_stem python tools/stem_probe.py > $O/ncu_stem.log 2>&1; echo "stem rc=$?"
env WHICH=gemm SPARSITY=70 REPS=3 python tools/ncu_target.py > $O/plain_gemm.log 2>&1 && \
env WHICH=gemm SPARSITY=70 REPS=3 ncu --set full --clock-control none --import-source on -k regex:gemm_ws_kernel -s 1 -c 1 -f -o $O/ncu_gemm4096 python tools/ncu_target.py > $O/ncu_gemm.log 2>&1; echo "gemm rc=$?"
# summaries here on the box (the reports are 16 MB each; gpurun brings back at most 64 MiB)
for r in $O/ncu_*.ncu-rep; do
  python tools/ncu_top.py $r --top 30 > ${r%.ncu-rep}.txt 2>&1
done
for r in $O/ncu_*.ncu-rep; do case $r in *layer1.0.conv1*|*stem*) ;; *) rm -f $r ;; esac; done
ls -la $O/ncu_*
