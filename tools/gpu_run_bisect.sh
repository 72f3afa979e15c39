#!/bin/bash
# per-layer probe of library variants (ACCEL_B200_LIB) and of an earlier commit (worktree) on ONE box
mkdir -p gpurun_out/r2
O=gpurun_out/r2
R=$PWD
for i in 1 2; do
  WHICH=net timeout 300 python tools/perf_probe.py > $O/pp_cur_$i.txt 2>&1
  for v in head fence dev; do
    ACCEL_B200_LIB=$R/tools/probe/libaccel_$v.so WHICH=net timeout 300 python tools/perf_probe.py > $O/pp_${v}_$i.txt 2>&1
  done
  (cd tools/probe/wt/99cad64 && WHICH=net timeout 300 python tools/perf_probe.py > $R/$O/pp_99cad64_$i.txt 2>&1)
done
for i in 1 2; do
echo "layer                    cur        head       fence       dev       99cad64"
paste <(cut -c1-32 $O/pp_cur_$i.txt) <(cut -c20-32 $O/pp_head_$i.txt) <(cut -c20-32 $O/pp_fence_$i.txt) <(cut -c20-32 $O/pp_dev_$i.txt) <(cut -c20-32 $O/pp_99cad64_$i.txt)
done
timeout 600 python bench.py --workload resnet50_fc_sharded --steps 10 --warmup 3 --sustain-seconds 0.5 --no-cpu-baseline > $O/bench_resnet50_n1.json 2> $O/bench_resnet50_n1.err; echo "resnet50 rc=$?"; tail -3 $O/bench_resnet50_n1.err
python -c "
import json;d=json.load(open('$O/bench_resnet50_n1.json'));print(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('bit_exact'))
for k in sorted(d['roofline'].get('kernels',[]),key=lambda k:-k['us'])[:30]: print('   ',k['name'],round(k['us'],1))"
