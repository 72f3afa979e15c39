#!/bin/bash
# state of HEAD on one box: full GPU suite, headline bench, ncu launch list of the same command
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_resnet18.json 2> $O/bench_resnet18.err; echo "resnet18 rc=$?"; tail -2 $O/bench_resnet18.err
python -c "
import json;d=json.load(open('$O/bench_resnet18.json'));print(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('bit_exact'))
for k in d['roofline'].get('kernels',[])[:60]: print('   ',k['name'],round(k['us'],1))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustain-seconds 0 > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
