#!/bin/bash
# full GPU suite + headline bench + ResNet-50 workload (one box)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_resnet18.json 2> $O/bench_resnet18.err; echo "resnet18 rc=$?"; tail -2 $O/bench_resnet18.err
python -c "
import json;d=json.load(open('$O/bench_resnet18.json'));print(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('bit_exact'))"
timeout 600 python bench.py --workload resnet50_fc_sharded --steps 10 --warmup 3 --sustain-seconds 0.5 > $O/bench_resnet50_n1.json 2> $O/bench_resnet50_n1.err; echo "resnet50 rc=$?"; tail -3 $O/bench_resnet50_n1.err
python -c "
import json;d=json.load(open('$O/bench_resnet50_n1.json'));print(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('bit_exact'))
for k in d['roofline'].get('kernels',[])[:60]: print('   ',k['name'],round(k['us'],1))"
