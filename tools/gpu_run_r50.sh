#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_conv_ws.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload resnet50_fc_sharded --steps 10 --warmup 3 --sustain-seconds 0.5 > $O/bench_resnet50_n1.json 2> $O/bench_resnet50_n1.err; echo "resnet50 rc=$?"; tail -3 $O/bench_resnet50_n1.err
python -c "
import json;d=json.load(open('$O/bench_resnet50_n1.json'));print(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('bit_exact'))
for k in d['roofline'].get('kernels',[])[:60]:
    if 'downsample' in k['name'] or 'layer1.0' in k['name']: print('   ',k['name'],round(k['us'],1))"
