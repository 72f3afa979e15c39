#!/bin/bash
# full GPU suite + headline bench (one box)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_resnet18.json 2> $O/bench_resnet18.err; echo "resnet18 rc=$?"; tail -2 $O/bench_resnet18.err
python -c "
import json;d=json.load(open('$O/bench_resnet18.json'));print(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('bit_exact'))"
