#!/bin/bash
# conv path check: parity of the weight-stationary kernels, then the per-layer probe and the headline bench
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_conv_ws.py tests/test_gpu_network.py tests/test_gpu_stem.py -m gpu -x -q 2>&1 | tail -15
WHICH=net timeout 300 python tools/perf_probe.py > $O/perf_probe.txt 2>&1; tail -30 $O/perf_probe.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_resnet18.json 2> $O/bench_resnet18.err; echo "resnet18 rc=$?"; tail -3 $O/bench_resnet18.err
python -c "
import json;d=json.load(open('$O/bench_resnet18.json'));print(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('bit_exact'))"
