#!/bin/bash
# Final evidence of a build on ONE box: full GPU suite, smoke, every bench workload + the reference arms, ncu launch list.
mkdir -p gpurun_out/r2/final
O=gpurun_out/r2/final
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tail -7
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_resnet18.json 2> $O/bench_resnet18.err; echo "resnet18 rc=$?"; tail -2 $O/bench_resnet18.err
timeout 600 python bench.py --steps 20 --warmup 5 --sparsity 90 --no-cpu-baseline > $O/bench_resnet18_s90.json 2> $O/bench_resnet18_s90.err; echo "resnet18@90 rc=$?"
for s in 0 50 70 90; do timeout 300 python bench.py --workload gemm4096 --sparsity $s --steps 20 --warmup 5 $( [ $s != 70 ] && echo --no-cpu-baseline ) > $O/bench_gemm4096_s$s.json 2> $O/bench_gemm4096_s$s.err; echo "gemm4096 $s rc=$?"; done
timeout 300 python bench.py --workload mnist --steps 20 --warmup 5 > $O/bench_mnist.json 2> $O/bench_mnist.err; echo "mnist rc=$?"
timeout 600 python bench.py --workload resnet50_fc_sharded --steps 10 --warmup 3 --sustain-seconds 0.5 > $O/bench_resnet50_n1.json 2> $O/bench_resnet50_n1.err; echo "resnet50 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $O/bench_ref_resnet18.json 2>&1; echo "ref resnet18 rc=$?"
timeout 300 python bench.py --impl reference --workload gemm4096 --steps 2 --warmup 0 > $O/bench_ref_gemm4096.json 2>&1; echo "ref gemm rc=$?"
timeout 300 python bench.py --impl reference --workload mnist --steps 2 --warmup 0 > $O/bench_ref_mnist.json 2>&1; echo "ref mnist rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2/final/bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r = d.get('roofline', {})
    print(f.split('/')[-1], round(d.get('value', 0), 1), d.get('unit'), 'ms', round(d.get('ms_per_step', 0), 4), 'e2e', round((d.get('e2e') or {}).get('value', 0), 1),
          'frac', r.get('frac'), 'bit_exact', (d.get('bit_exact') or {}).get('ok'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"conv_ws_kernel|stem_ws_kernel|gemm_ws_kernel|bsr_tc|avgpool|maxpool" -c 200 --csv --log-file $O/launches.csv \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline --sustain-seconds 0 > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
