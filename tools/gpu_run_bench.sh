#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_resnet18.json 2> $O/bench_resnet18.err; echo "resnet18 rc=$?"; tail -3 $O/bench_resnet18.err
for s in 0 50 70 90; do timeout 300 python bench.py --workload gemm4096 --sparsity $s --steps 20 --warmup 5 > $O/bench_gemm4096_s$s.json 2> $O/bench_gemm4096_s$s.err; echo "gemm4096 $s rc=$?"; tail -2 $O/bench_gemm4096_s$s.err; done
timeout 300 python bench.py --workload mnist --steps 20 --warmup 5 > $O/bench_mnist.json 2> $O/bench_mnist.err; echo "mnist rc=$?"; tail -3 $O/bench_mnist.err
timeout 600 python bench.py --workload resnet50_fc_sharded --steps 5 --warmup 3 --sustain-seconds 0.5 > $O/bench_resnet50_n1.json 2> $O/bench_resnet50_n1.err; echo "resnet50 rc=$?"; tail -3 $O/bench_resnet50_n1.err
timeout 300 python bench.py --impl reference --workload gemm4096 --steps 2 --warmup 0 > $O/bench_ref_gemm4096.json 2>&1; echo "ref gemm rc=$?"
timeout 300 python bench.py --impl reference --workload mnist --steps 2 --warmup 0 > $O/bench_ref_mnist.json 2>&1; echo "ref mnist rc=$?"
