#!/bin/bash
# multi-GPU evidence: ResNet-18 batch-sharded (no collective) and ResNet-50 with the block-row-sharded FC over NCCL
N=${1:-8}
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > $O/lscpu_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_resnet18_n$N.json 2> $O/bench_resnet18_n$N.err; echo "resnet18 n$N rc=$?"; tail -2 $O/bench_resnet18_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload resnet50_fc_sharded --steps 5 --warmup 3 --sustain-seconds 0.5 > $O/bench_resnet50_n$N.json 2> $O/bench_resnet50_n$N.err; echo "resnet50 n$N rc=$?"; tail -2 $O/bench_resnet50_n$N.err
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k nccl 2>&1 | tail -3 | tee $O/t_nccl_n$N.txt
