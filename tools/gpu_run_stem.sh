#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_stem.py -m gpu -x -q 2>&1 | tail -3
for f in 0 1 2 3; do
    echo "dbg=$f $(ACCEL_DBG_FLAGS=$f timeout 120 python tools/stem_probe.py 2>&1| head -1)"
done 2>&1 | tee $O/stem_ab.txt
timeout 120 python tools/stem_timeline.py > $O/stem_tl0.txt 2>&1
ACCEL_DBG_FLAGS=3 timeout 120 python tools/stem_timeline.py > $O/stem_tl3.txt 2>&1
sed -n 2,8p $O/stem_tl0.txt | cut -c90-230
sed -n 2,6p $O/stem_tl3.txt | cut -c1-230
