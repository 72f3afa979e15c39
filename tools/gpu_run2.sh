#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
run() { tag=$1; shift; env "$@" timeout 200 python tools/gemm_probe.py --tag $tag --no-check --pcts 0 --iters 10 2>&1 | grep '"n": 4096' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], 'flushed %.1f us  b2b %.1f us  dense-eq %.0f TOPS'%(d['us_flushed_median'], d['us_back_to_back'], d['dense_equiv_tops']))" | tee -a $O/p2.txt; }
echo "== correctness CG=2" | tee -a $O/p2.txt
ACCEL_GEMM_WS_CG=2 ACCEL_GEMM_WS_MIN_ROWS=1 timeout 600 python -m pytest tests/test_gpu_bsr_gemm.py -x -q 2>&1 | tail -1 | tee -a $O/p2.txt
ACCEL_GEMM_WS_CG=2 timeout 300 python tools/gemm_probe.py --tag cg2 --pcts 0,90 > $O/p2_cg2_check.txt 2>&1; cut -c1-330 $O/p2_cg2_check.txt | tee -a $O/p2.txt
run cg1 ACCEL_GEMM_WS_CG=1
run cg1_noepi ACCEL_GEMM_WS_CG=1 ACCEL_DBG_FLAGS=1
run cg1_nomma ACCEL_GEMM_WS_CG=1 ACCEL_DBG_FLAGS=2
run cg1_loadonly ACCEL_GEMM_WS_CG=1 ACCEL_DBG_FLAGS=3
run cg2 ACCEL_GEMM_WS_CG=2
run cg2_noepi ACCEL_GEMM_WS_CG=2 ACCEL_DBG_FLAGS=1
run cg2_nomma ACCEL_GEMM_WS_CG=2 ACCEL_DBG_FLAGS=2
run cg2_loadonly ACCEL_GEMM_WS_CG=2 ACCEL_DBG_FLAGS=3
run cg2_s3 ACCEL_GEMM_WS_CG=2 ACCEL_GEMM_WS_STAGES=3
run cg2_s5 ACCEL_GEMM_WS_CG=2 ACCEL_GEMM_WS_STAGES=5
