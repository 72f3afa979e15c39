#!/bin/bash
# round-2 GPU call 1: new dense-equivalent GEMM (CG=1, CG=2) parity + timing, INT8 peak, baseline tests
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
echo "== tests (default routing)"; timeout 600 python -m pytest tests/test_gpu_bsr_gemm.py tests/test_gpu_api.py -x -q > $O/t_default.txt 2>&1; echo rc=$? >> $O/t_default.txt; tail -3 $O/t_default.txt
echo "== tests CG=1 all rows"; ACCEL_GEMM_WS_CG=1 ACCEL_GEMM_WS_MIN_ROWS=1 timeout 600 python -m pytest tests/test_gpu_bsr_gemm.py -x -q > $O/t_cg1.txt 2>&1; echo rc=$? >> $O/t_cg1.txt; tail -3 $O/t_cg1.txt
echo "== tests CG=2 all rows"; ACCEL_GEMM_WS_CG=2 ACCEL_GEMM_WS_MIN_ROWS=1 timeout 600 python -m pytest tests/test_gpu_bsr_gemm.py -x -q > $O/t_cg2.txt 2>&1; echo rc=$? >> $O/t_cg2.txt; tail -3 $O/t_cg2.txt
echo "== probe old kernel"; ACCEL_NO_GEMM_WS=1 timeout 300 python tools/gemm_probe.py --tag old --no-check > $O/p_old.txt 2>&1; tail -8 $O/p_old.txt
echo "== probe CG=1"; ACCEL_GEMM_WS_CG=1 timeout 300 python tools/gemm_probe.py --tag cg1 > $O/p_cg1.txt 2>&1; tail -8 $O/p_cg1.txt
echo "== probe CG=2"; ACCEL_GEMM_WS_CG=2 timeout 300 python tools/gemm_probe.py --tag cg2 > $O/p_cg2.txt 2>&1; tail -8 $O/p_cg2.txt
echo "== probe CG=2 no-mma / no-epi"; ACCEL_DBG_FLAGS=2 timeout 300 python tools/gemm_probe.py --tag cg2_nomma --no-check --pcts 0 > $O/p_cg2_nomma.txt 2>&1; tail -2 $O/p_cg2_nomma.txt
echo "== int8 peak"; timeout 200 python tools/int8_peak.py --out $O/int8_peak.json > $O/int8_peak.txt 2>&1; tail -2 $O/int8_peak.txt
echo "== full gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q > $O/t_all.txt 2>&1; echo rc=$? >> $O/t_all.txt; tail -3 $O/t_all.txt
