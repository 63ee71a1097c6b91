#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== mode 1"; SDB_NO_EPI_W64=1 timeout 300 python tools/gemm_trace.py linear_qk_65536x320x640 --mode 1 2>&1 | tail -18
echo "== mode 8 (epilogue stamps: 0 before acc wait, 1 acc ready, 2 first chunk post start, 3 residual landed, 6 rows done, 7 fence+sync done, 4 store issued, 5 tile end)"; SDB_NO_EPI_W64=1 timeout 300 python tools/gemm_trace.py linear_qk_65536x320x640 --mode 8 2>&1 | tail -18
} > gpurun_out/r02ab_gemm_trace_qk.log 2>&1
cat gpurun_out/r02ab_gemm_trace_qk.log
