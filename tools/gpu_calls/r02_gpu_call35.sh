#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== mode 16, 32-column units: second chunk of warp 2 per tile: 0 prev store issued, 1 pre done, 2 tcgen05.ld landed, 3 next ld issued, 4 rows done, 5 st.shared done, 6 fence+sync, 7 store issued"
SDB_NO_EPI_W64=1 timeout 300 python tools/gemm_trace.py linear_qk_65536x320x640 --mode 16 2>&1 | tail -16
echo "== mode 16, fp32 out + fp32 residual (linear_proj)"
timeout 300 python tools/gemm_trace.py linear_proj_65536x320x320 --mode 16 2>&1 | tail -16
} > gpurun_out/r02af_gemm_trace_m16.log 2>&1
cat gpurun_out/r02af_gemm_trace_m16.log
