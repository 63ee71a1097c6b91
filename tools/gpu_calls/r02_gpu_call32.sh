#!/bin/bash
# round 2, call 32: batched shared loads in the TMA epilogue rows + 64-column units
set -u
mkdir -p gpurun_out
T=r02ac
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "tma_epilogue or linear or f16 or gn_partials" > gpurun_out/${T}_tests_k.log 2>&1; tail -5 gpurun_out/${T}_tests_k.log | cut -c1-250
timeout 600 python tools/epi16_probe.py > gpurun_out/${T}_epi16_probe.log 2>&1
grep -E "N=  320|N=  640|N= 1280" gpurun_out/${T}_epi16_probe.log | grep -v "epi_mode\|bn=64 "
echo "== SDB_NO_EPI_W64=1"
SDB_NO_EPI_W64=1 timeout 600 python tools/epi16_probe.py > gpurun_out/${T}_epi16_probe_now64.log 2>&1
grep -E "N=  320|N=  640|N= 1280" gpurun_out/${T}_epi16_probe_now64.log | grep -v "epi_mode\|bn=64 "
echo "== trace mode 8"; timeout 300 python tools/gemm_trace.py linear_qk_65536x320x640 --mode 8 2>&1 | tail -8
