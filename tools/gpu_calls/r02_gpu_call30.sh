#!/bin/bash
# round 2, call 30: 64-column units in the TMA epilogue for 16-bit results
set -u
mkdir -p gpurun_out
T=r02aa
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "tma_epilogue or linear or f16" > gpurun_out/${T}_tests_k.log 2>&1; tail -5 gpurun_out/${T}_tests_k.log | cut -c1-250
timeout 600 python tools/epi16_probe.py > gpurun_out/${T}_epi16_probe.log 2>&1
grep -E "N=  320|N=  640|N= 1280" gpurun_out/${T}_epi16_probe.log | grep -v "fp32\|epi_mode\|bn=64 "
echo "== SDB_NO_EPI_W64=1"
SDB_NO_EPI_W64=1 timeout 600 python tools/epi16_probe.py > gpurun_out/${T}_epi16_probe_now64.log 2>&1
grep -E "N=  320|N=  640|N= 1280" gpurun_out/${T}_epi16_probe_now64.log | grep -v "fp32\|epi_mode\|bn=64 "
