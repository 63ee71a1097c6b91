#!/bin/bash
# round 2, call 36: single-buffer TMA epilogue loop (instruction footprint)
set -u
mkdir -p gpurun_out
T=r02ag
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "tma_epilogue or linear or f16 or gn_partials" > gpurun_out/${T}_tests_k.log 2>&1; tail -5 gpurun_out/${T}_tests_k.log | cut -c1-250
timeout 600 python tools/epi16_probe.py > gpurun_out/${T}_epi16_probe.log 2>&1
grep -E "N=  320|N=  640|N= 1280" gpurun_out/${T}_epi16_probe.log | grep -v "epi_mode\|bn=64 \|bn=128\|bn=192"
echo "== trace mode 16"; SDB_NO_EPI_W64=1 timeout 300 python tools/gemm_trace.py linear_qk_65536x320x640 --mode 16 2>&1 | tail -8
