#!/bin/bash
# round 2, call 3: MMA dependency-chain micro-benchmark, 8x8-level conv timeline, remaining parity tests, resize test
set -u
mkdir -p gpurun_out
T=r02c
timeout 120 tools/micro/mma_chain > gpurun_out/${T}_mma_chain.log 2>&1; cat gpurun_out/${T}_mma_chain.log
timeout 300 python tools/gemm_trace.py conv3x3_res_1280_1280_8 > gpurun_out/${T}_trace_conv8.log 2>&1; head -12 gpurun_out/${T}_trace_conv8.log
timeout 300 python tools/small_kernels_bench.py > gpurun_out/${T}_small_kernels.log 2>&1; cat gpurun_out/${T}_small_kernels.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "resize or uint8" > gpurun_out/${T}_tests_resize.log 2>&1; tail -5 gpurun_out/${T}_tests_resize.log | cut -c1-200
timeout 1500 python -m pytest tests/test_parity_configs_gpu.py -m gpu -q -s -k "generate or graph or img2img" > gpurun_out/${T}_tests_parity.log 2>&1
grep -E "^\[|^\.\[|passed|failed|Error|assert" gpurun_out/${T}_tests_parity.log | cut -c1-250
