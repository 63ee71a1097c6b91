#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -s -k "c_host_program" > gpurun_out/r02au_test_c_host.log 2>&1; grep -E "abi_linear|passed|failed|Error|error" gpurun_out/r02au_test_c_host.log | cut -c1-250 | tail -8
