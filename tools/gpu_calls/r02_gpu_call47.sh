#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 500 -s -k "non_square" > gpurun_out/r02as_test_non_square.log 2>&1; grep -E "rel_err|PSNR|passed|failed|Error|error" gpurun_out/r02as_test_non_square.log | cut -c1-250 | tail -12
