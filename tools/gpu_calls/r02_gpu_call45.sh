#!/bin/bash
# round 2, call 45: polynomial share of the exponentials in the folded self-attention loop: 1/4 vs 3/8
set -u
mkdir -p gpurun_out
{
echo "== quarter (default)"; timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_self_S4096 --iters 10 2>&1 | grep attn_self
echo "== three of eight (SDB_ATTN_POLY38=1)"; SDB_ATTN_POLY38=1 timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_self_S4096 --iters 10 2>&1 | grep attn_self
SDB_ATTN_POLY38=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "qk_fold" 2>&1 | tail -2
} > gpurun_out/r02aq_attn_poly38.log 2>&1
cat gpurun_out/r02aq_attn_poly38.log
