#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== half token stream (default)"; timeout 600 python tools/diag_fold_error.py --96 2>&1 | grep latent; timeout 600 python tools/diag_fold_error.py 2>&1 | grep latent
echo "== SDB_TOK_FP32=1"; SDB_TOK_FP32=1 timeout 600 python tools/diag_fold_error.py --96 2>&1 | grep latent; SDB_TOK_FP32=1 timeout 600 python tools/diag_fold_error.py 2>&1 | grep latent
} > gpurun_out/r02z_error_tok_half_vs_fp32.log 2>&1
cat gpurun_out/r02z_error_tok_half_vs_fp32.log
