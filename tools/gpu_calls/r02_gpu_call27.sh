#!/bin/bash
set -u
mkdir -p gpurun_out
for B in 2 3 4 6 8; do echo "== SDB_GN_APPLY_BPS=$B"; SDB_GN_APPLY_BPS=$B timeout 300 python tools/gn_apply_bench.py 2>&1 | tail -12; done > gpurun_out/r02y_gn_apply_bench.log 2>&1
cat gpurun_out/r02y_gn_apply_bench.log
