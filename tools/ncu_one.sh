#!/bin/bash
# usage: [KB_ARGS='--attn-mode 3'] tools/ncu_one.sh <tag> <kernel_bench case substring> [kernel regex]
# Plain run first (must exit 0), then one `ncu --set full` capture of the 2nd matching launch.
TAG=$1; CASE=$2; KRE=${3:-'regex:^(gemm_tc|attn2?_tc)'}
python tools/kernel_bench.py ${KB_ARGS:-} --only $CASE --iters 1 --warmup 1 > gpurun_out/${TAG}_plain_$CASE.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$KRE" -s 1 -c 1 -f \
    -o gpurun_out/${TAG}_prof_$CASE python tools/kernel_bench.py ${KB_ARGS:-} --only $CASE --iters 1 --warmup 1 \
    > gpurun_out/${TAG}_ncu_$CASE.log 2>&1
tail -2 gpurun_out/${TAG}_plain_$CASE.log
