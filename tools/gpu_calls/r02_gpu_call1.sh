#!/bin/bash
# round 2, first GPU call: existing suite, parity at the benchmarked configs, bench line, per-layer error budget
set -u
mkdir -p gpurun_out
T=r02a
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_configs_gpu.py > gpurun_out/${T}_tests_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/${T}_tests_suite.log
timeout 1500 python -m pytest tests/test_parity_configs_gpu.py -m gpu -q -s > gpurun_out/${T}_tests_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/${T}_tests_parity.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 python tools/diag_layer_budget.py --hw 96 --seeds 31,32 > gpurun_out/${T}_layer_budget_96.log 2>&1
timeout 300 python tools/diag_layer_budget.py --hw 64 --seeds 21 > gpurun_out/${T}_layer_budget_64.log 2>&1
tail -5 gpurun_out/${T}_tests_suite.log; tail -40 gpurun_out/${T}_tests_parity.log | cut -c1-300; tail -3 gpurun_out/${T}_bench.err; head -c 600 gpurun_out/${T}_bench.json
