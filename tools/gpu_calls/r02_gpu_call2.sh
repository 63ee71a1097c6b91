#!/bin/bash
# round 2, call 2: filter-column staging (a3) + IEEE-half operands at the full-resolution level
set -u
mkdir -p gpurun_out
T=r02b
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/${T}_tests_kernels.log 2>&1
echo "kernel tests rc=$?" >> gpurun_out/${T}_tests_kernels.log
tail -25 gpurun_out/${T}_tests_kernels.log | cut -c1-250
timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -q -x > gpurun_out/${T}_tests_modules.log 2>&1
echo "module tests rc=$?" >> gpurun_out/${T}_tests_modules.log
tail -15 gpurun_out/${T}_tests_modules.log | cut -c1-250
timeout 900 python -m pytest tests/test_parity_configs_gpu.py -m gpu -q -s -k "768_unet or benched" > gpurun_out/${T}_tests_parity.log 2>&1
grep -E "^\[|passed|failed" gpurun_out/${T}_tests_parity.log | cut -c1-250
SDB_NO_A3=1 timeout 300 python tools/kernel_bench.py --graph --only conv3x3 > gpurun_out/${T}_kb_noa3.log 2>&1
timeout 300 python tools/kernel_bench.py --graph --only conv3x3 > gpurun_out/${T}_kb_a3.log 2>&1
SDB_A3_CHOOSER=0 timeout 300 python tools/kernel_bench.py --graph --only conv3x3 > gpurun_out/${T}_kb_a3_oldchooser.log 2>&1
paste <(cut -c1-75 gpurun_out/${T}_kb_noa3.log) <(cut -c50-75 gpurun_out/${T}_kb_a3.log) <(cut -c50-75 gpurun_out/${T}_kb_a3_oldchooser.log)
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
SDB_NO_A3=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_noa3.json 2> gpurun_out/${T}_bench_noa3.err
python - <<'PY'
import json
for f in ('gpurun_out/r02b_bench.json','gpurun_out/r02b_bench_noa3.json'):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, d['value'], (d['e2e'] or {}).get('value'), d['detail']['unet_step_ms'], d['clocks'])
    for r in d['detail']['rooflines']: print('  ', r['kernel'][:120], round(r['achieved'],1), round(r['frac'],3), round(r['us_per_launch'],1))
PY
