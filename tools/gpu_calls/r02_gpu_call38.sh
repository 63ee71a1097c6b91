#!/bin/bash
# round 2, call 38: all four up-sampling phases in one launch
set -u
mkdir -p gpurun_out
T=r02ai
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -x -k "conv_up2x or conv3x3" > gpurun_out/${T}_tests_k.log 2>&1; tail -5 gpurun_out/${T}_tests_k.log | cut -c1-250
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or benched_batch or vae or generate_short" > gpurun_out/${T}_tests.log 2>&1; grep -E "max=|rel_err|PSNR|passed|failed|Error" gpurun_out/${T}_tests.log | cut -c1-200 | tail -14
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
SDB_UP_SEPARATE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_sep.json 2> gpurun_out/${T}_bench_sep.err
python - <<'PY'
import json
for f in ('gpurun_out/r02ai_bench.json','gpurun_out/r02ai_bench_sep.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'], d['detail']['launches_per_graph'])
        for r in d['detail']['kernels']['shapes_eager_events']:
            if 'taps=4' in r['shape']: print('   ', r['launches'], r['us'], r['tflops'], r['shape'][:110])
    except Exception as e:
        print(f,'unreadable',e)
PY
