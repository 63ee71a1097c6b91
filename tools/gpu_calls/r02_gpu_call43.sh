#!/bin/bash
# round 2, call 43: epilogue statistics + apply preferred over the one-pass GroupNorm from the 16x16 level up
set -u
mkdir -p gpurun_out
T=r02ao
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or benched_batch or blocks or generate_short or 768_unet" > gpurun_out/${T}_tests.log 2>&1; grep -E "max=|rel_err|PSNR|passed|failed|Error" gpurun_out/${T}_tests.log | cut -c1-200 | tail -24
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
SDB_GN_PARTS_MIN_HW=1000000 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_onepass.json 2> gpurun_out/${T}_bench_onepass.err
python - <<'PY'
import json
for f in ('gpurun_out/r02ao_bench.json','gpurun_out/r02ao_bench_onepass.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'], d['detail']['launches_per_graph'])
        k=d['detail']['kernels']
        print({x:(k[x]['launches'],k[x]['ms']) for x in ('groupnorm','layernorm','attention','gemm_tc_conv3x3','gemm_tc_linear')})
    except Exception as e:
        print(f,'unreadable',e)
PY
