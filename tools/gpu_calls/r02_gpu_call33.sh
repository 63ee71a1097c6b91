#!/bin/bash
# round 2, call 33: 64-column epilogue units, bench A/B + parity
set -u
mkdir -p gpurun_out
T=r02ad
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or benched or blocks or generate_short or clip" > gpurun_out/${T}_tests.log 2>&1; grep -E "max=|rel_err|PSNR|passed|failed|Error" gpurun_out/${T}_tests.log | cut -c1-200 | tail -14
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
SDB_NO_EPI_W64=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_now64.json 2> gpurun_out/${T}_bench_now64.err
python - <<'PY'
import json
for f in ('gpurun_out/r02ad_bench.json','gpurun_out/r02ad_bench_now64.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['detail']['clip_ms'], d['clocks']['sm_mhz'])
        k=d['detail']['kernels']
        print({x:(k[x]['launches'],k[x]['ms']) for x in ('groupnorm','layernorm','attention','gemm_tc_conv3x3','gemm_tc_linear')})
    except Exception as e:
        print(f,'unreadable',e)
PY
