#!/bin/bash
# round 2, call 26: GroupNorm apply with raw loads staged in registers
set -u
mkdir -p gpurun_out
T=r02x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "groupnorm or gn_ or half" > gpurun_out/${T}_tests_k.log 2>&1; tail -3 gpurun_out/${T}_tests_k.log | cut -c1-200
timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or vae or blocks or generate_short" > gpurun_out/${T}_tests.log 2>&1; grep -E "rel_err|PSNR|passed|failed" gpurun_out/${T}_tests.log | cut -c1-200 | tail -16
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
f='gpurun_out/r02x_bench.json'
d=json.load(open(f)); print(f, round(d['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'])
k=d['detail']['kernels']
print({x:(k[x]['launches'],k[x]['ms']) for x in ('groupnorm','layernorm','attention','gemm_tc_conv3x3','gemm_tc_linear')})
for r in d['detail']['rooflines']:
    if 'gn_' in r['kernel']: print(r['kernel'][:120], r['us_per_launch'], r['achieved'])
PY
