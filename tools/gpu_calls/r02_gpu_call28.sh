#!/bin/bash
# round 2, call 28: IEEE-half token stream inside the attention blocks
set -u
mkdir -p gpurun_out
T=r02z
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "tma_epilogue or layernorm or linear" > gpurun_out/${T}_tests_k.log 2>&1; tail -5 gpurun_out/${T}_tests_k.log | cut -c1-250
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or benched or 768_unet or generate_short or blocks or cfg" > gpurun_out/${T}_tests.log 2>&1; grep -E "max=|rel_err|PSNR|passed|failed|Error" gpurun_out/${T}_tests.log | cut -c1-200 | tail -32
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
SDB_TOK_FP32=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_tokfp32.json 2> gpurun_out/${T}_bench_tokfp32.err
python - <<'PY'
import json
for f in ('gpurun_out/r02z_bench.json','gpurun_out/r02z_bench_tokfp32.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'])
        k=d['detail']['kernels']
        print({x:(k[x]['launches'],k[x]['ms']) for x in ('groupnorm','layernorm','attention','gemm_tc_conv3x3','gemm_tc_linear')})
    except Exception as e:
        print(f,'unreadable',e)
PY
