#!/bin/bash
# round 2, call 24: hidden tensor in IEEE half - parity and bench A/B
set -u
mkdir -p gpurun_out
T=r02v
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "groupnorm or half or gn_partials" > gpurun_out/${T}_tests_k.log 2>&1; tail -3 gpurun_out/${T}_tests_k.log | cut -c1-200
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or benched or 768_unet or generate_short or vae or blocks" > gpurun_out/${T}_tests.log 2>&1; grep -E "max=|rel_err|PSNR|passed|failed" gpurun_out/${T}_tests.log | cut -c1-200 | tail -32
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
SDB_HID_FP32=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_fp32hid.json 2> gpurun_out/${T}_bench_fp32hid.err
python - <<'PY'
import json
for f in ('gpurun_out/r02v_bench.json','gpurun_out/r02v_bench_fp32hid.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f,'unreadable',e)
PY
