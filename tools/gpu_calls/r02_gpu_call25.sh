#!/bin/bash
# round 2, call 25: gn_reduce_partials flat walk (VAE) + per-batch-size step time (L2 residency question)
set -u
mkdir -p gpurun_out
T=r02w
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "groupnorm or gn_" > gpurun_out/${T}_tests_k.log 2>&1; tail -3 gpurun_out/${T}_tests_k.log | cut -c1-200
timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or vae or blocks or generate_short" > gpurun_out/${T}_tests.log 2>&1; grep -E "rel_err|PSNR|passed|failed" gpurun_out/${T}_tests.log | cut -c1-200 | tail -16
for B in 8 4 2; do
timeout 600 python bench.py --steps 3 --warmup 3 --batch $B --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_b$B.json 2> gpurun_out/${T}_bench_b$B.err
done
python - <<'PY'
import json
for b in (8,4,2):
    f=f'gpurun_out/r02w_bench_b{b}.json'
    try:
        d=json.load(open(f)); print(f, round(d['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f,'unreadable',e)
PY
