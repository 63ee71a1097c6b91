#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02at_bench_b32.json 2> gpurun_out/r02at_bench_b32.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02at_bench_b32.json')); print(round(d['value'],3), d['config']['batch_per_gpu'], round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'])
PY
tail -3 gpurun_out/r02at_bench_b32.err
