#!/bin/bash
# round 2, call 10: two extra-source k-blocks per stage (conv_merged + skip conv), bench, the reference arm for real
set -u
mkdir -p gpurun_out
T=r02j
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "conv3x3" --timeout 300 > gpurun_out/${T}_tests_conv.log 2>&1; tail -3 gpurun_out/${T}_tests_conv.log | cut -c1-200
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 300 -k "diffusion or benched or 768_unet_eval_margin" > gpurun_out/${T}_tests_unet.log 2>&1; tail -3 gpurun_out/${T}_tests_unet.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench.json'))
print('bench', round(d['value'],3), round(d['e2e']['value'],3), round(d['detail']['unet_step_ms'],3), d['clocks']['sm_mhz'])
for r in d['detail']['rooflines']: print('  ', r['kernel'][:110], round(r['achieved'],1), round(r['frac'],3), round(r['us_per_launch'],1))
for s in d['detail']['kernels']['shapes_eager_events'][:40]:
    if 'conv3x3' in s['shape'] and 'cx=0' not in s['shape']: print('   ', s)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/${T}_bench_reference_arm.json
timeout 1500 python bench.py --impl reference --full --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference_full.json 2> gpurun_out/${T}_bench_reference_full.err; echo "ref full rc=$?"; cut -c1-300 gpurun_out/${T}_bench_reference_full.json
