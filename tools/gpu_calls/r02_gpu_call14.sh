#!/bin/bash
# round 2, call 14: Upsample folded into four 2x2 phase convolutions - parity, UNet-level parity, bench A/B
set -u
mkdir -p gpurun_out
T=r02n
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "up2x or conv3x3" --timeout 300 > gpurun_out/${T}_tests_up.log 2>&1; tail -12 gpurun_out/${T}_tests_up.log | cut -c1-250
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 600 -s -k "diffusion or benched or 768_unet or generate_short or graph" > gpurun_out/${T}_tests_unet.log 2>&1; grep -E "^\[|^\.\[|passed|failed" gpurun_out/${T}_tests_unet.log | cut -c1-230 | tail -25
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
SDB_NO_FOLD_UPSAMPLE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_nofold.json 2> gpurun_out/${T}_bench_nofold.err
python - <<'PY'
import json
for f in ('gpurun_out/r02n_bench.json','gpurun_out/r02n_bench_nofold.json'):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f,'unreadable',e); continue
    print(f, round(d['value'],3), d['e2e'] and round(d['e2e']['value'],3), round(d['detail']['unet_step_ms'],3), d['clocks']['sm_mhz'], d['config']['unet_gflop_per_image_step_algorithmic'], round(d['detail']['unet_tensor_frac_of_sustained_peak'],4))
    for s in d['detail']['kernels']['shapes_eager_events'][:40]:
        if 'taps=4' in s['shape']: print('   ', s)
PY
