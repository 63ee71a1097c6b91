#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02aj_bench.json 2> gpurun_out/r02aj_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02aj_bench.json'))
print(round(d['value'],3), round(d['detail']['unet_step_ms'],3))
tot=0
for r in d['detail']['kernels']['shapes_eager_events']:
    t=r['us']*r['launches']; tot+=t
    print(f"{t:8.1f} {r['launches']:3d} x {r['us']:7.1f} us  {str(r['tflops']):>7} TF {r['gbs']:7.1f} GB/s  {r['shape'][:140]}")
print(tot)
PY
