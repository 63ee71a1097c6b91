#!/bin/bash
# round 2, call 37: full GPU suite + bench after the epilogue / norm / token-stream changes
set -u
mkdir -p gpurun_out
T=r02ah
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/${T}_tests_all.log 2>&1; tail -4 gpurun_out/${T}_tests_all.log | cut -c1-250
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
f='gpurun_out/r02ah_bench.json'
d=json.load(open(f)); print(f, round(d['value'],3), d['e2e']['value'], round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks'])
k=d['detail']['kernels']
print({x:(k[x]['launches'],k[x]['ms']) for x in ('groupnorm','layernorm','attention','gemm_tc_conv3x3','gemm_tc_linear')})
print(json.dumps(d['roofline'])[:700]); print(json.dumps(d['roofline_hbm'])[:600])
PY
