#!/bin/bash
# round 2, call 5: two MMA issuer warps in the two-tile attention kernel - parity, timing, stagger sweep, bench
set -u
mkdir -p gpurun_out
T=r02e
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/${T}_tests_attn.log 2>&1; tail -4 gpurun_out/${T}_tests_attn.log | cut -c1-200
SDB_ATTN_ONE_ISSUER=1 timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_ > gpurun_out/${T}_kb_attn_one.log 2>&1
timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_ > gpurun_out/${T}_kb_attn_two.log 2>&1
paste <(cut -c1-75 gpurun_out/${T}_kb_attn_one.log) <(cut -c50-75 gpurun_out/${T}_kb_attn_two.log)
for sg in 0 -600 -1000 -1600 -2200; do
  echo "stagger $sg: $(SDB_ATTN_STAGGER=$sg timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_self_S4096 2>&1 | head -1)"
done | tee gpurun_out/${T}_stagger_sweep.log
timeout 600 python -m pytest tests/test_modules_gpu.py -m gpu -q -x -k "diffusion or generate_short or attention" > gpurun_out/${T}_tests_modules.log 2>&1; tail -3 gpurun_out/${T}_tests_modules.log | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print('bench', d['value'], (d['e2e'] or {}).get('value'), d['detail']['unet_step_ms'], d['clocks'])
for r in d['detail']['rooflines']: print('  ', r['kernel'][:120], round(r['achieved'],1), round(r['frac'],3), round(r['us_per_launch'],1))
PY
