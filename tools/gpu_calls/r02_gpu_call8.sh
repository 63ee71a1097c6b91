#!/bin/bash
# round 2, call 8: half-row softmax attention kernel - parity, timing vs attn2, timeline, bench
set -u
mkdir -p gpurun_out
T=r02h
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/${T}_tests_attn.log 2>&1; tail -6 gpurun_out/${T}_tests_attn.log | cut -c1-250
SDB_NO_ATTN_HALF=1 timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_self > gpurun_out/${T}_kb_attn2.log 2>&1
timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_self > gpurun_out/${T}_kb_attnh.log 2>&1
paste <(cut -c1-75 gpurun_out/${T}_kb_attn2.log) <(cut -c50-75 gpurun_out/${T}_kb_attnh.log)
for sg in 0 -600 -1300; do
  echo "half kernel, stagger $sg: $(SDB_ATTN_STAGGER=$sg timeout 300 python tools/kernel_bench.py --graph --attn-mode 3 --only attn_self_S4096 2>&1 | head -1)"
done | tee gpurun_out/${T}_stagger_sweep.log
timeout 300 python tools/attn_trace.py 3 > gpurun_out/${T}_attn_trace.log 2>&1; head -16 gpurun_out/${T}_attn_trace.log
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -x -k "diffusion_single or 768_unet_eval_margin or benched or generate_short" > gpurun_out/${T}_tests_modules.log 2>&1; tail -3 gpurun_out/${T}_tests_modules.log | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print('bench', d['value'], (d['e2e'] or {}).get('value'), d['detail']['unet_step_ms'], d['clocks'])
for r in d['detail']['rooflines']: print('  ', r['kernel'][:120], round(r['achieved'],1), round(r['frac'],3), round(r['us_per_launch'],1))
PY
