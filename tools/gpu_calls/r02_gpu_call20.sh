#!/bin/bash
# round 2, call 20: the driver's own end-of-round sequence on the final tree - GPU suite, smoke, both bench arms
set -u
mkdir -p gpurun_out
T=r02s
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/${T}_tests_all.log 2>&1
echo "all gpu tests rc=$?" >> gpurun_out/${T}_tests_all.log
tail -4 gpurun_out/${T}_tests_all.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log | cut -c1-250
SECONDS=0
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm rc=$? in ${SECONDS}s"
SECONDS=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? in ${SECONDS}s"
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02s_bench_reference.json')); d=json.load(open('gpurun_out/r02s_bench.json'))
print('reference', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'], r['ms_per_step'])
print('b200', round(d['value'],3), round(d['e2e']['value'],3), round(d['ms_per_step'],1), round(d['detail']['unet_step_ms'],3), d['clocks'], d['gpu_launches'])
print('ratio', d['value']/r['value'], 'e2e ratio', d['e2e']['value']/r['value'])
for k in ('roofline','roofline_attention','roofline_hbm'):
    x=d[k]; print(k, x['kernel'][:100], round(x['achieved'],1), x['unit'], round(x['frac'],3), round(x['us_per_launch'],1), x['traffic'])
print(d['cpu_baseline'])
PY
