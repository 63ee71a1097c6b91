#!/bin/bash
# round 2, call 16: final tree - full GPU suite, smoke, bench lines (512^2 with the CPU arm, 768^2)
set -u
mkdir -p gpurun_out
T=r02o
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_tests_all.log 2>&1
echo "all gpu tests rc=$?" >> gpurun_out/${T}_tests_all.log
tail -5 gpurun_out/${T}_tests_all.log | cut -c1-220
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log | cut -c1-250
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --config 4 --no-cpu-baseline > gpurun_out/${T}_bench_768.json 2> gpurun_out/${T}_bench_768.err; echo "bench768 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02o_bench.json','gpurun_out/r02o_bench_768.json'):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f,'unreadable',e); continue
    print(f, round(d['value'],3), round(d['e2e']['value'],3), round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['clocks']['sm_mhz'], round(d['detail']['whole_job_tensor_frac_of_sustained_peak'],4), d['gpu_launches'], d['cpu_baseline'] and (d['cpu_baseline']['kind'], round(d['cpu_baseline']['value'],5)))
    for r in d['detail']['rooflines']: print('  ', r['kernel'][:110], round(r['achieved'],1), round(r['frac'],3), round(r['us_per_launch'],1), r['traffic'])
PY
