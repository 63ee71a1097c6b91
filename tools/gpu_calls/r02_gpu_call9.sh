#!/bin/bash
# round 2, call 9: GroupNorm apply with prefetch, full GPU suite on the current tree, bench lines (512^2 and 768^2)
set -u
mkdir -p gpurun_out
T=r02i
timeout 300 python tools/small_kernels_bench.py > gpurun_out/${T}_small_kernels.log 2>&1; grep "gn " gpurun_out/${T}_small_kernels.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_tests_all.log 2>&1
echo "all gpu tests rc=$?" >> gpurun_out/${T}_tests_all.log
tail -6 gpurun_out/${T}_tests_all.log | cut -c1-220
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --config 4 --no-cpu-baseline > gpurun_out/${T}_bench_768.json 2> gpurun_out/${T}_bench_768.err
echo "bench768 rc=$?"; tail -2 gpurun_out/${T}_bench_768.err | cut -c1-200
python - <<'PY'
import json
for f in ('gpurun_out/r02i_bench.json','gpurun_out/r02i_bench_768.json'):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f,'unreadable',e); continue
    print(f, round(d['value'],3), round((d['e2e'] or {}).get('value',0),3), round(d['detail']['unet_step_ms'],3), d['detail']['vae_decode_ms'], d['clocks']['sm_mhz'], d['cpu_baseline'] and (d['cpu_baseline']['kind'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
    for r in d['detail']['rooflines']: print('  ', r['kernel'][:110], round(r['achieved'],1), round(r['frac'],3), round(r['us_per_launch'],1), r['traffic'])
PY
