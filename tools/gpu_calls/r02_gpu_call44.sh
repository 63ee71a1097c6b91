#!/bin/bash
# round 2, call 44: final confirmation - full GPU suite, smoke, both bench arms (driver-like), 768^2 config
set -u
mkdir -p gpurun_out
T=r02ap
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_tests_all.log 2>&1; tail -3 gpurun_out/${T}_tests_all.log | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log | cut -c1-250
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_config4.json 2> gpurun_out/${T}_bench_config4.err
python - <<'PY'
import json
for f in ('gpurun_out/r02ap_bench_reference.json','gpurun_out/r02ap_bench.json','gpurun_out/r02ap_bench_config4.json'):
    try:
        d=json.load(open(f)); print(f, d.get('impl'), round(d['value'],5), d['e2e']['value'], d.get('ms_per_step'), d.get('clocks'))
        if 'detail' in d: print('   ', round(d['detail']['unet_step_ms'],3), round(d['detail']['vae_decode_ms'],2), d['detail']['launches_per_graph'], d['gpu_launches'], d.get('cpu_baseline'))
        for k in ('roofline','roofline_hbm','roofline_attention'):
            if d.get(k): print('   ', k, d[k]['kernel'][:90], round(d[k]['achieved'],1), d[k]['unit'], round(d[k]['frac'],3), d[k]['traffic'], round(d[k]['us_per_launch'],1))
    except Exception as e:
        print(f,'unreadable',e)
PY
