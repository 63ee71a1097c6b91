#!/bin/bash
# round 2, call 17: CFG shared prefix - parity and bench A/B
set -u
mkdir -p gpurun_out
T=r02p
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --timeout 600 -s -k "cfg_pair or generate or graph or without_cfg" > gpurun_out/${T}_tests.log 2>&1; grep -E "^\[|^\.\[|passed|failed|Error" gpurun_out/${T}_tests.log | cut -c1-230 | tail -30
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err | cut -c1-200
SDB_NO_SHARE_CFG_PREFIX=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_noshare.json 2> gpurun_out/${T}_bench_noshare.err
python - <<'PY'
import json
for f in ('gpurun_out/r02p_bench.json','gpurun_out/r02p_bench_noshare.json'):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f,'unreadable',e); continue
    print(f, round(d['value'],3), d['e2e'] and round(d['e2e']['value'],3), round(d['detail']['unet_step_ms'],3), d['clocks']['sm_mhz'], d['config']['unet_gflop_per_image_step_algorithmic'], round(d['detail']['whole_job_tensor_frac_of_sustained_peak'],4), d['detail']['launches_per_graph'])
PY
