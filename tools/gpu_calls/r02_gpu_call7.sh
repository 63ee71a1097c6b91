#!/bin/bash
# round 2, 2-GPU call: NCCL sharding test, strong-scaling (configs[3]) and weak-scaling lines at N = 2
set -u
mkdir -p gpurun_out
T=r02g
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s > gpurun_out/${T}_tests_multigpu.log 2>&1
grep -E "sharded x|SHARDED_OK|passed|failed|skipped" gpurun_out/${T}_tests_multigpu.log | cut -c1-200
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 3 --config 3 > gpurun_out/${T}_bench_strong_n2.json 2> gpurun_out/${T}_bench_strong_n2.err
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_weak_n2.json 2> gpurun_out/${T}_bench_weak_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r02g_bench_strong_n2.json','gpurun_out/r02g_bench_weak_n2.json'):
    try:
        d=json.load(open(f)); print(f, d['value'], d['scaling'], d['config']['batch_per_gpu'], d['ms_per_step'], (d['e2e'] or {}).get('value'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -3 gpurun_out/${T}_bench_strong_n2.err
