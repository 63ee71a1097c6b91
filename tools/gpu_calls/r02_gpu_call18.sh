#!/bin/bash
# round 2, call 18 (8 GPUs): final tree - weak scaling at 8 and 4 ranks, configs[3] at 8 ranks, single-rank line
set -u
mkdir -p gpurun_out
T=r02q
for N in 8 4 2; do
  P=$((20000 + RANDOM % 20000))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_weak_n$N.json 2> gpurun_out/${T}_bench_weak_n$N.err
  echo "weak N=$N rc=$?"
done
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 5 --warmup 3 --config 3 > gpurun_out/${T}_bench_strong_n8.json 2> gpurun_out/${T}_bench_strong_n8.err
echo "strong N=8 rc=$?"
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_weak_n1.json 2> gpurun_out/${T}_bench_weak_n1.err
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 800 > gpurun_out/${T}_tests_multigpu.log 2>&1; tail -2 gpurun_out/${T}_tests_multigpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02q_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],2), d['scaling'], d['n_gpus'], d['config']['batch_per_gpu'], round(d['ms_per_step'],1), round((d['e2e'] or {}).get('value',0),2), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
