#!/bin/bash
# round 2, call 46 (8 GPUs): weak-scaling line of the final code
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02ar_bench_weak_n8.json 2> gpurun_out/r02ar_bench_weak_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02ar_bench_weak_n8.json')); print(round(d['value'],3), d['e2e']['value'], d['n_gpus'], d['ms_per_step'], d['clocks'])
PY
