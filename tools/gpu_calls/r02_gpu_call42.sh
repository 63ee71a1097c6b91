#!/bin/bash
# round 2, call 42 (2 GPUs): NCCL sharding test + weak-scaling bench at N=2 with the final code
set -u
mkdir -p gpurun_out
T=r02an
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 600 -s > gpurun_out/${T}_tests_multigpu.log 2>&1; tail -4 gpurun_out/${T}_tests_multigpu.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_weak_n2.json 2> gpurun_out/${T}_bench_weak_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02an_bench_weak_n2.json')); print(round(d['value'],3), d['e2e']['value'], d['n_gpus'], d['ms_per_step'], d['clocks'])
PY
