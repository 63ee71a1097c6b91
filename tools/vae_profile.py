"""Per-kernel timing of one VAE decode (eager, CUDA events per launch) at the benchmark's batch.
usage: python tools/vae_profile.py [--batch 8] [--hw 64]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_stable_diffusion_b200 import ops, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--hw", type=int, default=64)
args = ap.parse_args()
dev = "cuda"
models = synthetic.build_models(dev, which=("decoder",))
dec = models["decoder"]
lat = torch.randn(args.batch, 4, args.hw, args.hw, device=dev)
with torch.no_grad():
    for _ in range(2):
        dec.decode_nhwc(lat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dec.decode_nhwc(lat)
    e1.record()
    torch.cuda.synchronize()
    print(f"decode, batch {args.batch}: {e0.elapsed_time(e1):.2f} ms (eager, no profiler)")
    ops.PROFILER = ops.LaunchProfiler()
    dec.decode_nhwc(lat)
    summ = ops.PROFILER.summary(by_shape=True)
    ops.PROFILER = None
tot = sum(v["ms"] for v in summ.values())
print(f"sum of launches {tot:.2f} ms")
for (name, shape), v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    tf = v["flops"] / v["ms"] * 1e-9 if v["flops"] else 0.0
    print(f"{v['ms']:8.3f} ms {v['launches']:3d} x {1e3 * v['ms'] / v['launches']:8.1f} us {tf:7.0f} TF {v['bytes'] / v['ms'] * 1e-6:7.0f} GB/s  {name} {shape}")
