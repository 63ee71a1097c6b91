"""GroupNorm(+SiLU) apply pass with the statistics taken from epilogue partial sums (the resblock path): time per
reduce + apply pair inside a CUDA graph, cycling over operand sets whose footprint exceeds the L2.
usage: python tools/gn_apply_bench.py [--batch 8]      (env SDB_GN_APPLY_BPS = grid target, blocks per SM)"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_stable_diffusion_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--sets", type=int, default=6)
ap.add_argument("--iters", type=int, default=24)
args = ap.parse_args()
N = 2 * args.batch
dev = "cuda"
CASES = [("hid 320@64 f16", 64, 320, 0, torch.float16), ("x 320@64 f32", 64, 320, 0, torch.float32),
         ("cat 640@64 f32", 64, 320, 320, torch.float32), ("cat 960@64 f32", 64, 640, 320, torch.float32),
         ("hid 640@32 f16", 32, 640, 0, torch.float16), ("x 640@32 f32", 32, 640, 0, torch.float32),
         ("cat 1280@32 f32", 32, 640, 640, torch.float32), ("hid 1280@16 f16", 16, 1280, 0, torch.float16),
         ("x 1280@16 f32", 16, 1280, 0, torch.float32), ("cat 2560@16 f32", 16, 1280, 1280, torch.float32),
         ("vae 128@512 f16 (N=8)", 512, 128, 0, torch.float16), ("vae 256@256 f16 (N=8)", 256, 256, 0, torch.float16)]


def partials(x):
    n, h, w, c = x.shape
    k = max(1, h * w // 128)
    xs = x.float().view(n, k, -1, c)
    return torch.stack([xs.sum(2), (xs * xs).sum(2)], dim=-1).contiguous()


for name, hh, c0, c1, dt in CASES:
    n = args.batch if name.startswith("vae") else N
    sets = []
    for i in range(args.sets):
        x0 = torch.randn(n, hh, hh, c0, device=dev).to(dt)
        x1 = torch.randn(n, hh, hh, c1, device=dev) if c1 else None
        sets.append((x0, x1, partials(x0), partials(x1) if c1 else None))
    g = torch.randn(c0 + c1, device=dev)
    b = torch.randn(c0 + c1, device=dev)
    odt = torch.float16 if dt == torch.float16 else torch.bfloat16

    def run(i):
        x0, x1, p0, p1 = sets[i % len(sets)]
        return ops.groupnorm(x0, g, b, x1=x1, silu=True, part0=p0, part1=p1, out_dtype=odt, fused=False)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(args.iters):
            run(i)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * args.iters)
    nbytes = n * hh * hh * (c0 * (sets[0][0].element_size() + 2) + c1 * 6)
    print(f"{name:26s} {us:7.2f} us  {nbytes / us * 1e-3:7.1f} GB/s   ({nbytes / 1e6:.0f} MB)", flush=True)
