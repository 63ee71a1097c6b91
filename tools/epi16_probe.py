"""Probe: why do GEMMs with a 16-bit output run at ~2.9 TB/s when the fp32 + residual ones reach 5.5 TB/s?
CUDA-graph timing of 65536 x 320 x {320, 640, 768} with bf16 / fp32 outputs, forced epilogue modes and tile widths."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_stable_diffusion_b200 import ops  # noqa: E402

DEV = "cuda"


def timed(fn, iters=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


M, K = 65536, 320
a = [(torch.randn(M, K, device=DEV)).bfloat16() for _ in range(2)]
for N in (320, 640, 768, 1280):
    w = (torch.randn(N, K, device=DEV) * K ** -0.5).bfloat16()
    b = torch.randn(N, device=DEV)
    r = torch.randn(M, N, device=DEV)
    rh = r.half()
    i = [0]
    def A():
        i[0] ^= 1
        return a[i[0]]
    for name, kw, nbytes in (
            ("bf16 out            ", dict(), M * K * 2 + M * N * 2),
            ("bf16 out epi_mode=1 ", dict(epi_mode=1), M * K * 2 + M * N * 2),
            ("bf16 out bn=320     ", dict(block_n=320, nsplit=1), M * K * 2 + M * N * 2),
            ("bf16 out bn=256     ", dict(block_n=256, nsplit=1), M * K * 2 + M * N * 2),
            ("bf16 out bn=192     ", dict(block_n=192, nsplit=1), M * K * 2 + M * N * 2),
            ("bf16 out bn=160     ", dict(block_n=160, nsplit=1), M * K * 2 + M * N * 2),
            ("f16 out + f16 res        ", dict(residual=rh, out16=torch.float16), M * K * 2 + M * N * 4),
            ("f16 out + f16 res bn=320 ", dict(residual=rh, out16=torch.float16, block_n=320, nsplit=1), M * K * 2 + M * N * 4),
            ("f16 out + f16 res bn=256 ", dict(residual=rh, out16=torch.float16, block_n=256, nsplit=1), M * K * 2 + M * N * 4),
            ("f16 out + f16 res bn=192 ", dict(residual=rh, out16=torch.float16, block_n=192, nsplit=1), M * K * 2 + M * N * 4),
            ("f16 out + f16 res bn=160 ", dict(residual=rh, out16=torch.float16, block_n=160, nsplit=1), M * K * 2 + M * N * 4),
            ("f16 out + f16 res bn=128 ", dict(residual=rh, out16=torch.float16, block_n=128, nsplit=1), M * K * 2 + M * N * 4),
            ("bf16 out bn=128     ", dict(block_n=128, nsplit=1), M * K * 2 + M * N * 2),
            ("bf16 out bn=64      ", dict(block_n=64, nsplit=1), M * K * 2 + M * N * 2),
            ("fp32 out            ", dict(out_fp32=True), M * K * 2 + M * N * 4),
            ("fp32 out + bf16 copy", dict(out_fp32=True, out2=True), M * K * 2 + M * N * 6),
            ("fp32 out + fp32 res ", dict(out_fp32=True, residual=r), M * K * 2 + M * N * 8)):
        try:
            t = timed(lambda: ops.linear(A(), w, bias=b, **kw))
            print(f"N={N:5d} {name} {t:7.1f} us  {nbytes / t / 1e6:7.2f} TB/s", flush=True)
        except Exception as e:   # noqa: BLE001
            print(f"N={N:5d} {name} failed: {e}", flush=True)
