"""Sweep (block_n, nsplit, cta_pair) for the GEMM / conv shapes of the small UNet levels, CUDA-graph timed,
and print what ops._choose_tiling picks next to the best measured configuration."""
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


def bf(*s, scale=1.0):
    return (torch.randn(*s, device=DEV) * scale).bfloat16()


def conv_case(N, H, C0, C1, Cout):
    x0 = bf(N, H, H, C0)
    x1 = bf(N, H, H, C1) if C1 else None
    w = bf(Cout, 9 * (C0 + C1), scale=(9 * (C0 + C1)) ** -0.5)
    b = torch.randn(Cout, device=DEV)
    r = torch.randn(N * H * H, Cout, device=DEV)
    def run(bn, ns, pair):
        return ops.gemm(x0, w, Cout, kind=ops.GEMM_CONV3X3_S1, a1=x1, bias=b, conv_dims=(N, H, H), c0=C0, c1=C1,
                        residual=r, out_fp32=True, out2=True, block_n=bn, nsplit=ns, cta_pair=pair)
    return f"conv3x3 {C0 + C1}->{Cout} @{H}x{H} N={N}", run, 2.0 * N * H * H * Cout * 9 * (C0 + C1)


def lin_case(M, K, Cout):
    a = bf(M, K)
    w = bf(Cout, K, scale=K ** -0.5)
    b = torch.randn(Cout, device=DEV)
    r = torch.randn(M, Cout, device=DEV)
    def run(bn, ns, pair):
        return ops.linear(a, w, bias=b, residual=r, out_fp32=True, block_n=bn, nsplit=ns, cta_pair=pair)
    return f"linear {M}x{K}x{Cout} +res fp32", run, 2.0 * M * K * Cout


if "--big" in sys.argv:
    cases = [conv_case(16, 64, 320, 0, 320), conv_case(16, 32, 640, 0, 640), conv_case(16, 64, 320, 320, 320),
             conv_case(16, 32, 640, 640, 640), lin_case(65536, 320, 320)]
else:
  cases = [conv_case(16, 8, 1280, 0, 1280), conv_case(16, 8, 1280, 1280, 1280), conv_case(16, 16, 1280, 0, 1280),
         conv_case(16, 16, 1280, 1280, 1280), lin_case(4096, 1280, 1280), lin_case(1024, 1280, 1280),
         lin_case(16384, 640, 640)]
for name, run, fl in cases:
    res = []
    auto = timed(lambda: run(0, 0, 0))
    for bn in (80, 128, 160, 256, 320):
        for ns in ((1,) if "--big" in sys.argv else (1, 2, 3, 4, 6, 8)):
            for pair in (1, 2):
                try:
                    t = timed(lambda: run(bn, ns, pair), iters=10)
                except Exception as ex:   # unsupported combination
                    continue
                res.append((t, bn, ns, pair))
    res.sort()
    print(f"{name}: auto {auto:.1f} us ({fl / auto / 1e6:.0f} TFLOP/s) | best " +
          ", ".join(f"{t:.1f} us (bn={bn} ns={ns} cg={pair})" for t, bn, ns, pair in res[:4]), flush=True)
