"""Fixed cost of one kernel node inside a CUDA graph (launch + prologue + teardown), per kernel family."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_stable_diffusion_b200 import ops

dev = "cuda"
a = torch.randn(128, 64, device=dev).bfloat16(); w = torch.randn(128, 64, device=dev).bfloat16()
a2 = torch.randn(4096, 320, device=dev).bfloat16(); w2 = torch.randn(320, 320, device=dev).bfloat16()
x = torch.randn(256, 320, device=dev); g = torch.randn(320, device=dev); b = torch.randn(320, device=dev)
xg = torch.randn(2, 8, 8, 320, device=dev)
cases = {
    "gemm 128x64x128 (1 CTA)": lambda: ops.linear(a, w),
    "gemm 4096x320x320 (pairs)": lambda: ops.linear(a2, w2),
    "layernorm 256x320": lambda: ops.layernorm(x, g, b),
    "groupnorm 2x8x8x320 (2 kernels)": lambda: ops.groupnorm(xg, g, b, silu=True),
}
for name, fn in cases.items():
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    n = 200
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:36s} {1e3 * e0.elapsed_time(e1) / n:7.2f} us per call")
