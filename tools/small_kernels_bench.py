import sys, torch
sys.path.insert(0, '/root/repo')
from pytorch_stable_diffusion_b200 import ops
dev='cuda'
x=torch.randn(16,64,64,4,device=dev)
w=torch.randn(320,9,4,device=dev); b=torch.randn(320,device=dev)
xs=torch.randn(16,32,32,1280,device=dev).bfloat16()
lat=torch.randn(8,4,64,64,device=dev); eps=torch.randn(16,64,64,4,device=dev); nz=torch.randn(8,4,64,64,device=dev)
coef=torch.rand(50,5,device=dev); nxt=torch.empty(16,64,64,4,device=dev)
cases={"conv_direct 4->320 @64 (fp32 in, fp32+bf16 out)": lambda: ops.conv_direct(x,w,b,320,3,out_fp32=True,out2=True),
       "upsample2x 16x32x32x1280": lambda: ops.upsample2x(xs),
       "cfg_ddpm_step": lambda: ops.cfg_ddpm_step(lat, eps, nz, coef, 3, 7.5, True, nxt)}
for name,fn in cases.items():
    fn(); torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:50s} {1e3*e0.elapsed_time(e1)/20:8.1f} us")

# GroupNorm statistics from epilogue partials: the reduce kernel alone, then reduce + apply, vs stats + apply
import ctypes
from pytorch_stable_diffusion_b200 import _ext
lib = _ext.lib()
for (n, hw, c) in ((16, 4096, 320), (16, 1024, 640)):
    xg = torch.randn(n, hw, c, device=dev)
    gam = torch.randn(c, device=dev); bet = torch.randn(c, device=dev)
    part = torch.randn(n, hw // 32, c, 2, device=dev)
    stats = torch.empty((lib.sdb_groupnorm_stats_bytes(n, 32) // 8,), device=dev, dtype=torch.float64)
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cases2 = {f"gn reduce_partials only {n}x{hw}x{c}": lambda: lib.sdb_groupnorm_reduce_partials(
                  ctypes.c_void_p(part.data_ptr()), None, ctypes.c_void_p(stats.data_ptr()), n, hw // 32, 0, c, 0, 32, st()),
              f"gn reduce + apply {n}x{hw}x{c}": lambda: ops.groupnorm(xg, gam, bet, silu=True, fused=False, part0=part),
              f"gn stats + apply {n}x{hw}x{c}": lambda: ops.groupnorm(xg, gam, bet, silu=True, fused=False)}
    for name, fn in cases2.items():
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print(f"{name:50s} {1e3*e0.elapsed_time(e1)/20:8.1f} us")
