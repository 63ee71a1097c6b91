"""Per-tile timeline of CTA 0 of the persistent GEMM kernel (debug aid).
usage: python tools/gemm_trace.py <kernel_bench case substring> [--pair N]"""
import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import kernel_bench as kb  # noqa: E402
from pytorch_stable_diffusion_b200 import _ext  # noqa: E402

only = sys.argv[1]
if "--pair" in sys.argv:
    kb.PAIR = int(sys.argv[sys.argv.index("--pair") + 1])
MODE = int(sys.argv[sys.argv.index("--mode") + 1]) if "--mode" in sys.argv else 1
lib = _ext.lib()
for name, count, fl, fn in kb.cases(8):
    if only not in name:
        continue
    fn(); fn()
    torch.cuda.synchronize()
    lib.sdb_debug_gemm_trace(MODE, None)
    fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 512)()
    lib.sdb_debug_gemm_trace(-1, buf)
    lib.sdb_debug_gemm_trace(0, None)
    t = [list(buf[i * 8:(i + 1) * 8]) for i in range(64)]
    t = [r for r in t if any(r)]
    base = min(x for r in t for x in (r[:3] + r[4:]) if x)
    if MODE & 4:
        for r in t:
            r[3] += base      # slot 3 holds accumulated cycles waiting on full barriers
    print(name, "tiles on CTA 0:", len(t))
    print("tile   prodStart prodDone | issWait  issFirst  issDone | epiWait accReady  stored   (SM clocks from first stamp)")
    for i, r in enumerate(t):
        print(f"{i:4d} " + " ".join(f"{(x - base) if x else -1:9d}" for x in r))
    break
