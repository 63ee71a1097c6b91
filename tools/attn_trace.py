"""Per-key-block timeline of CTA 0 of the two-tile attention kernel (debug aid).
usage: python tools/attn_trace.py [attn_mode]"""
import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import kernel_bench as kb  # noqa: E402
from pytorch_stable_diffusion_b200 import _ext  # noqa: E402

kb.ATTN_MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib = _ext.lib()
for name, count, fl, fn in kb.cases(8):
    if "attn_self_S4096" not in name:
        continue
    fn(); fn()
    torch.cuda.synchronize()
    lib.sdb_debug_gemm_trace(1, None)
    fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 512)()
    lib.sdb_debug_gemm_trace(-1, buf)
    lib.sdb_debug_gemm_trace(0, None)
    t = [list(buf[i * 8:(i + 1) * 8]) for i in range(32)]
    base = min(x for r in t for x in r if x)
    print(name, "mode", kb.ATTN_MODE)
    print("  j   smWait  sFull   ldDone  maxDone expDone pStored | pv0Issued qkNextIssued")
    for i, r in enumerate(t[:20]):
        print(f"{i:3d} " + " ".join(f"{(x - base) if x else -1:8d}" for x in r))
    break
