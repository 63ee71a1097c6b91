"""Per-kernel micro-benchmark over the distinct problem sizes of one UNet evaluation at batch B
(M scales with N = 2B). CUDA-event timing, `--iters` back-to-back launches after
`--warmup`, an L2-sized scratch write between timed groups.

  python tools/kernel_bench.py [--batch 8] [--only substr] [--iters 10] [--json out.json]
A single launch of one case for ncu:  python tools/kernel_bench.py --only conv3x3_320_320_64 --iters 1 --warmup 0
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_stable_diffusion_b200 import _ext, ops  # noqa: E402

DEV = "cuda"
PAIR = 0
ATTN_MODE = 0     # 0 plain, 1 ones-row denominator, 2 ones-row + f16x2 exponentials


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).bfloat16()


def cases(B):
    N = 2 * B
    out = []

    def lin(name, S, K, Cout, count, res=False, fp32=False, half=False):
        """half=True: IEEE-half residual and output (the attention blocks' token stream)."""
        M = N * S
        a = bf(M, K)
        w = bf(Cout, K, scale=K ** -0.5)
        b = torch.randn(Cout, device=DEV)
        r = torch.randn(M, Cout, device=DEV) if res else None
        if half and r is not None:
            r = r.half()
        fl = 2.0 * M * K * Cout
        out.append((f"linear_{name}_{M}x{K}x{Cout}", count, fl,
                    lambda: ops.linear(a, w, bias=b, residual=r, out_fp32=fp32, cta_pair=PAIR,
                                       out16=torch.float16 if half else torch.bfloat16)))

    def conv(name, Hh, C0, C1, Cout, count, kind=ops.GEMM_CONV3X3_S1, res=False, cx=0, f16=False):
        """cx > 0: conv_merged of a channel-changing resblock - the 1x1 skip convolution of a cx-channel block input
        rides as extra k-blocks, IEEE-half operands, fp32 output with GroupNorm partial sums (the launch bench.py's
        `roofline` reports)."""
        dt = torch.float16 if f16 else torch.bfloat16
        x0 = bf(N, Hh, Hh, C0).to(dt)
        x1 = bf(N, Hh, Hh, C1).to(dt) if C1 else None
        ax = bf(N, Hh, Hh, cx).to(dt) if cx else None
        ktot = 9 * (C0 + C1) + cx
        w = bf(Cout, ktot, scale=ktot ** -0.5).to(dt)
        b = torch.randn(Cout, device=DEV)
        ho = Hh // 2 if kind != ops.GEMM_CONV3X3_S1 else Hh
        fl = 2.0 * N * ho * ho * Cout * ktot
        r = torch.randn(N * ho * ho, Cout, device=DEV) if res else None
        out.append((f"conv3x3_{name}_{C0 + C1}_{Cout}_{Hh}", count, fl,
                    lambda: ops.gemm(x0, w, Cout, kind=kind, a1=x1, bias=b, conv_dims=(N, Hh, Hh), c0=C0, c1=C1,
                                     residual=r, out_fp32=True, out2=True if res else None, cta_pair=PAIR, ax0=ax,
                                     gn_samples=N if cx else None)))

    def attn(name, S, Skv, d, count):
        heads = 8
        c = heads * d
        skv_pad = (Skv + 7) // 8 * 8
        q = bf(N * S, c)
        k = bf(N * skv_pad, c)
        sr = ATTN_MODE > 0 and d <= 112
        rows = heads * ((d + 16) // 16 * 16) if sr else c
        vt = bf(rows, N * skv_pad)
        if sr and ATTN_MODE == 2:
            vt = vt.to(torch.float16)
        o = torch.empty(N * S, c, device=DEV, dtype=torch.bfloat16)
        fl = 4.0 * N * heads * S * Skv * d
        if ATTN_MODE == 3 and sr and d <= 47 and Skv == S:
            # what the UNet's 64x64 self-attention runs: 48-column heads, pre-scaled q, ones column in k (qk_fold)
            R = 48
            qp = torch.zeros(N * S, heads, R, device=DEV, dtype=torch.bfloat16)
            kp = torch.zeros(N * S, heads, R, device=DEV, dtype=torch.bfloat16)
            qp[:, :, :d] = (q.view(N * S, heads, d).float() * (1.4427 / d ** 0.5)).bfloat16()
            kp[:, :, :d] = k.view(N * S, heads, d)
            kp[:, :, d] = 1.0
            vb = vt.to(torch.bfloat16)
            out.append((f"attn_{name}_S{S}_kv{Skv}_d{d}", count, fl,
                        lambda: ops.attention(qp.view(N * S, heads * R), kp.view(N * S, heads * R), vb, o, NB=N,
                                              heads=heads, d=d, S=S, Skv=Skv, Skv_pad=skv_pad, ldq=heads * R,
                                              ldk=heads * R, ldo=c, sum_row=True, q_prescaled=True, qk_cols=R,
                                              qk_fold=True)))
            return
        out.append((f"attn_{name}_S{S}_kv{Skv}_d{d}", count, fl,
                    lambda: ops.attention(q, k, vt, o, NB=N, heads=heads, d=d, S=S, Skv=Skv, Skv_pad=skv_pad,
                                          ldq=c, ldk=c, ldo=c, sum_row=sr, p_f16=sr and ATTN_MODE == 2)))

    def gnorm(name, Hh, C0, C1, count, fp32=True):
        x0 = torch.randn(N, Hh, Hh, C0, device=DEV)
        x0 = x0 if fp32 else x0.bfloat16()
        x1 = torch.randn(N, Hh, Hh, C1, device=DEV) if C1 else None
        g = torch.randn(C0 + C1, device=DEV)
        b = torch.randn(C0 + C1, device=DEV)
        nbytes = (2 * (4 if fp32 else 2) + 2) * N * Hh * Hh * C0 + (2 * 4 + 2) * N * Hh * Hh * C1
        out.append((f"groupnorm_{name}_{C0 + C1}_{Hh}_{'f32' if fp32 else 'bf16'}", count, nbytes * 1e3,   # "flops" column = bytes*1e3 -> reads as GB/s
                    lambda: ops.groupnorm(x0, g, b, x1=x1, silu=True)))

    def lnorm(S, C, count, half=False):
        x = torch.randn(N * S, C, device=DEV)
        x = x.half() if half else x
        g = torch.randn(C, device=DEV)
        b = torch.randn(C, device=DEV)
        out.append((f"layernorm_{'h16_' if half else ''}{N * S}x{C}", count, (4.0 if half else 6.0) * N * S * C * 1e3,
                    lambda: ops.layernorm(x, g, b)))

    def gnorm_parts(name, Hh, C0, count):
        """IEEE-half input, statistics from epilogue partial sums (the resblock's hidden tensor): reduce + apply."""
        x0 = torch.randn(N, Hh, Hh, C0, device=DEV).half()
        k = max(1, Hh * Hh // 128)
        xs = x0.float().view(N, k, -1, C0)
        part = torch.stack([xs.sum(2), (xs * xs).sum(2)], dim=-1).contiguous()
        g = torch.randn(C0, device=DEV)
        b = torch.randn(C0, device=DEV)
        out.append((f"groupnorm_{name}_{C0}_{Hh}_f16", count, 4.0 * N * Hh * Hh * C0 * 1e3,
                    lambda: ops.groupnorm(x0, g, b, silu=True, part0=part, out_dtype=torch.float16, fused=False)))

    gnorm("x", 64, 320, 0, 13)
    gnorm("hid", 64, 320, 0, 7, fp32=False)
    gnorm("cat", 64, 320, 320, 2)
    gnorm("x", 32, 640, 0, 8)
    gnorm("hid", 32, 640, 0, 6, fp32=False)
    gnorm("cat", 32, 640, 640, 2)
    gnorm("x", 16, 1280, 0, 8)
    gnorm("hid", 16, 1280, 0, 6, fp32=False)
    gnorm("cat", 16, 1280, 1280, 3)
    gnorm("x", 8, 1280, 0, 6)
    gnorm("cat", 8, 1280, 1280, 3)
    gnorm_parts("hidparts", 64, 320, 7)
    gnorm_parts("hidparts", 32, 640, 6)
    lnorm(4096, 320, 15)
    lnorm(1024, 640, 15)
    lnorm(256, 1280, 15)
    lnorm(4096, 320, 15, half=True)
    lnorm(1024, 640, 15, half=True)
    lnorm(256, 1280, 15, half=True)
    for S, C in ((4096, 320), (1024, 640), (256, 1280)):
        lin("proj", S, C, C, 15, res=True, fp32=True)     # conv_output / block outputs: fp32 residual stream
        lin("tok", S, C, C, 10, res=True, half=True)      # out_proj x2: IEEE-half token stream
        lin("qk", S, C, 2 * C, 5)
        lin("geglu1", S, C, 4 * C, 5)
        lin("geglu2", S, 4 * C, C, 5, res=True)
        attn("self", S, S, C // 8, 5)
        attn("cross", S, 77, C // 8, 5)
    conv("res", 64, 320, 0, 320, 4)
    conv("mergedskip640", 64, 320, 0, 320, 2, cx=640, f16=True)   # decoders 10/11: conv_merged + 1x1 skip of 640 channels
    conv("merged", 64, 320, 0, 320, 5, res=True)      # conv_merged: fp32 residual in, fp32 + bf16 out
    conv("res", 32, 640, 0, 640, 3)
    conv("merged", 32, 640, 0, 640, 5, res=True)
    conv("res", 16, 1280, 0, 1280, 3)
    conv("merged", 16, 1280, 0, 1280, 6, res=True)
    conv("res", 8, 1280, 0, 1280, 4)
    conv("merged", 8, 1280, 0, 1280, 7, res=True)
    conv("cat", 8, 1280, 1280, 1280, 3)
    conv("cat", 16, 1280, 1280, 1280, 2)
    conv("cat", 16, 1280, 640, 1280, 1)
    conv("cat", 32, 1280, 640, 640, 1)
    conv("cat", 32, 640, 640, 640, 1)
    conv("cat", 32, 640, 320, 640, 1)
    conv("cat", 64, 640, 320, 320, 1)
    conv("cat", 64, 320, 320, 320, 2)
    conv("down", 64, 320, 0, 320, 1, kind=ops.GEMM_CONV3X3_S2)
    conv("down", 32, 640, 0, 640, 1, kind=ops.GEMM_CONV3X3_S2)
    conv("down", 16, 1280, 0, 1280, 1, kind=ops.GEMM_CONV3X3_S2)
    # VAE decoder tail (per-sample sizes; batch B)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--json", default="")
    ap.add_argument("--pair", type=int, default=0, help="0 auto, 1 single-CTA tiles, 2 CTA pairs")
    ap.add_argument("--graph", action="store_true", help="time a CUDA-graph replay of the launches")
    ap.add_argument("--attn-mode", type=int, default=0, help="0 plain, 1 ones-row denominator, 2 + f16x2 exps, 3 ones-row + pre-scaled q + qk_fold")
    args = ap.parse_args()
    global PAIR, ATTN_MODE
    PAIR = args.pair
    ATTN_MODE = args.attn_mode
    torch.manual_seed(0)
    scratch = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
    rows = []
    tot_ms = 0.0
    tot_fl = 0.0
    for name, count, fl, fn in cases(args.batch):
        if args.only and args.only not in name:
            continue
        for _ in range(args.warmup):
            fn()
        scratch.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.graph:
            # the iters launches replayed from a CUDA graph: no host launch cost between kernels (what the
            # captured denoising loop sees); outputs are allocated by fn() inside the capture pool
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(args.iters):
                    fn()
            g.replay()
            scratch.zero_()
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
        else:
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        tf = fl / (ms * 1e-3) / 1e12
        rows.append({"case": name, "count_per_unet_eval": count, "ms": ms, "tflops": tf})
        tot_ms += ms * count
        tot_fl += fl * count
        print(f"{name:44s} x{count:<3d} {ms * 1e3:9.1f} us  {tf:8.1f} TFLOP/s", flush=True)
    fault = _ext.read_fault()
    print(f"weighted total: {tot_ms:.2f} ms per UNet evaluation in these kernels, "
          f"{tot_fl / max(tot_ms, 1e-9) / 1e9:.1f} TFLOP/s aggregate; fault={fault}")
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
