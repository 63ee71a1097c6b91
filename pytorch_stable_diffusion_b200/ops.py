"""Thin torch-tensor wrappers over the C ABI (include/sdb200.h).

PyTorch is used for device memory and streams only; every function here launches one (or, for
GroupNorm and split-K GEMMs, two) hand-written kernels on the current CUDA stream.
"""
import ctypes
import math

import torch

from . import _ext
from ._ext import AttnArgs, GemmArgs

GEMM_LINEAR, GEMM_CONV3X3_S1, GEMM_CONV3X3_S2, GEMM_CONV3X3_S2_PAD_RB, GEMM_CONV2X2_UP = 0, 1, 2, 3, 4
ACT_NONE, ACT_QUICK_GELU, ACT_SILU = 0, 1, 2
NUM_SMS = 148


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class LaunchProfiler:
    """Brackets every C-ABI kernel call with CUDA events on the launching stream and records its
    algorithmic work (bench.py's roofline pass; eager execution only, never under graph capture)."""

    def __init__(self):
        self.records = []

    def begin(self, name, flops=0.0, nbytes=0.0, shape=None, relaunch=None):
        """relaunch: optional zero-argument callable that issues the identical launch again (same device
        buffers, same arguments) - bench.py re-times the dominant launch inside a CUDA graph with it."""
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self.records.append([name, float(flops), float(nbytes), e0, e1, shape, relaunch])
        return e1

    def summary(self, by_shape=False):
        """{name (or (name, shape)): launches, ms, flops, bytes} over the recorded launches."""
        torch.cuda.synchronize()
        out = {}
        for name, flops, nbytes, e0, e1, shape, relaunch in self.records:
            key = (name, shape) if by_shape else name
            d = out.setdefault(key, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


PROFILER = None


def _prof(name, flops=0.0, nbytes=0.0, shape=None, relaunch=None):
    return PROFILER.begin(name, flops, nbytes, shape, relaunch) if PROFILER is not None else None


def _prof_end(ev):
    if ev is not None:
        ev.record()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t


import os as _os
_UP_SEPARATE = _os.environ.get("SDB_UP_SEPARATE") == "1"   # A/B switch: four launches per up-sampling conv (one per phase)
_NO_EPI_W64 = _os.environ.get("SDB_NO_EPI_W64") == "1"   # A/B switch (also read by the library): 32-column stores only
_NO_WIDE = _os.environ.get("SDB_NO_WIDE") == "1"     # A/B switch: never pick the wide (320-column) tiles
_A3_CHOOSER = _os.environ.get("SDB_A3_CHOOSER") != "0"   # A/B switch: the tile chooser ignores filter-column staging


def _choose_tiling(rows, cout, nkb, out_bytes=2, res_bytes=0, a3_bytes=0):
    """(block_n, nsplit) from a per-SM cycle model of gemm_tc_kernel, calibrated on B200 timelines
    (tools/gemm_trace.py, tools/micro/mma_rate.cu):

      * one tcgen05.mma (M=128 per CTA, K=16) costs max(88, N/2) cycles;
      * the TMA engine of an SM delivers ~63 B/clk: a k-block of a narrow tile moves 16 KiB of A plus
        N/cg rows of W, which makes 256 x 160 pair tiles TMA-bound (414 instead of 354 cycles);
      * narrow tiles (two TMEM buffers) overlap the epilogue with the next tile but pay a ~3000-cycle
        pipeline refill at every tile start; wide tiles (block_n = 320: two 160-column accumulators sharing
        the A tile, one buffer) are MMA-bound but their epilogue is exposed;
      * an epilogue moves its bytes at ~20 B/clk per SM (HBM share).

      * a3_bytes > 0: the conv runs with filter-column staging when it is not split (sdb_gemm_conv_a3_bytes): a
        k-block then moves a3_bytes / 3 of A instead of 16 KiB.

    Tiles are 256 rows (CTA pairs, 74 slots on a 148-SM part) or 128 rows when there is one row tile."""
    m_tiles = (rows + 127) // 128
    cg = 2 if m_tiles >= 2 else 1
    row_tiles = (m_tiles + cg - 1) // cg
    slots = NUM_SMS // cg
    if cout % 160 == 0:
        # the UNet's channel counts (320 / 640 / 1280): 160-column tiles waste nothing and were tuned on
        # the timelines; 256-column tiles lose to them through padding and wave quantisation
        cands = [160]
    else:
        cands = [bn for bn in (256, 192, 160, 128, 96, 64) if bn <= cout]
        small = ((cout + 15) // 16) * 16
        if small <= 256 and small not in cands:
            cands.append(small)
    if cout % 320 == 0 and not _NO_WIDE:
        cands.append(320)
    # 16-bit result only (no fp32 tensor, no fp32 residual) and a short reduction: the TMA epilogue stores 64-column
    # units when the tile boundaries allow it (sdb_gemm_args: epi_w64) - 256-column tiles then beat the 160-column ones
    # from 640 output channels up (tools/epi16_probe.py: 65536 x 320 x 1280 94 -> 79 us, x 640 50 -> 47 us)
    w64 = out_bytes == 2 and res_bytes in (0, 2) and nkb <= 32 and not _NO_EPI_W64
    if w64 and cout % 160 == 0 and cout >= 640 and 256 not in cands:
        cands.append(256)
    best = None
    for bn in cands:
        n_tiles = (cout + bn - 1) // bn
        for ns in (1, 2, 3, 4, 6, 8):
            if ns > 1 and nkb < 16 * ns:
                continue
            per = (nkb + ns - 1) // ns
            tiles = row_tiles * n_tiles * ((nkb + per - 1) // per)
            rounds = (tiles + slots - 1) // slots
            epi_bytes = 128.0 * bn * ((4 if ns > 1 else out_bytes) + (0 if ns > 1 else res_bytes))
            epi = epi_bytes / 20.0 + (1000.0 if (w64 and bn % 64 == 0 and ns == 1) else 1500.0)
            if bn == 320:
                tile = per * 8 * 88.0 + 1500.0 + epi
            else:
                a_kb = a3_bytes / 3.0 if (a3_bytes and ns == 1 and nkb > 32) else 16384.0
                tma = (a_kb + (bn // cg) * 128.0) / 63.0
                mma = 4.0 * max(88.0, bn / 2.0)
                tile = max(per * max(tma, mma) + 3000.0, epi)
            cost = rounds * tile
            if ns > 1:                                 # fp32 partials written + read back by the finalize kernel
                cost += 6000.0 + rows * cout * 4.0 * (ns + 1) / (NUM_SMS * 20.0)
            if best is None or cost < best[0]:
                best = (cost, bn, ns)
    return best[1], best[2]


def gemm(a0, w, cout, *, kind=GEMM_LINEAR, a1=None, bias=None, residual=None, act=ACT_NONE,
         out=None, out_fp32=False, out2=None, bias_per_row=False, M=None, conv_dims=None, c0=None, c1=0,
         lda0=0, lda1=0, ldw=0, ldo=0, ldr=0, block_n=0, nsplit=0, cta_pair=0, out_f16=False, epi_mode=0,
         gn_samples=None, ax0=None, ax1=None, out16=None, up_phase=None, gn_part=None):
    """out = act(A . W^T + bias) + residual through sdb_gemm_tc. See include/sdb200.h.

    gn_samples=N: the output is a GroupNorm input of N samples - the epilogue also writes per-slab, per-channel
    partial statistics and the call returns (out, out2 or None, part or None); part is None when the geometry or
    the tiling (split-K) cannot provide them and the consumer has to run its own statistics pass.

    The 16-bit operands (a0, a1, ax0, ax1, w) are all bf16 or all IEEE half (sdb_gemm_args::ab_f16 follows a0's
    dtype). out16 = torch.bfloat16 | torch.float16: type of the 16-bit tensor written (`out`, or `out2` next to an
    fp32 `out`); default bf16."""
    lib = _ext.lib()
    op16 = a0.dtype if a0.dtype == torch.float16 else torch.bfloat16
    _chk(a0, op16, "a0")
    _chk(w, op16, "w")
    for t_, nm_ in ((a1, "a1"), (ax0, "ax0"), (ax1, "ax1")):
        if t_ is not None:
            _chk(t_, op16, nm_)
    if out16 is None:
        out16 = torch.float16 if out_f16 else torch.bfloat16
    out_f16 = out16 == torch.float16
    args = GemmArgs()
    args.ab_f16 = 1 if op16 == torch.float16 else 0
    args.kind = kind
    args.a0, args.a1, args.w = _p(a0), _p(a1), _p(w)
    args.bias = _p(_chk(bias, torch.float32, "bias")) if bias is not None else None
    if residual is not None:
        if residual.dtype not in (torch.bfloat16, torch.float32, torch.float16):
            raise ValueError("residual must be bf16, fp32 or IEEE half")
        args.residual = _p(residual)
        args.res_fp32 = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}[residual.dtype]
    if kind == GEMM_LINEAR:
        if M is None:
            M = a0.shape[0] if a0.dim() == 2 else a0.numel() // a0.shape[-1]
        if c0 is None:
            c0 = a0.shape[-1]
        args.M = M
        rows = M
        ntaps = 1
        m_tiles = (M + 127) // 128
    else:
        nb, hi, wi = conv_dims
        if c0 is None:
            c0 = a0.shape[-1]
        args.NB, args.HI, args.WI = nb, hi, wi
        s2 = kind in (GEMM_CONV3X3_S2, GEMM_CONV3X3_S2_PAD_RB)
        ho, wo = (hi // 2, wi // 2) if s2 else (hi, wi)
        rows = nb * ho * wo                      # CONV2X2_UP: rows of ONE phase (the low-resolution grid)
        ntaps = 4 if kind == GEMM_CONV2X2_UP else 9
        m_tiles = (rows + 127) // 128
        if kind == GEMM_CONV2X2_UP:
            if out is None or up_phase is None:
                raise ValueError("CONV2X2_UP writes one phase of a caller-allocated [NB, 2H, 2W, Cout] output")
            args.up_phase = up_phase
            nsplit = 1
    args.C0, args.C1, args.Cout = c0, c1, cout
    args.lda0, args.lda1, args.ldw, args.ldo, args.ldr = lda0, lda1, ldw, ldo, ldr
    if out is None:
        out = torch.empty((rows, cout), device=a0.device, dtype=torch.float32 if out_fp32 else out16)
    args.out = _p(out)
    args.out_fp32 = 1 if out_fp32 else 0
    if out2 is True:
        out2 = torch.empty(out.shape, device=out.device, dtype=out16)
    args.out2 = _p(out2)
    args.bias_per_row = 1 if bias_per_row else 0
    args.act = act
    nkb = ntaps * ((c0 + 63) // 64 + (c1 + 63) // 64)
    cx0 = cx1 = 0
    if ax0 is not None:        # extra 1x1 source (resblock skip conv) accumulated behind the nine taps
        cx0 = ax0.shape[-1]
        cx1 = ax1.shape[-1] if ax1 is not None else 0
        args.ax0, args.ax1 = _p(ax0), _p(ax1)
        args.Cx0, args.Cx1 = cx0, cx1
        nkb += (cx0 + cx1) // 64
    if block_n == 0 or nsplit == 0:
        ob = (4 if out_fp32 else 2) + (2 if out2 is not None else 0)
        rb = 0 if residual is None else residual.element_size()
        a3b = lib.sdb_gemm_conv_a3_bytes(*conv_dims) if (kind in (GEMM_CONV3X3_S1, GEMM_CONV2X2_UP) and _A3_CHOOSER) else 0
        bn_auto, ns_auto = _choose_tiling(rows, cout, nkb, ob, rb, a3b)
        if block_n == 0:
            block_n = bn_auto
            if nsplit == 0:
                nsplit = ns_auto
        elif nsplit == 0:
            nsplit = 1
    args.block_n = block_n
    ws = None
    if nsplit > 1:
        ws = torch.empty((nsplit, rows, cout), device=a0.device, dtype=torch.float32)
        args.workspace = _p(ws)
    args.nsplit = nsplit
    args.cta_pair = cta_pair
    args.out_f16 = 1 if out_f16 else 0
    args.epi_mode = epi_mode
    part = None
    if gn_part is not None:                      # shared partial-sum tensor of the four up-sampling phases
        part = gn_part
        args.gn_part = _p(part)
    elif gn_samples is not None and (out_fp32 or kind != GEMM_LINEAR) and nsplit == 1 and cout % 32 == 0 \
            and block_n % 32 == 0 and ldo == 0:
        hw = rows // gn_samples
        k_slabs = lib.sdb_gemm_gn_slabs(kind, args.NB, args.HI, args.WI, args.M, hw if kind == GEMM_LINEAR else 0)
        if k_slabs > 0 and rows % gn_samples == 0:
            part = torch.empty((gn_samples, k_slabs, cout, 2), device=a0.device, dtype=torch.float32)
            args.gn_part = _p(part)
            args.gn_hw = hw
    ev = None
    if PROFILER is not None:
        # one key per launch VARIANT: everything that changes the FLOPs or the bytes of the launch is in it, so the
        # per-shape averages bench.py reports are averages over identical launches
        k_total = ntaps * (c0 + c1) + cx0 + cx1
        nph = 4 if (kind == GEMM_CONV2X2_UP and up_phase == 4) else 1       # phases computed by this launch
        res_b = residual.element_size() if residual is not None else 0
        shape = (f"rows={rows} cin={c0 + c1} cout={cout} taps={ntaps} cx={cx0 + cx1} "
                 f"out={'f32' if out_fp32 else ('f16' if out_f16 else 'bf16')} res={str(residual.dtype).replace('torch.', '').replace('float', 'f') if res_b else 'none'} "
                 f"out2={1 if out2 is not None else 0} gn={1 if part is not None else 0} bn={block_n} split={nsplit} "
                 f"ab={'f16' if op16 == torch.float16 else 'bf16'}"
                 + (f" hw={rows // conv_dims[0]}" if conv_dims is not None else "") + (" phases=4" if nph == 4 else ""))
        keep = (a0, a1, w, bias, residual, out, out2, ws, part, ax0, ax1)     # the relaunch closure owns its operands
        ev = _prof("gemm_tc_conv3x3" if ntaps != 1 else "gemm_tc_linear",
                   2.0 * rows * cout * k_total * nph,
                   2.0 * (rows * (c0 + c1 + cx0 + cx1) + cout * k_total * nph) + out.numel() * out.element_size()
                   + rows * cout * res_b + (rows * cout * 2 if out2 is not None else 0),
                   shape=shape,
                   relaunch=lambda: (keep, _ext.check(lib.sdb_gemm_tc(ctypes.byref(args), _stream()), "sdb_gemm_tc"))[1])
    _ext.check(lib.sdb_gemm_tc(ctypes.byref(args), _stream()), "sdb_gemm_tc")
    _prof_end(ev)
    if gn_samples is not None:
        return out, out2, part
    return (out, out2) if out2 is not None else out


def linear(x, w, bias=None, **kw):
    """x: bf16 [M, K] (row stride may exceed K), w: bf16 [Cout, K]."""
    lda = x.stride(0) if x.dim() == 2 and x.stride(0) != x.shape[1] else 0
    return gemm(x, w, w.shape[0], kind=GEMM_LINEAR, bias=bias, M=x.shape[0], c0=x.shape[1],
                lda0=lda, **kw)


def conv3x3(x, w, cout, bias=None, kind=GEMM_CONV3X3_S1, **kw):
    """x: bf16 NHWC [N, H, W, C]; w: bf16 [Cout, 9*C] packed (ky, kx, c). Returns [N, Ho, Wo, Cout]."""
    n, h, wd, c = x.shape
    out = gemm(x, w, cout, kind=kind, bias=bias, conv_dims=(n, h, wd), c0=c, **kw)
    s2 = kind != GEMM_CONV3X3_S1
    shape = (n, h // 2 if s2 else h, wd // 2 if s2 else wd, cout)
    if kw.get("gn_samples") is not None:
        return out[0].view(shape), (out[1].view(shape) if out[1] is not None else None), out[2]
    if isinstance(out, tuple):
        return out[0].view(shape), out[1].view(shape)
    return out.view(shape)


def conv_up2x(x, w4, cout, bias=None, *, out2=False, gn_samples=None, out16=torch.bfloat16, block_n=0):
    """conv3x3(nearest_upsample_x2(x)) without the up-sampled tensor (Upsample, sd/diffusion.py:412-435): four parity
    phases, each a 2x2 convolution of the low-resolution x [N, H, W, C] with pre-summed taps (w4: [4][Cout][4*C],
    engine.pack_upsample_phases). Returns (out fp32 [N, 2H, 2W, Cout], out2 16-bit or None, GroupNorm partials or None)."""
    lib = _ext.lib()
    n, h, wd, c = x.shape
    out = torch.empty((n, 2 * h, 2 * wd, cout), device=x.device, dtype=torch.float32)
    o2 = torch.empty(out.shape, device=x.device, dtype=out16) if out2 else None
    part = None
    if gn_samples is not None and cout % 32 == 0:
        k = lib.sdb_gemm_gn_slabs(GEMM_CONV2X2_UP, n, h, wd, 0, 0)
        if k > 0:
            part = torch.empty((n, 4 * k, cout, 2), device=x.device, dtype=torch.float32)
    if _UP_SEPARATE:                             # A/B switch: one launch per phase
        for phase in range(4):
            gemm(x, w4[phase], cout, kind=GEMM_CONV2X2_UP, bias=bias, conv_dims=(n, h, wd), c0=c, out=out,
                 out_fp32=True, out2=o2, out16=out16, up_phase=phase, gn_part=part, block_n=block_n)
    else:
        # up_phase = 4: ONE launch, the tile index carries the phase (4x the tiles: the 8x8 -> 16x16 layer of a batch of
        # 16 has 32 tiles per phase for 74 CTA-pair slots)
        gemm(x, w4.view(4 * cout, -1), cout, kind=GEMM_CONV2X2_UP, bias=bias, conv_dims=(n, h, wd), c0=c, out=out,
             out_fp32=True, out2=o2, out16=out16, up_phase=4, gn_part=part, block_n=block_n)
    return out, o2, part


def attention(q, k, vt, out, *, NB, heads, d, S, Skv, Skv_pad, ldq, ldk, ldo, causal=False, vt_ld=0,
              variant=0, sum_row=False, p_f16=False, exp_poly=0, q_prescaled=False, qk_cols=0, qk_fold=False):
    lib = _ext.lib()
    a = AttnArgs()
    a.q, a.k, a.vt, a.out = _p(_chk(q, torch.bfloat16, "q")), _p(_chk(k, torch.bfloat16, "k")), \
        _p(_chk(vt, torch.float16 if p_f16 else torch.bfloat16, "vt")), _p(_chk(out, torch.bfloat16, "out"))
    a.NB, a.heads, a.d, a.S, a.Skv, a.Skv_pad, a.vt_ld = NB, heads, d, S, Skv, Skv_pad, vt_ld
    a.ldq, a.ldk, a.ldo = ldq, ldk, ldo
    a.causal = 1 if causal else 0
    a.scale = 1.0 / math.sqrt(d)
    a.variant = variant
    a.sum_row = 1 if sum_row else 0
    a.p_f16 = 1 if p_f16 else 0
    a.exp_poly = exp_poly
    a.q_prescaled = 1 if q_prescaled else 0
    a.qk_cols = qk_cols
    a.qk_fold = 1 if qk_fold else 0
    ev = None
    if PROFILER is not None:
        keep = (q, k, vt, out)
        # algorithmic bytes: Q and O once, K and V once per sample and head (re-reads by the query tiles hit L2)
        ev = _prof("attention", 4.0 * NB * heads * S * Skv * d * (0.5 if causal else 1.0),
                   2.0 * NB * heads * d * (2 * S + 2 * Skv),
                   shape=f"NB={NB} heads={heads} d={d} S={S} Skv={Skv} causal={int(causal)} fold={int(qk_fold)}",
                   relaunch=lambda: (keep, _ext.check(lib.sdb_attention(ctypes.byref(a), _stream()), "sdb_attention"))[1])
    _ext.check(lib.sdb_attention(ctypes.byref(a), _stream()), "sdb_attention")
    _prof_end(ev)
    return out


_NO_GN_FUSED = _os.environ.get("SDB_NO_GN_FUSED") == "1"     # A/B switch: always stats + apply
# pixels per sample from which a GroupNorm prefers the producers' epilogue statistics + the apply kernel over the
# one-pass kernel when both are possible (A/B: SDB_GN_PARTS_MIN_HW=1000000 restores "one-pass wherever it fits")
GN_PARTS_MIN_HW = int(_os.environ.get("SDB_GN_PARTS_MIN_HW", "256"))
_NO_GN_EPI = _os.environ.get("SDB_NO_GN_EPI") == "1"         # A/B switch: ignore epilogue partial statistics


def groupnorm(x0, gamma, beta, *, x1=None, groups=32, eps=1e-5, silu=False, fused=None, part0=None, part1=None,
              out_dtype=torch.bfloat16):
    """GroupNorm (+SiLU) over NHWC x0 ++ x1 (channel concat; each bf16 or fp32); returns bf16 (or IEEE half:
    out_dtype) [N, H, W, C0+C1]. fp32 inputs whose (sample, group slab) fits a cluster's shared memory take the
    one-pass kernel (fused=None: when supported; True: required; False: never)."""
    o16 = 1 if out_dtype == torch.float16 else 0
    lib = _ext.lib()
    kind_of = {torch.float32: 1, torch.bfloat16: 0, torch.float16: 2}      # input kinds of the norm kernels
    f0 = kind_of[x0.dtype]
    f1 = kind_of[x1.dtype] if x1 is not None else 0
    n = x0.shape[0]
    c0 = x0.shape[-1]
    hw = x0.numel() // (n * c0)
    c1 = x1.shape[-1] if x1 is not None else 0
    if fused is None:
        parts_ok = part0 is not None and (x1 is None or part1 is not None) and not _NO_GN_EPI
        fused = (not _NO_GN_FUSED) and f0 == 1 and (x1 is None or f1 == 1) and \
            not (parts_ok and hw >= GN_PARTS_MIN_HW) and lib.sdb_groupnorm_fused_supported(hw, c0, c1, groups) == 2
    out = torch.empty(tuple(x0.shape[:-1]) + (c0 + c1,), device=x0.device, dtype=out_dtype)
    nel0, nel1 = n * hw * c0, n * hw * c1
    in_bytes = nel0 * x0.element_size() + (nel1 * x1.element_size() if x1 is not None else 0)
    if fused:
        _chk(x0, torch.float32, "x0")

        def launch():
            _ext.check(lib.sdb_groupnorm_fused(_p(x0), _p(x1), _p(gamma), _p(beta), _p(out), n, hw, c0, c1, groups,
                                               float(eps), 1 if silu else 0, o16, _stream()), "sdb_groupnorm_fused")
        passes, form = 1.0, "one-pass"
    else:
        stats = torch.empty((lib.sdb_groupnorm_stats_bytes(n, groups) // 8,), device=x0.device, dtype=torch.float64)
        have_parts = part0 is not None and (x1 is None or part1 is not None) and not _NO_GN_EPI

        def launch():
            if have_parts:
                # statistics from the partial sums the producers' epilogues wrote: no pass over the tensor
                _ext.check(lib.sdb_groupnorm_reduce_partials(_p(part0), _p(part1), _p(stats), n, part0.shape[1],
                                                             part1.shape[1] if part1 is not None else 0, c0, c1,
                                                             groups, _stream()), "sdb_groupnorm_reduce_partials")
            else:
                _ext.check(lib.sdb_groupnorm_stats(_p(x0), _p(x1), _p(stats), n, hw, c0, c1, groups, f0, f1, _stream()),
                           "sdb_groupnorm_stats")
            _ext.check(lib.sdb_groupnorm_apply(_p(x0), _p(x1), _p(stats), _p(gamma), _p(beta), _p(out), n, hw,
                                               c0, c1, groups, float(eps), 1 if silu else 0, f0, f1,
                                               1 if have_parts else 0, o16, _stream()), "sdb_groupnorm_apply")
        passes, form = (1.0, "epilogue-stats+apply") if have_parts else (2.0, "stats+apply")
    ev = _prof("groupnorm", 0.0, passes * in_bytes + 2.0 * (nel0 + nel1),
               shape=f"n={n} hw={hw} c0={c0} c1={c1} in={('bf16', 'f32', 'f16')[f0]} silu={int(silu)} {form}",
               relaunch=launch)
    launch()
    _prof_end(ev)
    return out


def layernorm(x, gamma, beta, eps=1e-5, out_fp32=False, out_dtype=torch.bfloat16):
    lib = _ext.lib()
    c = x.shape[-1]
    rows = x.numel() // c
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32 if out_fp32 else out_dtype)
    kind = 1 if out_fp32 else (2 if out_dtype == torch.float16 else 0)
    in_kind = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}[x.dtype]

    def launch():
        _ext.check(lib.sdb_layernorm(_p(x), _p(gamma), _p(beta), _p(out), rows, c, float(eps), in_kind, kind, _stream()),
                   "sdb_layernorm")
    ev = _prof("layernorm", 0.0, x.numel() * x.element_size() + out.numel() * out.element_size(),
               shape=f"rows={rows} c={c} in={('bf16', 'f32', 'f16')[in_kind]} "
                     f"out={'f32' if out_fp32 else 'bf16'}", relaunch=launch)
    launch()
    _prof_end(ev)
    return out


def softmax_rows(scores, scale):
    lib = _ext.lib()
    _chk(scores, torch.float32, "scores")
    rows, cols = scores.shape
    out = torch.empty((rows, cols), device=scores.device, dtype=torch.bfloat16)
    _ext.check(lib.sdb_softmax_rows(_p(scores), _p(out), rows, cols, float(scale), _stream()),
               "sdb_softmax_rows")
    return out


def nchw_to_nhwc(x, repeat=1, scale=1.0, out_fp32=False):
    lib = _ext.lib()
    _chk(x, torch.float32, "x")
    x = x.contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n * repeat, h, w, c), device=x.device,
                      dtype=torch.float32 if out_fp32 else torch.bfloat16)
    _ext.check(lib.sdb_nchw_f32_to_nhwc(_p(x), _p(out), n, c, h, w, repeat, float(scale),
                                        1 if out_fp32 else 0, _stream()), "sdb_nchw_f32_to_nhwc")
    return out


def nchw_to_nhwc_bf16(x, repeat=1, scale=1.0):
    return nchw_to_nhwc(x, repeat, scale, False)


def nhwc_to_nchw_f32(x):
    lib = _ext.lib()
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    _ext.check(lib.sdb_nhwc_to_nchw_f32(_p(x), _p(out), n, c, h, w, 1 if x.dtype == torch.float32 else 0,
                                        _stream()), "sdb_nhwc_to_nchw_f32")
    return out


def upsample2x(x):
    lib = _ext.lib()
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=x.dtype)     # any 16-bit type: a pure copy
    _ext.check(lib.sdb_upsample2x_nhwc(_p(x), _p(out), n, h, w, c, _stream()), "sdb_upsample2x_nhwc")
    return out


def conv_direct(x, w, bias, cout, ksize, out_fp32=False, out2=False, out2_dtype=torch.bfloat16):
    """x NHWC (bf16 or fp32) with Cin <= 8; w fp32 [Cout, k*k, Cin]. out2=True also returns a bf16 copy
    of an fp32 output."""
    lib = _ext.lib()
    n, h, wd, cin = x.shape
    out = torch.empty((n, h, wd, cout), device=x.device,
                      dtype=torch.float32 if out_fp32 else torch.bfloat16)
    o2 = torch.empty((n, h, wd, cout), device=x.device, dtype=out2_dtype) if out2 else None
    _ext.check(lib.sdb_conv_direct(_p(x), _p(w), _p(bias), _p(out), _p(o2), n, h, wd, cin, cout, ksize,
                                   1 if out_fp32 else 0, 1 if x.dtype == torch.float32 else 0,
                                   1 if out2_dtype == torch.float16 else 0, _stream()),
               "sdb_conv_direct")
    return (out, o2) if out2 else out


def small_linear(x, w, bias, act_in=ACT_NONE, act_out=ACT_NONE):
    lib = _ext.lib()
    _chk(x, torch.float32, "x")
    r, k = x.shape
    n = w.shape[0]
    out = torch.empty((r, n), device=x.device, dtype=torch.float32)
    _ext.check(lib.sdb_small_linear(_p(x), _p(w), _p(bias), _p(out), r, k, n, act_in, act_out, _stream()),
               "sdb_small_linear")
    return out


def cfg_ddpm_step(latents, eps, noise, coef, step, cfg_scale, do_cfg, next_in, eps_nchw=False):
    lib = _ext.lib()
    n, c, h, w = latents.shape
    _ext.check(lib.sdb_cfg_ddpm_step(_p(latents), _p(eps), _p(noise), _p(coef), step, float(cfg_scale),
                                     1 if do_cfg else 0, _p(next_in), n, c, h, w, 1 if eps_nchw else 0,
                                     1 if (next_in is not None and next_in.dtype == torch.float32) else 0,
                                     _stream()),
               "sdb_cfg_ddpm_step")
    return latents


def vae_attn_scramble_add(y, res):
    """y bf16, res fp32 [N, HW, C] -> (fp32, bf16) outputs."""
    lib = _ext.lib()
    _chk(res, torch.float32, "res")
    n = res.shape[0]
    c = res.shape[-1]
    hw = res.numel() // (n * c)
    out = torch.empty_like(res)
    out2 = torch.empty(res.shape, device=res.device, dtype=torch.bfloat16)
    _ext.check(lib.sdb_vae_attn_scramble_add(_p(y), _p(res), _p(out), _p(out2), n, hw, c, _stream()),
               "sdb_vae_attn_scramble_add")
    return out, out2


def zeros(shape, dtype, device):
    lib = _ext.lib()
    out = torch.empty(shape, device=device, dtype=dtype)
    _ext.check(lib.sdb_fill_zero(_p(out), out.numel() * out.element_size(), _stream()), "sdb_fill_zero")
    return out


def f32_to_bf16(x, dtype=torch.bfloat16):
    """fp32 -> 16-bit shadow (bf16, or IEEE half with dtype=torch.float16)."""
    lib = _ext.lib()
    _chk(x, torch.float32, "x")
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    _ext.check(lib.sdb_f32_to_bf16(_p(x.contiguous()), _p(out), x.numel(), 1 if dtype == torch.float16 else 0, _stream()),
               "sdb_f32_to_bf16")
    return out


def vae_encode_tail(moments, noise):
    lib = _ext.lib()
    n, h, w, _ = moments.shape
    out = torch.empty((n, 4, h, w), device=moments.device, dtype=torch.float32)
    _ext.check(lib.sdb_vae_encode_tail(_p(moments), _p(noise.contiguous()), _p(out), n, h, w, _stream()),
               "sdb_vae_encode_tail")
    return out


def axpby(x, y, a, b):
    lib = _ext.lib()
    out = torch.empty_like(x)
    _ext.check(lib.sdb_axpby(_p(x), _p(y), _p(out), float(a), float(b), x.numel(), _stream()), "sdb_axpby")
    return out


def image_to_uint8(x):
    lib = _ext.lib()
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    _ext.check(lib.sdb_image_to_uint8(_p(x), _p(out), x.numel(), _stream()), "sdb_image_to_uint8")
    return out


def uint8_to_image(x, out_fp32=False):
    lib = _ext.lib()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    _ext.check(lib.sdb_uint8_to_image(_p(x), _p(out), x.numel(), 1 if out_fp32 else 0, _stream()),
               "sdb_uint8_to_image")
    return out


def clip_embed(tokens, table, pos, t_pad):
    lib = _ext.lib()
    nb, t = tokens.shape
    vocab, d = table.shape
    out = torch.empty((nb, t_pad, d), device=tokens.device, dtype=torch.float32)
    _ext.check(lib.sdb_clip_embed(_p(tokens), _p(table), _p(pos), _p(out), nb, t, t_pad, d, vocab, _stream()),
               "sdb_clip_embed")
    return out


def matmul_f64(a, b):
    """Pack-time fp64 product of row-major matrices (fp32 or fp64 in, fp64 out) on the CUDA cores - no library GEMM."""
    lib = _ext.lib()
    for t, nm in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype not in (torch.float32, torch.float64) or not t.is_contiguous():
            raise ValueError(f"{nm} must be a contiguous fp32/fp64 CUDA tensor")
    m, k = a.shape
    k2, n = b.shape
    if k != k2:
        raise ValueError("inner dimensions differ")
    out = torch.empty((m, n), device=a.device, dtype=torch.float64)
    _ext.check(lib.sdb_matmul_f64(_p(a), 1 if a.dtype == torch.float64 else 0, _p(b), 1 if b.dtype == torch.float64 else 0,
                                  _p(out), m, n, k, _stream()), "sdb_matmul_f64")
    return out


def repeat2(x):
    """[B, ...] -> [2B, ...] with both halves equal to x (two stream-ordered device copies): the classifier-free-
    guidance pair of a tensor computed once."""
    lib = _ext.lib()
    x = x.contiguous()
    out = torch.empty((2 * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    nbytes = x.numel() * x.element_size()
    for half in range(2):
        _ext.check(lib.sdb_copy_bytes(ctypes.c_void_p(out.data_ptr() + half * nbytes), _p(x), nbytes, _stream()),
                   "sdb_copy_bytes")
    return out
