"""Execution engine: packs module parameters into kernel-friendly 16-bit device tensors and runs
the four networks as sequences of C-ABI kernel launches over NHWC activations.

Data layout in HBM
  activations   NHWC ([N, H, W, C]; the (B, S, C) token view of the attention blocks is the same memory). The
                residual stream proper (block inputs / outputs, skips, VAE stream, latents, eps, moments, images)
                is fp32; tensor-core operands are 16-bit - IEEE half at the UNet's full-resolution level (TOP_F16),
                bf16 elsewhere; the resblock's hidden tensor (HID_F16) and the attention blocks' token stream
                (TOK_F16) are IEEE half
  conv weights  16-bit [Cout][ky][kx][Cin] (K-major rows of 9*Cin), linear weights 16-bit [out][in]
  attention     QK projection output [tokens, 2C]; V produced transposed [C, N, S] by a swapped GEMM
  biases/affine fp32

Every tensor operation of a forward pass (run_* / forward*) is a kernel from libsdb200.so. What PyTorch (ATen)
still does, none of it in the sampling loop: dtype casts / permutes / concatenations of parameters at pack time,
torch.empty allocations, the context gather / zero-padding per generate() call and the copies into the CUDA graph's
static input buffers. The pack-time weight compositions (GEGLU fold) run on ops.matmul_f64 - no library GEMM.
"""
import math
import os
from types import SimpleNamespace as NS

import torch

from . import ops

CTX_PAD = 80  # 77 CLIP tokens padded to a multiple of 8 (TMA stride alignment)

# The reference's feed-forward is linear_geglu_2(linear_geglu_1(x)[..., :4C]) with NO non-linearity in
# between (the GEGLU gate is computed and dropped, sd/diffusion.py:359-363), i.e. one affine map. With
# FOLD_GEGLU the two weight matrices are composed once at pack time (fp64) into a single C x C matrix:
# W = W2 . W1[:4C], b = W2 . b1[:4C] + b2 (SURVEY.md "Hard parts"; the FLOP numerator in bench.py drops
# by the saved 179 GFLOP per image-step accordingly). Set to False to run the two GEMMs separately.
FOLD_GEGLU = True
# ... and conv_output composed with it (see pack_unet_attn): one dual-source GEMM per attention block
FOLD_FF_OUT = os.environ.get("SDB_NO_FOLD_FF_OUT") != "1"
# channel-changing resblocks: the 1x1 skip convolution rides in conv_merged's GEMM as extra k-blocks
FUSE_SKIP_CONV = os.environ.get("SDB_NO_FUSE_SKIP") != "1"
# The resblock's hidden tensor (conv_feature output + time bias: read ONLY by the following GroupNorm) is stored as IEEE
# half where its GroupNorm statistics come from the fp32 values in conv_feature's epilogue anyway: only the normalised
# value sees the 2^-11 rounding, 4 bytes per element less HBM traffic per resblock. (Round 1 tried bf16 here: 0.115 ms
# per step for 6 % more RMS error - left off; half costs 1/8 of that error.) SDB_HID_FP32=1 restores the fp32 tensor.
HID_F16 = os.environ.get("SDB_HID_FP32") != "1"
# The token stream INSIDE a UNet attention block (t0 = conv_input, t1 = t0 + self-attention, t2 = t1 + cross-attention;
# sd/diffusion.py:306-353) is IEEE half: each tensor is written once and read twice (LayerNorm + the next residual add),
# so fp32 cost 12 bytes per element and block where half costs 6, and t2 is its own 16-bit GEMM operand (no shadow).
# The block's input and output - the UNet's residual stream proper - stay fp32. SDB_TOK_FP32=1 restores fp32.
TOK_F16 = os.environ.get("SDB_TOK_FP32") != "1"
# self-attention of heads <= 112 channels: softmax denominator from a ones row in V^T (see pack_unet_attn)
SUM_ROW_ATTENTION = os.environ.get("SDB_NO_SUM_ROW") != "1"
# ... and the softmax row offset folded into Q.K^T through a ones column of k (heads padded to R columns)
QK_OFFSET_FOLD = os.environ.get("SDB_NO_QK_FOLD") != "1"
# GroupNorm statistics accumulated by the epilogue of the GEMM that produces the tensor (sdb_gemm_args.gn_part)
GN_EPILOGUE_STATS = os.environ.get("SDB_NO_GN_EPI") != "1"
# The UNet's FULL-RESOLUTION level (encoders 0-3, decoders 9-11, the output layer: everything whose tensor-core
# operands are tensors of latent resolution) takes IEEE-half operands instead of bf16: same width, same tcgen05
# kind::f16 rate, fp32 accumulation, but 11 instead of 8 significand bits. tools/diag_layer_budget.py measured that
# this level carries 76 % of the squared operand-rounding error of a UNet evaluation (decoders.11 alone 30 %, the
# output layer 11 %); with it in half the per-evaluation error against the fp32 reference drops from 7-11e-3 (over the
# 1e-2 tolerance for some 768^2 inputs) to ~4e-3. The operands there are GroupNorm / LayerNorm outputs and 16-bit
# shadows of the fp32 stream written with a saturating conversion; the fp32 master tensors are untouched. Every other
# level, the attention kernels (Q, K, V, P, O), CLIP and the VAE stay bf16. SDB_NO_TOP_F16=1 restores bf16 everywhere.
TOP_F16 = os.environ.get("SDB_NO_TOP_F16") != "1"
# Upsample (nearest x2 -> conv3x3, sd/diffusion.py:412-435) as four parity-phase 2x2 convolutions of the LOW-resolution
# tensor: the 3x3 taps that land on the same input pixel are summed at pack time (fp32, rounded once), so the layer costs
# 4/9 of the FLOPs and the 4x up-sampled tensor never exists (ops.conv_up2x, SDB_GEMM_CONV2X2_UP).
FOLD_UPSAMPLE = os.environ.get("SDB_NO_FOLD_UPSAMPLE") != "1"
# Classifier-free guidance feeds the UNet the same latent twice (cond / uncond differ only in the text context): the part
# of the network before the first cross-attention is evaluated once per pair (UNetEngine._shared_cfg_prefix).
SHARE_CFG_PREFIX = os.environ.get("SDB_NO_SHARE_CFG_PREFIX") != "1"
BF16, F16 = torch.bfloat16, torch.float16


# ------------------------------------------------------------------------------------------------
# packing
def _bf16(t, dev, dtype=torch.bfloat16):
    """Weights as a 16-bit tensor-core operand: bf16, or IEEE half for the layers of the full-resolution level."""
    return t.detach().to(device=dev, dtype=dtype).contiguous()


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def pack_conv3x3(conv, dev, dtype=torch.bfloat16):
    w = conv.weight.detach()
    cout = w.shape[0]
    return _bf16(w.permute(0, 2, 3, 1).reshape(cout, -1), dev, dtype), _f32(conv.bias, dev)


def pack_conv1x1(conv, dev, dtype=torch.bfloat16):
    w = conv.weight.detach()
    return _bf16(w.reshape(w.shape[0], w.shape[1]), dev, dtype), _f32(conv.bias, dev)


def pack_upsample_phases(conv, dev, dtype=torch.bfloat16):
    """[4][Cout][4 * Cin] weights of the four parity phases of conv3x3(nearest_upsample_x2(.)): output pixel
    (2y + a, 2x + b) reads input rows (y + a - 1 + u), u = 0, 1, through the filter rows {0} | {1, 2} (a = 0) or
    {0, 1} | {2} (a = 1) - the same along x with b - so each phase is a 2x2 convolution whose taps are sums of the 3x3
    taps (summed in fp32, rounded to the operand type once). Phase index = 2a + b, tap order [u][v][Cin]."""
    w = conv.weight.detach().to(torch.float32)                    # [Cout, Cin, 3, 3]
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    phases = []
    for a in (0, 1):
        for b in (0, 1):
            taps = []
            for u in (0, 1):
                for v in (0, 1):
                    acc = torch.zeros_like(w[:, :, 0, 0])
                    for ky in groups[a][u]:
                        for kx in groups[b][v]:
                            acc = acc + w[:, :, ky, kx]
                    taps.append(acc)                              # [Cout, Cin]
            phases.append(torch.stack(taps, dim=1).reshape(w.shape[0], -1))
    return torch.stack(phases).to(device=dev, dtype=dtype).contiguous(), _f32(conv.bias, dev)


def pack_direct(conv, dev):
    w = conv.weight.detach()
    cout, cin, k, _ = w.shape
    return NS(w=_f32(w.permute(0, 2, 3, 1).reshape(cout, k * k, cin), dev), b=_f32(conv.bias, dev),
              cout=cout, k=k)


def pack_linear(lin, dev):
    return _bf16(lin.weight, dev), (_f32(lin.bias, dev) if lin.bias is not None else None)


def pack_norm(norm, dev):
    return _f32(norm.weight, dev), _f32(norm.bias, dev)


def fingerprint(module):
    """Cheap identity of a module's parameters (storage pointer + in-place version counter)."""
    return tuple((p.data_ptr(), p._version, p.device.index if p.is_cuda else -1) for p in module.parameters())


def pack_resblock(m, dev, time=True, dt=torch.bfloat16):
    """UNET_ResidualBlock (sd/diffusion.py:111-143) or VAE_ResidualBlock (sd/decoder.py:103-133). dt: 16-bit operand
    type of the block's convolutions (weights here, activations in run_resblock)."""
    if time:
        gn1, conv1, gn2, conv2 = m.groupnorm_feature, m.conv_feature, m.groupnorm_merged, m.conv_merged
    else:
        gn1, conv1, gn2, conv2 = m.groupnorm_1, m.conv_1, m.groupnorm_2, m.conv_2
    pk = NS()
    pk.dt = dt
    pk.gn1_w, pk.gn1_b = pack_norm(gn1, dev)
    pk.conv1_w, pk.conv1_b = pack_conv3x3(conv1, dev, dt)
    pk.gn2_w, pk.gn2_b = pack_norm(gn2, dev)
    pk.conv2_w, pk.conv2_b = pack_conv3x3(conv2, dev, dt)
    pk.cin, pk.cout = conv1.in_channels, conv1.out_channels
    pk.conv2x_w = None
    if isinstance(m.residual_layer, torch.nn.Conv2d):
        pk.skip_w, pk.skip_b = pack_conv1x1(m.residual_layer, dev, dt)
        # the 1x1 skip convolution as extra k-blocks of conv_merged: its weights behind the 9 * Cout columns of the
        # 3x3 filter, the two biases added - one GEMM, no fp32 round trip of the skip branch (FUSE_SKIP_CONV)
        if FUSE_SKIP_CONV and pk.cin % 64 == 0 and pk.cout % 64 == 0:
            pk.conv2x_w = torch.cat([pk.conv2_w, pk.skip_w], dim=1).contiguous()
            pk.conv2x_b = (pk.conv2_b + pk.skip_b).contiguous()
    else:
        pk.skip_w = pk.skip_b = None
    if time:
        pk.time_w, tb = pack_linear(m.linear_time, dev)
        pk.time_b = (tb + pk.conv1_b).contiguous()   # conv bias folded into the time projection bias
    return pk


def pack_unet_attn(m, dev, dt=torch.bfloat16):
    """UNET_AttentionBlock (sd/diffusion.py:243-269). dt: 16-bit operand type of conv_input and of the feed-forward /
    conv_output GEMMs (the block's error-relevant operators, tools/diag_layer_budget.py); the projections around the
    two attention kernels stay bf16 like the kernels themselves."""
    pk = NS()
    pk.dt = dt
    c = m.conv_input.in_channels
    pk.c, pk.heads = c, m.attention_1.n_heads
    pk.gn_w, pk.gn_b = pack_norm(m.groupnorm, dev)
    pk.cin_w, pk.cin_b = pack_conv1x1(m.conv_input, dev, dt)
    pk.ln1 = pack_norm(m.layernorm_1, dev)
    w = m.attention_1.in_proj.weight.detach()
    pk.wqk = _bf16(w[:2 * c], dev)
    pk.wv = _bf16(w[2 * c:], dev)
    b = m.attention_1.in_proj.bias
    pk.bqk = _f32(b[:2 * c], dev) if b is not None else None
    pk.bv = _f32(b[2 * c:], dev) if b is not None else None
    # Heads of up to 112 channels run the two-tile attention kernel with the softmax denominator accumulated by
    # the P.V product itself: V^T gets R = round16(d + 1) rows per head, row d all ones (a zero weight row
    # with bias 1), the rest zero. The padding is free: the tensor core pays for N = R either way.
    d = c // pk.heads
    pk.vt_rows = 0
    pk.qk_cols = 0
    if SUM_ROW_ATTENTION and d <= 112:
        r = (d + 1 + 15) // 16 * 16
        wv = torch.zeros((pk.heads, r, c), dtype=torch.float32)
        wv[:, :d] = w[2 * c:].float().cpu().view(pk.heads, d, c)
        bv = torch.zeros((pk.heads, r), dtype=torch.float32)
        if b is not None:
            bv[:, :d] = b[2 * c:].detach().float().cpu().view(pk.heads, d)
        bv[:, d] = 1.0
        pk.wv = _bf16(wv.view(pk.heads * r, c), dev)
        pk.bv = _f32(bv.view(-1), dev)
        pk.vt_rows = r
        # ... and the queries carry log2(e)/sqrt(d) (sd/attention.py:66 divides the scores by sqrt(d_head)): the
        # scores leave the tensor core in log2 units and the softmax argument is a subtraction, not an FFMA
        sl2 = math.log2(math.e) / math.sqrt(d)
        wq = w[:2 * c].detach().float().cpu().clone()
        wq[:c] *= sl2
        bq = None
        if b is not None:
            bq = b[:2 * c].detach().float().cpu().clone()
            bq[:c] *= sl2
        pk.qk_cols = 0
        if QK_OFFSET_FOLD and r <= 64 and d % 8 == 0:
            # ... and the heads of q and k are padded to the same R columns: column d of k is all ones (zero weight
            # row, bias 1), column d of q is zero - the attention kernel writes each row's -round(max) there and the
            # scores of every later key block leave the tensor core relative to the row's reference (qk_fold)
            wp = torch.zeros((2, pk.heads, r, c), dtype=torch.float32)
            wp[:, :, :d] = wq.view(2, pk.heads, d, c)
            bp = torch.zeros((2, pk.heads, r), dtype=torch.float32)
            if bq is not None:
                bp[:, :, :d] = bq.view(2, pk.heads, d)
            bp[1, :, d] = 1.0
            wq, bq = wp.view(2 * pk.heads * r, c), bp.view(-1)
            pk.qk_cols = r
        pk.wqk = _bf16(wq, dev)
        if bq is not None:
            pk.bqk = _f32(bq, dev)
    pk.wo1, pk.bo1 = pack_linear(m.attention_1.out_proj, dev)
    pk.ln2 = pack_norm(m.layernorm_2, dev)
    pk.wq2, pk.bq2 = pack_linear(m.attention_2.q_proj, dev)
    pk.wk2, pk.bk2 = pack_linear(m.attention_2.k_proj, dev)
    pk.wv2, pk.bv2 = pack_linear(m.attention_2.v_proj, dev)
    pk.wo2, pk.bo2 = pack_linear(m.attention_2.out_proj, dev)
    pk.ln3 = pack_norm(m.layernorm_3, dev)
    pk.tok16 = False
    # the GEGLU gate half is dead in the reference (sd/diffusion.py:359-363): keep the first 4C rows only
    if FOLD_GEGLU:
        # composed in fp64 by our own CUDA-core kernel (ops.matmul_f64): no library GEMM anywhere, pack time included
        w1 = _f32(m.linear_geglu_1.weight.detach()[:4 * c], dev)
        b1 = _f32(m.linear_geglu_1.bias.detach()[:4 * c], dev)
        w2 = _f32(m.linear_geglu_2.weight, dev)
        b2 = m.linear_geglu_2.bias.detach().to(device=dev, dtype=torch.float64)
        w21 = ops.matmul_f64(w2, w1)                                        # [C, C] fp64
        b21 = ops.matmul_f64(w2, b1.view(-1, 1)).view(-1) + b2              # [C] fp64
        pk.wg = w21.to(dt).contiguous()
        pk.bg = b21.to(torch.float32).contiguous()
        pk.wg1 = None
        pk.w_ffout = None
        if FOLD_FF_OUT:
            # conv_output is a 1x1 convolution of (feed-forward + its residual t2): one more affine map, so
            #   conv_output(W_g l3 + b_g + t2) = [W_o W_g | W_o] . [l3 ; t2] + (W_o b_g + b_o)
            # - ONE dual-source GEMM over the LayerNorm output and the bf16 shadow of the token stream instead of
            # two GEMMs with a bf16 round trip of t3 in between (sd/diffusion.py:355-381)
            wo = _f32(m.conv_output.weight.detach().reshape(c, c), dev)
            bo = m.conv_output.bias.detach().to(device=dev, dtype=torch.float64)
            # with a half token stream the second source IS the stream: this GEMM's operands are IEEE half at every level
            pk.tok16 = TOK_F16
            pk.w_ffout = torch.cat([ops.matmul_f64(wo, w21), wo.to(torch.float64)], dim=1) \
                .to(torch.float16 if pk.tok16 else dt).contiguous()
            pk.b_ffout = (ops.matmul_f64(wo, b21.view(-1, 1).contiguous()).view(-1) + bo).to(torch.float32).contiguous()
    else:
        pk.wg1 = _bf16(m.linear_geglu_1.weight[:4 * c], dev, dt)
        pk.bg1 = _f32(m.linear_geglu_1.bias[:4 * c], dev)
        pk.wg2, pk.bg2 = _bf16(m.linear_geglu_2.weight, dev, dt), _f32(m.linear_geglu_2.bias, dev)
    pk.cout_w, pk.cout_b = pack_conv1x1(m.conv_output, dev, dt)
    return pk


def pack_self_attention(att, dev):
    """SelfAttention (sd/attention.py:7-25) as used standalone by the VAE and CLIP."""
    pk = NS()
    c = att.in_proj.in_features
    pk.c, pk.heads = c, att.n_heads
    w = att.in_proj.weight.detach()
    b = att.in_proj.bias
    pk.wqk, pk.wv = _bf16(w[:2 * c], dev), _bf16(w[2 * c:], dev)
    pk.bqk = _f32(b[:2 * c], dev) if b is not None else None
    pk.bv = _f32(b[2 * c:], dev) if b is not None else None
    pk.wo, pk.bo = pack_linear(att.out_proj, dev)
    return pk


# ------------------------------------------------------------------------------------------------
# residual stream
class Stream:
    """A residual-stream activation: fp32 NHWC master `f` plus an optional 16-bit shadow `b`.

    What is added to again from block to block (block inputs/outputs, UNet skips, the VAE stream) stays fp32
    so that rounding does not accumulate along the residual additions of a UNet evaluation; the 16-bit shadow
    (bf16, or IEEE half at the full-resolution level) exists only where a tensor-core kernel reads the stream
    directly as its A operand (skip 1x1 convs, down/up-sampling convs, the VAE attention projections). Branch
    tensors (GroupNorm/LayerNorm outputs, Q/K/V, attention outputs) are 16-bit; the tensors that live inside one
    block - the resblock's hidden tensor, the attention block's token stream - are IEEE half (HID_F16, TOK_F16)."""
    __slots__ = ("f", "b", "gp")

    def __init__(self, f, b=None, gp=None):
        # gp: GroupNorm partial statistics of `f` written by the epilogue of the GEMM that produced it
        # (ops.gemm(gn_samples=...)), or None - the consuming GroupNorm then runs its own statistics pass
        self.f, self.b, self.gp = f, b, gp

    @property
    def shape(self):
        return self.f.shape

    def b16(self, dtype=torch.bfloat16):
        """The 16-bit shadow in the operand type the consumer wants (converted from the fp32 master when the producer
        did not write one, or wrote the other type)."""
        if self.b is None or self.b.dtype != dtype:
            self.b = ops.f32_to_bf16(self.f, dtype)
        return self.b

    def bf16(self):
        return self.b16(torch.bfloat16)


def _stream(out, shape):
    """(fp32, bf16[, partials]) tuple or a lone fp32 tensor from a kernel wrapper -> Stream viewed as `shape`."""
    if isinstance(out, tuple):
        return Stream(out[0].view(shape), out[1].view(shape) if out[1] is not None else None,
                      out[2] if len(out) > 2 else None)
    return Stream(out.view(shape))


def _gn_samples(n, hw, c):
    """n when a GroupNorm over (hw, c) per sample wants its statistics from the producer's epilogue, else None
    (the small levels take the one-pass kernel, which reads the tensor once anyway)."""
    if not GN_EPILOGUE_STATS:
        return None
    if hw >= ops.GN_PARTS_MIN_HW:
        # from the 16 x 16 level up the epilogue statistics + the apply kernel beat the one-pass kernel (9.9 - 12.4 us
        # against 14.9 us for 16 x 256 x 1280), and a skip concatenation needs the partial sums of BOTH its sources
        return n
    from . import _ext
    return n if _ext.lib().sdb_groupnorm_fused_supported(hw, c, 0, 32) != 2 else None


# ------------------------------------------------------------------------------------------------
# block runners (Stream in / Stream out)
def run_resblock(pk, x, x1=None, bias1=None, want_b16=False, out16=torch.bfloat16):
    """conv2(silu(GN(conv1(silu(GN(x ++ x1))) + t))) + skip(x ++ x1).  bias1 = conv1 bias (+ time).
    UNET_ResidualBlock (sd/diffusion.py:145-209) / VAE_ResidualBlock (sd/decoder.py:135-189). pk.dt = 16-bit type of
    this block's operands, out16 = type of the output's 16-bit shadow (what its consumer reads)."""
    n, h, w, c0 = x.shape
    c1 = x1.shape[-1] if x1 is not None else 0
    dt = pk.dt
    a = ops.groupnorm(x.f, pk.gn1_w, pk.gn1_b, x1=x1.f if x1 is not None else None, silu=True,
                      part0=x.gp, part1=x1.gp if x1 is not None else None, out_dtype=dt)
    gs = _gn_samples(n, h * w, pk.cout)
    # the hidden tensor is only ever read by GroupNorm, never as a tensor-core operand: IEEE half (HID_F16) or fp32
    hid = ops.conv3x3(a, pk.conv1_w, pk.cout, bias=bias1 if bias1 is not None else pk.conv1_b,
                      out_fp32=not (HID_F16 and gs is not None), gn_samples=gs, out16=torch.float16)
    hid_gp = None
    if gs is not None:
        hid, _, hid_gp = hid
    a2 = ops.groupnorm(hid, pk.gn2_w, pk.gn2_b, silu=True, part0=hid_gp, out_dtype=dt)
    if pk.conv2x_w is not None and c0 % 64 == 0 and c1 % 64 == 0:
        out = ops.conv3x3(a2, pk.conv2x_w, pk.cout, bias=pk.conv2x_b, out_fp32=True, out2=True if want_b16 else None,
                          gn_samples=gs, ax0=x.b16(dt), ax1=x1.b16(dt) if x1 is not None else None, out16=out16)
        if isinstance(out, tuple):
            return Stream(*out)
        return Stream(out)
    if pk.skip_w is None:
        res = x.f.view(-1, c0)
    else:
        res = ops.gemm(x.b16(dt).view(-1, c0), pk.skip_w, pk.cout,
                       a1=x1.b16(dt).view(-1, c1) if x1 is not None else None,
                       M=n * h * w, c0=c0, c1=c1, bias=pk.skip_b, out_fp32=True)
    out = ops.conv3x3(a2, pk.conv2_w, pk.cout, bias=pk.conv2_b, residual=res, out_fp32=True,
                      out2=True if want_b16 else None, gn_samples=gs, out16=out16)
    if isinstance(out, tuple):
        return Stream(*out)
    return Stream(out)


def project_vt(w, bias, x, n, s):
    """V^T = Wv . X^T + bv for the attention kernel: [C, n, s_pad] bf16 with the per-sample key stride
    padded to a multiple of 8 elements (TMA needs 16-byte strides). Returns (vt, s_pad). The common
    case (s % 8 == 0) is one swapped-operand GEMM over all n*s tokens."""
    c = w.shape[0]
    if s % 8 == 0:
        return ops.gemm(w, x, n * s, M=c, c0=x.shape[1], bias=bias, bias_per_row=True), s
    s_pad = (s + 7) // 8 * 8
    vt = ops.zeros((c, n, s_pad), torch.bfloat16, x.device)
    for i in range(n):
        ops.gemm(w, x[i * s:(i + 1) * s], s, M=c, c0=x.shape[1], bias=bias, bias_per_row=True,
                 out=vt[:, i], ldo=n * s_pad)
    return vt, s_pad


def context_kv(pk, ctx_pad):
    """Cross-attention K and V^T of one block from the padded context [N, CTX_PAD, 768] (bf16).
    Constant over the denoising loop, so computed once per prompt (sd/attention.py:194-198)."""
    n = ctx_pad.shape[0]
    flat = ctx_pad.view(n * CTX_PAD, -1)
    k = ops.linear(flat, pk.wk2, bias=pk.bk2)
    vt = ops.gemm(pk.wv2, flat, n * CTX_PAD, M=pk.c, c0=flat.shape[1], bias=pk.bv2, bias_per_row=True)
    return k, vt


def unet_attn_prefix(pk, x):
    """UNET_AttentionBlock.forward up to (and including) the self-attention residual: GroupNorm, conv_input,
    LayerNorm, self-attention, out_proj (sd/diffusion.py:271-326). Nothing here depends on the text context - for a
    classifier-free-guidance pair this half of the block is the same for both members. Returns the token stream
    t1 [n*s, c] (IEEE half with TOK_F16, else fp32)."""
    n, h, w, c = x.shape
    s = h * w
    m = n * s
    d = c // pk.heads
    dev = x.f.device
    a = ops.groupnorm(x.f, pk.gn_w, pk.gn_b, eps=1e-6, silu=False, part0=x.gp, out_dtype=pk.dt)
    tok = pk.tok16
    t0 = ops.linear(a.view(m, c), pk.cin_w, bias=pk.cin_b, out_fp32=not tok, out16=torch.float16)
    l1 = ops.layernorm(t0, *pk.ln1)
    qk = ops.linear(l1, pk.wqk, bias=pk.bqk)
    vt, vt_ld = project_vt(pk.wv, pk.bv, l1, n, s)
    o = torch.empty((m, c), device=dev, dtype=torch.bfloat16)
    cq = pk.heads * pk.qk_cols if pk.qk_cols else c          # columns of the q (and k) part of qk
    fold = pk.qk_cols > 0 and s % 8 == 0 and s > 128
    ops.attention(qk, qk[:, cq:], vt, o, NB=n, heads=pk.heads, d=d, S=s, Skv=s, Skv_pad=s, vt_ld=vt_ld,
                  ldq=2 * cq, ldk=2 * cq, ldo=c, sum_row=pk.vt_rows > 0, q_prescaled=pk.vt_rows > 0,
                  qk_cols=pk.qk_cols, qk_fold=fold)
    return ops.linear(o, pk.wo1, bias=pk.bo1, residual=t0, out_fp32=not tok, out16=torch.float16)


def unet_attn_suffix(pk, x, t1, kv, want_b16=False, out16=torch.bfloat16):
    """The rest of UNET_AttentionBlock.forward (sd/diffusion.py:328-381): cross-attention over the CLIP tokens,
    feed-forward, conv_output, + block input x. t1: token stream after the self-attention, same batch as x."""
    dt = pk.dt
    n, h, w, c = x.shape
    s = h * w
    m = n * s
    d = c // pk.heads
    dev = x.f.device
    l2 = ops.layernorm(t1, *pk.ln2)
    q = ops.linear(l2, pk.wq2, bias=pk.bq2)
    k2, vt2 = kv
    o2 = torch.empty((m, c), device=dev, dtype=torch.bfloat16)
    ops.attention(q, k2, vt2, o2, NB=n, heads=pk.heads, d=d, S=s, Skv=77, Skv_pad=CTX_PAD,
                  ldq=c, ldk=c, ldo=c)
    ff_out = pk.wg1 is None and pk.w_ffout is not None
    tok = pk.tok16 and ff_out
    t2 = ops.linear(o2, pk.wo2, bias=pk.bo2, residual=t1, out_fp32=not tok, out2=True if (ff_out and not tok) else None,
                    out16=torch.float16 if tok else dt)
    # feed-forward: linear_geglu_2(linear_geglu_1(x)[:, :4C]) — gate unused, no GELU
    if ff_out:
        t2, t2_b = (t2, t2) if tok else t2
        l3 = ops.layernorm(t2, *pk.ln3, out_dtype=torch.float16 if tok else dt)
        out = ops.gemm(l3, pk.w_ffout, c, a1=t2_b, M=m, c0=c, c1=c, bias=pk.b_ffout, residual=x.f.view(m, c),
                       out_fp32=True, out2=True if want_b16 else None, gn_samples=_gn_samples(n, s, c), out16=out16)
        return _stream(out, (n, h, w, c))
    l3 = ops.layernorm(t2, *pk.ln3, out_dtype=dt)
    if pk.wg1 is None:
        t3 = ops.linear(l3, pk.wg, bias=pk.bg, residual=t2, out16=dt)   # folded affine map; only conv_output reads it
    else:
        g = ops.linear(l3, pk.wg1, bias=pk.bg1, out16=dt)
        t3 = ops.linear(g, pk.wg2, bias=pk.bg2, residual=t2, out16=dt)
    out = ops.linear(t3, pk.cout_w, bias=pk.cout_b, residual=x.f.view(m, c), out_fp32=True,
                     out2=True if want_b16 else None, gn_samples=_gn_samples(n, s, c), out16=out16)
    return _stream(out, (n, h, w, c))


def run_unet_attn(pk, x, kv, want_b16=False, out16=torch.bfloat16):
    """UNET_AttentionBlock.forward (sd/diffusion.py:271-381); the token stream t0..t2 is IEEE half (TOK_F16) or fp32."""
    return unet_attn_suffix(pk, x, unet_attn_prefix(pk, x), kv, want_b16=want_b16, out16=out16)


def run_vae_attn(pk, x, want_b16=False):
    """VAE_AttentionBlock.forward (sd/decoder.py:34-73): single-head d=C attention over h*w tokens,
    no GroupNorm, output re-viewed raw as (n, c, h, w) before the residual add. d = 512 does not fit
    one TMEM accumulator, so this runs as GEMM -> row softmax -> GEMM per sample."""
    n, h, w, c = x.shape
    s = h * w
    m = n * s
    xf = x.bf16().view(m, c)
    qk = ops.linear(xf, pk.wqk, bias=pk.bqk)                                   # [m, 2c]
    vt = ops.gemm(pk.wv, xf, m, M=c, c0=c, bias=pk.bv, bias_per_row=True)      # [c, m]
    o = torch.empty((m, c), device=xf.device, dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(c // pk.heads)
    for i in range(n):
        q_i = qk[i * s:(i + 1) * s, :c]
        k_i = qk[i * s:(i + 1) * s, c:]
        scores = ops.gemm(q_i, k_i, s, M=s, c0=c, lda0=2 * c, ldw=2 * c, out_fp32=True, nsplit=1)
        probs = ops.softmax_rows(scores, scale)
        ops.gemm(probs, vt[:, i * s:(i + 1) * s], c, M=s, c0=s, ldw=m, out=o[i * s:(i + 1) * s], nsplit=1)
    y = ops.linear(o, pk.wo, bias=pk.bo)
    f, b = ops.vae_attn_scramble_add(y.view(n, s, c), x.f.view(n, s, c))
    return Stream(f.view(n, h, w, c), b.view(n, h, w, c))


def run_clip_layer(pk, x, n, t_pad):
    """CLIPLayer.forward (sd/clip.py:123-176) on fp32 [n*t_pad, 768] rows (rows >= 77 are padding)."""
    c = pk.att.c
    d = c // pk.att.heads
    l1 = ops.layernorm(x, *pk.ln1)
    qk = ops.linear(l1, pk.att.wqk, bias=pk.att.bqk)
    vt = ops.gemm(pk.att.wv, l1, n * t_pad, M=c, c0=c, bias=pk.att.bv, bias_per_row=True)
    o = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    ops.attention(qk, qk[:, c:], vt, o, NB=n, heads=pk.att.heads, d=d, S=t_pad, Skv=77, Skv_pad=t_pad,
                  ldq=2 * c, ldk=2 * c, ldo=c, causal=True)
    x = ops.linear(o, pk.att.wo, bias=pk.att.bo, residual=x, out_fp32=True)
    l2 = ops.layernorm(x, *pk.ln2)
    hdn = ops.linear(l2, pk.w1, bias=pk.b1, act=ops.ACT_QUICK_GELU)
    return ops.linear(hdn, pk.w2, bias=pk.b2, residual=x, out_fp32=True)


def _reads_bf16(kind, pk):
    """Does program entry (kind, pk) read its stream input as a bf16 tensor-core operand?"""
    if kind == "res":
        return pk.skip_w is not None
    return kind in ("up", "upconv", "conv", "conv1", "attn_vae")


# ------------------------------------------------------------------------------------------------
# whole-network engines
class UNetEngine:
    """Diffusion (sd/diffusion.py:751-837) = TimeEmbedding + UNET + UNET_OutputLayer."""

    def __init__(self, diffusion, dev):
        from .diffusion import UNET_AttentionBlock, UNET_ResidualBlock, Upsample
        self.dev = dev
        te = diffusion.time_embedding
        self.t1_w, self.t1_b = pack_linear(te.linear_1, dev)
        self.t2_w, self.t2_b = pack_linear(te.linear_2, dev)
        self.res_blocks = []

        def pack_seq(seq, top=False, out_top=None):
            """Program of one SwitchSequential: entries [kind, packed, want_b16, out16]. top: the block's 16-bit
            operands are tensors of the full-resolution level (IEEE half when TOP_F16); out_top: so is its OUTPUT
            (differs for the down-sampling conv of encoders.3 and the Upsample conv that ends decoders.8)."""
            if out_top is None:
                out_top = top
            dt = F16 if (top and TOP_F16) else BF16
            dt_out = F16 if (out_top and TOP_F16) else BF16
            prog = []
            for layer in seq:
                if isinstance(layer, UNET_ResidualBlock):
                    pk = pack_resblock(layer, dev, time=True, dt=dt)
                    self.res_blocks.append(pk)
                    prog.append(["res", pk, False, dt])
                elif isinstance(layer, UNET_AttentionBlock):
                    prog.append(["attn", pack_unet_attn(layer, dev, dt), False, dt])
                elif isinstance(layer, Upsample):
                    # nearest x2 of the block's (lower-level) output, then a conv whose OUTPUT is one level up
                    w, b = pack_conv3x3(layer.conv, dev, dt)
                    w4 = pack_upsample_phases(layer.conv, dev, dt)[0] if FOLD_UPSAMPLE else None
                    prog.append(["up", NS(w=w, b=b, w4=w4, cout=layer.conv.out_channels, dt=dt), False, dt_out])
                elif isinstance(layer, torch.nn.Conv2d):
                    if layer.in_channels <= 8:
                        prog.append(["direct", pack_direct(layer, dev), False, dt_out])
                    else:
                        w, b = pack_conv3x3(layer, dev, dt)
                        kind = ops.GEMM_CONV3X3_S2 if layer.stride[0] == 2 else ops.GEMM_CONV3X3_S1
                        prog.append(["conv", NS(w=w, b=b, cout=layer.out_channels, kind=kind, dt=dt), False, dt_out])
                else:
                    raise TypeError(f"unexpected layer {type(layer)}")
            return prog

        u = diffusion.unet
        # full-resolution level: encoders 0-2 and decoders 9-11 entirely; encoders.3 (stride-2 conv) reads it and
        # writes the next level; the Upsample conv closing decoders.8 reads the lower level and writes it
        self.encoders = [pack_seq(s, top=i <= 3, out_top=i <= 2) for i, s in enumerate(u.encoders)]
        self.bottleneck = pack_seq(u.bottleneck)
        self.decoders = [pack_seq(s, top=i >= 9, out_top=i >= 8) for i, s in enumerate(u.decoders)]
        # third field of every entry: does a later tensor-core kernel read this output directly?
        flat = [e for prog in self.encoders + [self.bottleneck] + self.decoders for e in prog]
        for cur, nxt in zip(flat, flat[1:]):
            cur[2] = _reads_bf16(nxt[0], nxt[1])
        for prog in self.encoders:        # skips feed the decoders' 1x1 skip convs (sd/diffusion.py:671)
            prog[-1][2] = True
        self.attn_blocks = [pk for kind, pk, _, _ in flat if kind == "attn"]
        # all linear_time projections as one [sum(Cout), 1280] matrix
        offs, off = [], 0
        for pk in self.res_blocks:
            offs.append(off)
            off += pk.cout
        self.time_offsets, self.time_total = offs, off
        self.time_w = torch.cat([pk.time_w for pk in self.res_blocks], 0).contiguous()
        self.time_b = torch.cat([pk.time_b for pk in self.res_blocks], 0).contiguous()
        for pk, o in zip(self.res_blocks, offs):
            pk.time_off = o
            pk.time_w = None
        fin = diffusion.final
        self.fin_dt = F16 if TOP_F16 else BF16
        self.fin_gn = pack_norm(fin.groupnorm, dev)
        self.fin_w, self.fin_b = pack_conv3x3(fin.conv, dev, self.fin_dt)
        self.fin_cout = fin.conv.out_channels

    def time_vectors(self, time):
        """time fp32 [R, 320] -> fp32 [R, sum(Cout)]: per-ResBlock conv_feature bias + linear_time(
        silu(TimeEmbedding(time))) (sd/diffusion.py:64-76,184-194). Batch-independent."""
        h1 = ops.small_linear(time, self.t1_w, self.t1_b, act_out=ops.ACT_SILU)
        temb = ops.small_linear(h1, self.t2_w, self.t2_b)
        return ops.small_linear(temb, self.time_w, self.time_b, act_in=ops.ACT_SILU)

    def context_kv(self, context):
        """context fp32/bf16 [N, 77, 768] -> per-attention-block (K, V^T)."""
        n, t, dc = context.shape
        ctx = torch.zeros((n, CTX_PAD, dc), device=self.dev, dtype=torch.bfloat16)
        ctx[:, :t] = context.to(device=self.dev, dtype=torch.bfloat16)
        return [context_kv(pk, ctx) for pk in self.attn_blocks]

    def _prefix_shareable(self):
        e = self.encoders
        return len(e) > 2 and [k for k, *_ in e[0]] == ["direct"] and [k for k, *_ in e[1]] == ["res", "attn"]

    def _run_seq(self, prog, x, x1, tvec, kv_iter):
        for kind, pk, want, o16 in prog:
            if kind == "res":
                x = run_resblock(pk, x, x1, tvec[pk.time_off:pk.time_off + pk.cout], want_b16=want, out16=o16)
                x1 = None
            elif kind == "attn":
                x = run_unet_attn(pk, x, next(kv_iter), want_b16=want, out16=o16)
            elif kind == "up":
                nn_, hh_, ww_, _ = x.shape
                gs = _gn_samples(nn_, 4 * hh_ * ww_, pk.cout)
                if pk.w4 is not None:
                    x = Stream(*ops.conv_up2x(x.b16(pk.dt), pk.w4, pk.cout, bias=pk.b, out2=bool(want), gn_samples=gs,
                                              out16=o16))
                else:
                    o = ops.conv3x3(ops.upsample2x(x.b16(pk.dt)), pk.w, pk.cout, bias=pk.b, out_fp32=True,
                                    out2=True if want else None, gn_samples=gs, out16=o16)
                    x = Stream(*o) if isinstance(o, tuple) else Stream(o)
            elif kind == "conv":
                nn_, hh_, ww_, _ = x.shape
                s2_ = pk.kind != ops.GEMM_CONV3X3_S1
                o = ops.conv3x3(x.b16(pk.dt), pk.w, pk.cout, bias=pk.b, kind=pk.kind, out_fp32=True,
                                out2=True if want else None,
                                gn_samples=_gn_samples(nn_, (hh_ * ww_) // (4 if s2_ else 1), pk.cout), out16=o16)
                x = Stream(*o) if isinstance(o, tuple) else Stream(o)
            elif kind == "direct":
                o = ops.conv_direct(x, pk.w, pk.b, pk.cout, pk.k, out_fp32=True, out2=want, out2_dtype=o16)
                x = Stream(*o) if isinstance(o, tuple) else Stream(o)
        return x

    def _shared_cfg_prefix(self, x, tvec, kv_iter):
        """encoders.0 and encoders.1 for a batch whose second half repeats the first (classifier-free guidance:
        latents.repeat(2, 1, 1, 1), sd/pipeline.py:221, one time step for all): everything before the first
        cross-attention - the stem conv, the first resblock, GroupNorm / conv_input / LayerNorm / SELF-attention /
        out_proj of the first attention block - sees identical inputs in both halves, so it runs once on B samples
        and is duplicated (three device copies) where the text context first enters. Per step at batch 8: one
        S = 4096 self-attention, two 320 -> 320 convs and their norms for 8 instead of 16 samples."""
        b = x.shape[0] // 2
        s0 = self._run_seq(self.encoders[0], x[:b], None, tvec, kv_iter)           # Stream [B, h, w, 320]
        (k_res, pk_res, want_res, o16_res), (k_att, pk_att, want_att, o16_att) = self.encoders[1]
        r = run_resblock(pk_res, s0, None, tvec[pk_res.time_off:pk_res.time_off + pk_res.cout], want_b16=want_res,
                         out16=o16_res)
        t1 = unet_attn_prefix(pk_att, r)
        r2 = Stream(ops.repeat2(r.f))                                               # block input: residual of conv_output
        x1 = unet_attn_suffix(pk_att, r2, ops.repeat2(t1), next(kv_iter), want_b16=want_att, out16=o16_att)
        skip0 = Stream(ops.repeat2(s0.f), ops.repeat2(s0.b) if s0.b is not None else None,
                       ops.repeat2(s0.gp) if s0.gp is not None else None)
        return skip0, x1

    def forward_nhwc(self, x, tvec, kvs, cfg_pairs=False):
        """x fp32 (or bf16) NHWC [N, h, w, 4]; tvec fp32 [sum(Cout)] (one row of time_vectors); returns eps fp32
        NHWC [N, h, w, 4]. UNET.forward sd/diffusion.py:628-676 without materialising torch.cat.
        cfg_pairs=True: the caller guarantees x[N/2:] == x[:N/2] (the pipeline's CFG batch) - the context-independent
        prefix of the network is then evaluated once (see _shared_cfg_prefix)."""
        kv_iter = iter(kvs)
        skips = []
        encoders = self.encoders
        if cfg_pairs and SHARE_CFG_PREFIX and x.shape[0] % 2 == 0 and self._prefix_shareable():
            skip0, x = self._shared_cfg_prefix(x, tvec, kv_iter)
            skips += [skip0, x]
            encoders = self.encoders[2:]
        for prog in encoders:
            x = self._run_seq(prog, x, None, tvec, kv_iter)
            skips.append(x)
        x = self._run_seq(self.bottleneck, x, None, tvec, kv_iter)
        for prog in self.decoders:
            x = self._run_seq(prog, x, skips.pop(), tvec, kv_iter)
        a = ops.groupnorm(x.f, *self.fin_gn, silu=True, part0=x.gp, out_dtype=self.fin_dt)
        return ops.conv3x3(a, self.fin_w, self.fin_cout, bias=self.fin_b, out_fp32=True)


def _pack_vae_sequential(seq, dev, pad_rb):
    from .decoder import VAE_AttentionBlock, VAE_ResidualBlock
    prog = []
    for layer in seq:
        if isinstance(layer, VAE_ResidualBlock):
            prog.append(["res", pack_resblock(layer, dev, time=False), False])
        elif isinstance(layer, VAE_AttentionBlock):
            prog.append(["attn_vae", pack_self_attention(layer.attention, dev), False])
        elif isinstance(layer, torch.nn.Conv2d):
            if layer.in_channels <= 8:
                prog.append(["direct", pack_direct(layer, dev), False])
            elif layer.kernel_size[0] == 1:
                w, b = pack_conv1x1(layer, dev)
                prog.append(["conv1", NS(w=w, b=b, cout=layer.out_channels), False])
            else:
                w, b = pack_conv3x3(layer, dev)
                kind = ops.GEMM_CONV3X3_S1
                if layer.stride[0] == 2:
                    kind = ops.GEMM_CONV3X3_S2_PAD_RB if pad_rb else ops.GEMM_CONV3X3_S2
                w4 = None
                if (FOLD_UPSAMPLE and kind == ops.GEMM_CONV3X3_S1 and prog and prog[-1][0] == "up"
                        and layer.in_channels % 64 == 0):
                    w4 = pack_upsample_phases(layer, dev)[0]
                prog.append(["conv", NS(w=w, b=b, w4=w4, cout=layer.out_channels, kind=kind), False])
        elif isinstance(layer, torch.nn.Upsample):
            prog.append(["up", None, False])
        elif isinstance(layer, torch.nn.GroupNorm):
            prog.append(["gn", pack_norm(layer, dev), False])
        elif isinstance(layer, torch.nn.SiLU):
            prog.append(["silu", None, False])
        else:
            raise TypeError(f"unexpected layer {type(layer)}")
    # GroupNorm followed by SiLU is one kernel; nn.Upsample(scale_factor=2) followed by a stride-1 3x3 conv
    # (sd/decoder.py:269-273, 295-299, 321-325) is four parity-phase 2x2 convolutions of the low-resolution tensor
    fused = []
    i = 0
    while i < len(prog):
        if prog[i][0] == "gn" and i + 1 < len(prog) and prog[i + 1][0] == "silu":
            fused.append(["gn_silu", prog[i][1], False])
            i += 2
        elif (FOLD_UPSAMPLE and prog[i][0] == "up" and i + 1 < len(prog) and prog[i + 1][0] == "conv"
              and prog[i + 1][1].kind == ops.GEMM_CONV3X3_S1 and prog[i + 1][1].w4 is not None):
            fused.append(["upconv", prog[i + 1][1], False])
            i += 2
        else:
            fused.append(prog[i])
            i += 1
    for cur, nxt in zip(fused, fused[1:]):
        cur[2] = _reads_bf16(nxt[0], nxt[1])
    return fused


def _run_vae_sequential(prog, x):
    """x: bf16 NHWC tensor (network input). Returns the last entry's fp32 NHWC output."""
    for kind, pk, want in prog:
        o2 = True if want else None
        if kind == "res":
            x = run_resblock(pk, x, want_b16=want)
        elif kind == "attn_vae":
            x = run_vae_attn(pk, x)
        elif kind == "direct":
            src = x.f if isinstance(x, Stream) else x          # the direct convolution reads fp32 or bf16
            o = ops.conv_direct(src, pk.w, pk.b, pk.cout, pk.k, out_fp32=True, out2=True)
            x = Stream(*o)
        elif kind == "conv1":
            n, h, w, c = x.shape
            o = ops.linear(x.bf16().view(-1, c), pk.w, bias=pk.b, out_fp32=True, out2=o2)
            x = _stream(o, (n, h, w, pk.cout))
        elif kind == "conv":
            src = x.bf16() if isinstance(x, Stream) else x
            o = ops.conv3x3(src, pk.w, pk.cout, bias=pk.b, kind=pk.kind, out_fp32=True, out2=o2)
            x = Stream(*o) if isinstance(o, tuple) else Stream(o)
        elif kind == "upconv":
            o, ob, _ = ops.conv_up2x(x.bf16(), pk.w4, pk.cout, bias=pk.b, out2=bool(want))
            x = Stream(o, ob)
        elif kind == "up":
            x = ops.upsample2x(x.bf16())        # bf16 tensor: the next entry is a conv reading it
        elif kind == "gn_silu":
            x = ops.groupnorm(x.f, *pk, silu=True)
        elif kind == "gn":
            x = ops.groupnorm(x.f, *pk, silu=False)
        else:
            raise RuntimeError(f"VAE program entry {kind} has no kernel")
    return x.f if isinstance(x, Stream) else x


class VAEDecoderEngine:
    """VAE_Decoder (sd/decoder.py:192-374)."""

    def __init__(self, decoder, dev):
        self.dev = dev
        self.prog = _pack_vae_sequential(decoder, dev, pad_rb=False)

    def forward_nhwc(self, latents):
        """latents fp32 NCHW [B, 4, h, w] -> image fp32 NHWC [B, 8h, 8w, 3] (x / 0.18215 first)."""
        x = ops.nchw_to_nhwc(latents, scale=1.0 / 0.18215, out_fp32=True)
        return _run_vae_sequential(self.prog, x)


class VAEEncoderEngine:
    """VAE_Encoder (sd/encoder.py:8-155)."""

    def __init__(self, encoder, dev):
        self.dev = dev
        self.prog = _pack_vae_sequential(encoder, dev, pad_rb=True)

    def forward_from_nhwc(self, x, noise):
        """x fp32 (or bf16) NHWC [B, H, W, 3]; noise fp32 NCHW [B, 4, H/8, W/8] -> latents fp32 NCHW."""
        moments = _run_vae_sequential(self.prog, x)
        return ops.vae_encode_tail(moments, noise)


class CLIPEngine:
    """CLIP (sd/clip.py:179-261)."""

    def __init__(self, clip, dev):
        self.dev = dev
        self.table = _f32(clip.embedding.token_embedding.weight, dev)
        self.pos = _f32(clip.embedding.position_embedding, dev)
        self.layers = []
        for layer in clip.layers:
            pk = NS()
            pk.ln1 = pack_norm(layer.layernorm_1, dev)
            pk.att = pack_self_attention(layer.attention, dev)
            pk.ln2 = pack_norm(layer.layernorm_2, dev)
            pk.w1, pk.b1 = pack_linear(layer.linear_1, dev)
            pk.w2, pk.b2 = pack_linear(layer.linear_2, dev)
            self.layers.append(pk)
        self.ln = pack_norm(clip.layernorm, dev)

    def forward(self, tokens):
        """tokens int64 [B, 77] -> fp32 [B, 77, 768]."""
        n, t = tokens.shape
        t_pad = (t + 7) // 8 * 8
        x = ops.clip_embed(tokens.to(self.dev).contiguous(), self.table, self.pos, t_pad)
        d = x.shape[-1]
        x = x.view(n * t_pad, d)
        for pk in self.layers:
            x = run_clip_layer(pk, x, n, t_pad)
        out = ops.layernorm(x, *self.ln, out_fp32=True)
        return out.view(n, t_pad, d)[:, :t].contiguous()


class EngineCache:
    """Mixin: lazily built, parameter-fingerprinted engine for a top-level nn.Module."""

    _engine_cls = None

    def _engine(self):
        params = list(self.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__} runs on hand-written CUDA kernels only; move it to a CUDA device "
                "(there is no CPU fallback)")
        fp = fingerprint(self)
        cached = self.__dict__.get("_sdb_engine")
        if cached is None or cached[0] != fp:
            cached = (fp, type(self)._engine_cls(self, dev))
            self.__dict__["_sdb_engine"] = cached
        return cached[1]

    def invalidate_packed(self):
        self.__dict__.pop("_sdb_engine", None)
