"""UNet denoiser with the reference's classes, constructor signatures and state_dict keys
(sd/diffusion.py:8-837). Modules hold fp32 parameters (so model_loader / load_state_dict(strict=True)
work unchanged); forward() executes hand-written sm_100a kernels via engine.py."""
import torch
import torch.nn as nn

from . import engine, ops
from .attention import CrossAttention, SelfAttention, _Packed, _require_cuda


class TimeEmbedding(nn.Module):
    def __init__(self, n_embed: int):
        super().__init__()
        self.linear_1 = nn.Linear(n_embed, 4 * n_embed)
        self.linear_2 = nn.Linear(4 * n_embed, 4 * n_embed)

    def forward(self, x):
        """(R, 320) -> (R, 1280): linear, SiLU, linear (sd/diffusion.py:44-80)."""
        _require_cuda(x, "TimeEmbedding")
        dev = x.device
        w1, b1 = engine.pack_linear(self.linear_1, dev)
        w2, b2 = engine.pack_linear(self.linear_2, dev)
        h = ops.small_linear(x.to(torch.float32).contiguous(), w1, b1, act_out=ops.ACT_SILU)
        return ops.small_linear(h, w2, b2)


class UNET_ResidualBlock(nn.Module, _Packed):
    def __init__(self, in_channels: int, out_channels: int, n_time=1280):
        super().__init__()
        self.groupnorm_feature = nn.GroupNorm(32, in_channels)
        self.conv_feature = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.linear_time = nn.Linear(n_time, out_channels)
        self.groupnorm_merged = nn.GroupNorm(32, out_channels)
        self.conv_merged = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
        if in_channels == out_channels:
            self.residual_layer = nn.Identity()
        else:
            self.residual_layer = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)

    def forward(self, feature, time):
        """feature (N, C_in, H, W), time (1, 1280) -> (N, C_out, H, W) (sd/diffusion.py:145-209)."""
        _require_cuda(feature, "UNET_ResidualBlock")
        if time.shape[0] != 1:
            raise ValueError("the time embedding is batch-independent: expected shape (1, n_time)")
        pk = self._packed(lambda m, dev: engine.pack_resblock(m, dev, time=True))
        tvec = ops.small_linear(time.to(torch.float32).contiguous(), pk.time_w, pk.time_b,
                                act_in=ops.ACT_SILU).view(-1)
        x = engine.Stream(ops.nchw_to_nhwc(feature.to(torch.float32), out_fp32=True))
        return ops.nhwc_to_nchw_f32(engine.run_resblock(pk, x, None, tvec).f)


class UNET_AttentionBlock(nn.Module, _Packed):
    def __init__(self, n_head: int, n_embed: int, d_context=768):
        super().__init__()
        channels = n_head * n_embed
        self.groupnorm = nn.GroupNorm(32, channels, eps=1e-6)
        self.conv_input = nn.Conv2d(channels, channels, kernel_size=1, padding=0)
        self.layernorm_1 = nn.LayerNorm(channels)
        self.attention_1 = SelfAttention(n_head, channels, in_proj_bias=False)
        self.layernorm_2 = nn.LayerNorm(channels)
        self.attention_2 = CrossAttention(n_head, channels, d_context, in_proj_bias=False)
        self.layernorm_3 = nn.LayerNorm(channels)
        self.linear_geglu_1 = nn.Linear(channels, 4 * channels * 2)
        self.linear_geglu_2 = nn.Linear(4 * channels, channels)
        self.conv_output = nn.Conv2d(channels, channels, kernel_size=1, padding=0)

    def forward(self, x, context):
        """x (N, C, H, W), context (N, 77, 768) -> (N, C, H, W) (sd/diffusion.py:271-381)."""
        _require_cuda(x, "UNET_AttentionBlock")
        pk = self._packed(engine.pack_unet_attn)
        n, t, dc = context.shape
        ctx = torch.zeros((n, engine.CTX_PAD, dc), device=x.device, dtype=torch.bfloat16)
        ctx[:, :t] = context.to(torch.bfloat16)
        kv = engine.context_kv(pk, ctx)
        xn = engine.Stream(ops.nchw_to_nhwc(x.to(torch.float32), out_fp32=True))
        return ops.nhwc_to_nchw_f32(engine.run_unet_attn(pk, xn, kv).f)


class Upsample(nn.Module, _Packed):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, padding=1)

    def forward(self, x):
        """nearest x2 then conv3x3 (sd/diffusion.py:412-435)."""
        _require_cuda(x, "Upsample")
        w, b = self._packed(lambda m, dev: engine.pack_conv3x3(m.conv, dev))
        xn = ops.upsample2x(ops.nchw_to_nhwc_bf16(x.to(torch.float32)))
        return ops.nhwc_to_nchw_f32(ops.conv3x3(xn, w, self.conv.out_channels, bias=b, out_fp32=True))


class SwitchSequential(nn.Sequential):
    """Container with the reference's dispatch rule (sd/diffusion.py:458-496). Whole-UNet execution
    goes through engine.UNetEngine; calling a SwitchSequential directly runs its layers one by one."""

    def forward(self, x: torch.Tensor, context: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        for layer in self:
            if isinstance(layer, UNET_AttentionBlock):
                x = layer(x, context)
            elif isinstance(layer, UNET_ResidualBlock):
                x = layer(x, time)
            elif isinstance(layer, nn.Conv2d):
                x = _conv2d_kernel(layer, x)
            else:
                x = layer(x)
        return x


def _conv2d_kernel(conv, x):
    """A bare nn.Conv2d entry (stem / stride-2 downsample) executed by the conv kernels."""
    _require_cuda(x, "Conv2d")
    dev = x.device
    if conv.in_channels <= 8:
        xn = ops.nchw_to_nhwc(x.to(torch.float32), out_fp32=True)
        pk = engine.pack_direct(conv, dev)
        return ops.nhwc_to_nchw_f32(ops.conv_direct(xn, pk.w, pk.b, pk.cout, pk.k, out_fp32=True))
    xn = ops.nchw_to_nhwc_bf16(x.to(torch.float32))
    w, b = engine.pack_conv3x3(conv, dev)
    kind = ops.GEMM_CONV3X3_S2 if conv.stride[0] == 2 else ops.GEMM_CONV3X3_S1
    return ops.nhwc_to_nchw_f32(ops.conv3x3(xn, w, conv.out_channels, bias=b, kind=kind, out_fp32=True))


class UNET(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoders = nn.ModuleList([
            SwitchSequential(nn.Conv2d(4, 320, kernel_size=3, padding=1)),
            SwitchSequential(UNET_ResidualBlock(320, 320), UNET_AttentionBlock(8, 40)),
            SwitchSequential(UNET_ResidualBlock(320, 320), UNET_AttentionBlock(8, 40)),
            SwitchSequential(nn.Conv2d(320, 320, kernel_size=3, stride=2, padding=1)),
            SwitchSequential(UNET_ResidualBlock(320, 640), UNET_AttentionBlock(8, 80)),
            SwitchSequential(UNET_ResidualBlock(640, 640), UNET_AttentionBlock(8, 80)),
            SwitchSequential(nn.Conv2d(640, 640, kernel_size=3, stride=2, padding=1)),
            SwitchSequential(UNET_ResidualBlock(640, 1280), UNET_AttentionBlock(8, 160)),
            SwitchSequential(UNET_ResidualBlock(1280, 1280), UNET_AttentionBlock(8, 160)),
            SwitchSequential(nn.Conv2d(1280, 1280, kernel_size=3, stride=2, padding=1)),
            SwitchSequential(UNET_ResidualBlock(1280, 1280)),
            SwitchSequential(UNET_ResidualBlock(1280, 1280)),
        ])
        self.bottleneck = SwitchSequential(
            UNET_ResidualBlock(1280, 1280),
            UNET_AttentionBlock(8, 160),
            UNET_ResidualBlock(1280, 1280),
        )
        self.decoders = nn.ModuleList([
            SwitchSequential(UNET_ResidualBlock(2560, 1280)),
            SwitchSequential(UNET_ResidualBlock(2560, 1280)),
            SwitchSequential(UNET_ResidualBlock(2560, 1280), Upsample(1280)),
            SwitchSequential(UNET_ResidualBlock(2560, 1280), UNET_AttentionBlock(8, 160)),
            SwitchSequential(UNET_ResidualBlock(2560, 1280), UNET_AttentionBlock(8, 160)),
            SwitchSequential(UNET_ResidualBlock(1920, 1280), UNET_AttentionBlock(8, 160), Upsample(1280)),
            SwitchSequential(UNET_ResidualBlock(1920, 640), UNET_AttentionBlock(8, 80)),
            SwitchSequential(UNET_ResidualBlock(1280, 640), UNET_AttentionBlock(8, 80)),
            SwitchSequential(UNET_ResidualBlock(960, 640), UNET_AttentionBlock(8, 80), Upsample(640)),
            SwitchSequential(UNET_ResidualBlock(960, 320), UNET_AttentionBlock(8, 40)),
            SwitchSequential(UNET_ResidualBlock(640, 320), UNET_AttentionBlock(8, 40)),
            SwitchSequential(UNET_ResidualBlock(640, 320), UNET_AttentionBlock(8, 40)),
        ])

    def forward(self, x, context, time):
        """Layer-by-layer execution with NCHW fp32 between layers (sd/diffusion.py:628-676). The fused
        NHWC path used by Diffusion.forward / pipeline.generate is engine.UNetEngine."""
        skips = []
        for layers in self.encoders:
            x = layers(x, context, time)
            skips.append(x)
        x = self.bottleneck(x, context, time)
        for layers in self.decoders:
            x = torch.cat((x, skips.pop()), dim=1)
            x = layers(x, context, time)
        return x


class UNET_OutputLayer(nn.Module, _Packed):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.groupnorm = nn.GroupNorm(32, in_channels)
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)

    def forward(self, x):
        """GroupNorm, SiLU, conv3x3 (sd/diffusion.py:714-748)."""
        _require_cuda(x, "UNET_OutputLayer")
        gn, (w, b) = self._packed(lambda m, dev: (engine.pack_norm(m.groupnorm, dev),
                                                   engine.pack_conv3x3(m.conv, dev)))
        xn = ops.groupnorm(ops.nchw_to_nhwc(x.to(torch.float32), out_fp32=True), *gn, silu=True)
        return ops.nhwc_to_nchw_f32(ops.conv3x3(xn, w, self.conv.out_channels, bias=b, out_fp32=True))


class Diffusion(nn.Module, engine.EngineCache):
    _engine_cls = engine.UNetEngine

    def __init__(self):
        super().__init__()
        self.time_embedding = TimeEmbedding(320)
        self.unet = UNET()
        self.final = UNET_OutputLayer(320, 4)

    def forward(self, latent: torch.Tensor, context: torch.Tensor, time: torch.Tensor):
        """latent (N, 4, h, w), context (N, 77, 768), time (1, 320) -> (N, 4, h, w) fp32
        (sd/diffusion.py:797-837)."""
        _require_cuda(latent, "Diffusion")
        eng = self._engine()
        if time.shape[0] != 1:
            raise ValueError("the time embedding is batch-independent: expected shape (1, 320)")
        tvec = eng.time_vectors(time.to(device=latent.device, dtype=torch.float32).contiguous())[0]
        # cross-attention K / V^T depend only on the context: cached while the caller keeps passing the
        # same (unmodified) tensor object. The cache holds a reference, so the storage cannot be reused
        # by another tensor behind our back.
        cached = self.__dict__.get("_sdb_ctx")
        if (cached is None or cached[0] is not context or cached[1] != context._version or cached[2] is not eng):
            cached = (context, context._version, eng, eng.context_kv(context))
            self.__dict__["_sdb_ctx"] = cached
        x = ops.nchw_to_nhwc(latent.to(torch.float32), out_fp32=True)
        return ops.nhwc_to_nchw_f32(eng.forward_nhwc(x, tvec, cached[3]))
