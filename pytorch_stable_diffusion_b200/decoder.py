"""VAE decoder with the reference's classes and state_dict keys (sd/decoder.py:7-374)."""
import torch
from torch import nn

from . import engine, ops
from .attention import SelfAttention, _Packed, _require_cuda


class VAE_AttentionBlock(nn.Module, _Packed):
    def __init__(self, channels: int):
        super().__init__()
        # declared but never applied by the reference's forward (sd/decoder.py:31 vs :34-73); kept so
        # that the state_dict keys match
        self.groupnorm = nn.GroupNorm(32, channels)
        self.attention = SelfAttention(1, channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(N, C, H, W) -> (N, C, H, W), reproducing sd/decoder.py:34-73 as written (no GroupNorm;
        attention output re-viewed raw as (n, c, h, w) before the residual add)."""
        _require_cuda(x, "VAE_AttentionBlock")
        pk = self._packed(lambda m, dev: engine.pack_self_attention(m.attention, dev))
        xn = engine.Stream(ops.nchw_to_nhwc(x.to(torch.float32), out_fp32=True))
        return ops.nhwc_to_nchw_f32(engine.run_vae_attn(pk, xn).f)


class VAE_ResidualBlock(nn.Module, _Packed):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.groupnorm_1 = nn.GroupNorm(32, in_channels)
        self.conv_1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.groupnorm_2 = nn.GroupNorm(32, out_channels)
        self.conv_2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
        if in_channels != out_channels:
            self.residual_layer = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        else:
            self.residual_layer = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """sd/decoder.py:135-189."""
        _require_cuda(x, "VAE_ResidualBlock")
        pk = self._packed(lambda m, dev: engine.pack_resblock(m, dev, time=False))
        xn = engine.Stream(ops.nchw_to_nhwc(x.to(torch.float32), out_fp32=True))
        return ops.nhwc_to_nchw_f32(engine.run_resblock(pk, xn).f)


class VAE_Decoder(nn.Sequential, engine.EngineCache):
    _engine_cls = engine.VAEDecoderEngine

    def __init__(self):
        super().__init__(
            nn.Conv2d(4, 4, kernel_size=1),
            nn.Conv2d(4, 512, kernel_size=3, padding=1),
            VAE_ResidualBlock(512, 512),
            VAE_AttentionBlock(512),
            VAE_ResidualBlock(512, 512),
            VAE_ResidualBlock(512, 512),
            VAE_ResidualBlock(512, 512),
            VAE_ResidualBlock(512, 512),
            nn.Upsample(scale_factor=2),
            nn.Conv2d(512, 512, kernel_size=3, padding=1),
            VAE_ResidualBlock(512, 512),
            VAE_ResidualBlock(512, 512),
            VAE_ResidualBlock(512, 512),
            nn.Upsample(scale_factor=2),
            nn.Conv2d(512, 512, kernel_size=3, padding=1),
            VAE_ResidualBlock(512, 256),
            VAE_ResidualBlock(256, 256),
            VAE_ResidualBlock(256, 256),
            nn.Upsample(scale_factor=2),
            nn.Conv2d(256, 256, kernel_size=3, padding=1),
            VAE_ResidualBlock(256, 128),
            VAE_ResidualBlock(128, 128),
            VAE_ResidualBlock(128, 128),
            nn.GroupNorm(32, 128),
            nn.SiLU(),
            nn.Conv2d(128, 3, kernel_size=3, padding=1),
        )

    def forward(self, x):
        """(B, 4, h, w) -> (B, 3, 8h, 8w) fp32 (sd/decoder.py:342-374). Unlike the reference (:364) the
        caller's tensor is not scaled in place."""
        _require_cuda(x, "VAE_Decoder")
        img = self._engine().forward_nhwc(x.to(torch.float32).contiguous())
        return ops.nhwc_to_nchw_f32(img)

    def decode_nhwc(self, x):
        """(B, 4, h, w) latents -> fp32 NHWC image [B, 8h, 8w, 3] (the layout pipeline.generate needs)."""
        _require_cuda(x, "VAE_Decoder")
        return self._engine().forward_nhwc(x.to(torch.float32).contiguous())
