"""VAE encoder with the reference's class and state_dict keys (sd/encoder.py:8-155)."""
import torch
from torch import nn

from . import engine, ops
from .attention import _require_cuda
from .decoder import VAE_AttentionBlock, VAE_ResidualBlock


class VAE_Encoder(nn.Sequential, engine.EngineCache):
    _engine_cls = engine.VAEEncoderEngine

    def __init__(self):
        super().__init__(
            nn.Conv2d(3, 128, kernel_size=3, padding=1),
            VAE_ResidualBlock(128, 128),
            VAE_ResidualBlock(128, 128),
            nn.Conv2d(128, 128, kernel_size=3, stride=2, padding=0),
            VAE_ResidualBlock(128, 256),
            VAE_ResidualBlock(256, 256),
            nn.Conv2d(256, 256, kernel_size=3, stride=2, padding=0),
            VAE_ResidualBlock(256, 512),
            VAE_ResidualBlock(512, 512),
            nn.Conv2d(512, 512, kernel_size=3, stride=2, padding=0),
            VAE_ResidualBlock(512, 512),
            VAE_ResidualBlock(512, 512),
            VAE_ResidualBlock(512, 512),
            VAE_AttentionBlock(512),
            VAE_ResidualBlock(512, 512),
            nn.GroupNorm(32, 512),
            nn.SiLU(),
            nn.Conv2d(512, 8, kernel_size=3, padding=1),
            nn.Conv2d(8, 8, kernel_size=1, padding=0),
        )

    def forward(self, x: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """x (B, 3, H, W) in [-1, 1], noise (B, 4, H/8, W/8) -> latents (B, 4, H/8, W/8)
        (sd/encoder.py:95-155: right/bottom pad before the stride-2 convs, reparameterisation,
        x0.18215)."""
        _require_cuda(x, "VAE_Encoder")
        xn = ops.nchw_to_nhwc(x.to(torch.float32), out_fp32=True)
        return self._engine().forward_from_nhwc(xn, noise.to(device=x.device, dtype=torch.float32))
