"""LightweightDecoder parameter tree (src/models/decoder.py:21-123): 4 x [UpsampleBlock, 2 x ResidualBlock] + head."""
from __future__ import annotations

import logging
from typing import List

import torch.nn as nn

from .blocks import ResidualBlock, UpsampleBlock

logger = logging.getLogger(__name__)


class LightweightDecoder(nn.Module):
    def __init__(self, in_channels: int = 768, channels: List[int] = (384, 192, 96, 48), out_channels: int = 13,
                 num_residual_blocks: int = 2, use_pixel_shuffle: bool = False):
        super().__init__()
        channels = list(channels)
        self.in_channels, self.out_channels, self.num_stages = in_channels, out_channels, len(channels)
        self.channels, self.num_residual_blocks = channels, num_residual_blocks
        layers, prev = [], in_channels
        for ch in channels:
            layers.append(UpsampleBlock(prev, ch, use_pixel_shuffle=use_pixel_shuffle))
            layers.extend(ResidualBlock(ch) for _ in range(num_residual_blocks))
            prev = ch
        self.upsample_blocks = nn.Sequential(*layers)
        self.head = nn.Sequential(nn.Conv2d(channels[-1], channels[-1], 3, padding=1), nn.BatchNorm2d(channels[-1]),
                                  nn.GELU(), nn.Conv2d(channels[-1], out_channels, kernel_size=1))
        self.apply(self._init_weights)
        logger.info("Initialized LightweightDecoder: %s parameters", f"{self.count_parameters():,}")

    @staticmethod
    def _init_weights(m):
        """Kaiming-normal (fan_out) convs, unit BatchNorm (decoder.py:91-99)."""
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    def count_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def forward(self, x):  # pragma: no cover - holder only
        raise RuntimeError("LightweightDecoder holds parameters only; call SatMAERestoration.forward")
