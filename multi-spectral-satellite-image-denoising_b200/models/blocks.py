"""Decoder building blocks as parameter holders with the reference's key names (src/models/blocks.py:12-123).
ConvBlock is kept because ``src.models`` exports it (src/models/__init__.py:7-18); the reference never instantiates it."""
from __future__ import annotations

import torch.nn as nn


class ConvBlock(nn.Module):
    """Conv2d -> BatchNorm2d -> GELU holder (blocks.py:12-38)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 use_bn: bool = True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=not use_bn)
        self.bn = nn.BatchNorm2d(out_channels) if use_bn else nn.Identity()
        self.activation = nn.GELU()


class ResidualBlock(nn.Module):
    """x -> GELU(x + BN(conv3x3(GELU(BN(conv3x3(x)))))) holder (blocks.py:66-85): block.{0,1,3,4}."""

    def __init__(self, channels: int):
        super().__init__()
        self.block = nn.Sequential(
            nn.Conv2d(channels, channels, 3, padding=1, bias=False), nn.BatchNorm2d(channels), nn.GELU(),
            nn.Conv2d(channels, channels, 3, padding=1, bias=False), nn.BatchNorm2d(channels))
        self.activation = nn.GELU()


class UpsampleBlock(nn.Module):
    """ConvTranspose2d(k2,s2) -> BN -> GELU holder (blocks.py:88-123): upsample.{0,1}."""

    def __init__(self, in_channels: int, out_channels: int, use_pixel_shuffle: bool = False):
        super().__init__()
        if use_pixel_shuffle:
            raise NotImplementedError("pixel-shuffle upsampling is never enabled by the reference configs")
        self.upsample = nn.Sequential(nn.ConvTranspose2d(in_channels, out_channels, kernel_size=2, stride=2),
                                      nn.BatchNorm2d(out_channels), nn.GELU())
