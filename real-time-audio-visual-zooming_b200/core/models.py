"""Mask estimators of the reference, as torch modules with the reference's parameter names so that its
checkpoints (`mask_3.pth`, `mask_estimator.pth`) load unchanged.  The networks run on cuDNN through torch; only the
feature extraction in front of them and the MVDR behind them are this project's kernels.

FreqPreservingUNet: rt_av_zoom/core/full_audio_generating_pipeline/inference.py:29-67 (1.84 M parameters).
ResBlock / DeepFPU: rt_av_zoom/core/resnet_model_mvdr/inference.py:38-137 (16.05 M parameters).
Both map (B, 2, F, T) float32 -> (B, F, T) in (0, 1); time is pooled by 2 three / four times, frequency never.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _double_conv(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(),
                         nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU())


def _fit(x: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Nearest-neighbour resize when the transposed conv overshoots an odd time length."""
    if x.shape[3] != like.shape[3]:
        x = F.interpolate(x, size=like.shape[2:], mode="nearest")
    return x


class FreqPreservingUNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.pool = nn.MaxPool2d(kernel_size=(1, 2))
        widths = [(2, 32), (32, 64), (64, 128)]
        self.enc1, self.enc2, self.enc3 = (_double_conv(i, o) for i, o in widths)
        self.bot = _double_conv(128, 256)
        self.up3 = nn.ConvTranspose2d(256, 128, (1, 2), stride=(1, 2))
        self.dec3 = _double_conv(256, 128)
        self.up2 = nn.ConvTranspose2d(128, 64, (1, 2), stride=(1, 2))
        self.dec2 = _double_conv(128, 64)
        self.up1 = nn.ConvTranspose2d(64, 32, (1, 2), stride=(1, 2))
        self.dec1 = _double_conv(64, 32)
        self.out = nn.Sequential(nn.Conv2d(32, 1, 1), nn.Sigmoid())

    def forward(self, x):
        skips = []
        for enc in (self.enc1, self.enc2, self.enc3):
            x = enc(x)
            skips.append(x)
            x = self.pool(x)
        x = self.bot(x)
        for up, dec in ((self.up3, self.dec3), (self.up2, self.dec2), (self.up1, self.dec1)):
            skip = skips.pop()
            x = dec(torch.cat([_fit(up(x), skip), skip], dim=1))
        return self.out(x).squeeze(1)


class ResBlock(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(channels, channels, 3, padding=1), nn.BatchNorm2d(channels), nn.ReLU(),
                                  nn.Conv2d(channels, channels, 3, padding=1), nn.BatchNorm2d(channels))
        self.relu = nn.ReLU()

    def forward(self, x):
        return self.relu(x + self.conv(x))


def _res_stage(cin: int, cout: int, n_res: int = 1) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(),
                         *[ResBlock(cout) for _ in range(n_res)])


class DeepFPU(nn.Module):
    def __init__(self):
        super().__init__()
        self.pool = nn.MaxPool2d(kernel_size=(1, 2))
        self.enc1_conv = _double_conv(2, 32)
        self.enc2_conv = _res_stage(32, 64)
        self.enc3_conv = _res_stage(64, 128)
        self.enc4_conv = _res_stage(128, 256)
        self.bottleneck = _res_stage(256, 512, n_res=2)
        self.up4 = nn.ConvTranspose2d(512, 256, (1, 2), stride=(1, 2))
        self.dec4_conv = _res_stage(512, 256)
        self.up3 = nn.ConvTranspose2d(256, 128, (1, 2), stride=(1, 2))
        self.dec3_conv = _res_stage(256, 128)
        self.up2 = nn.ConvTranspose2d(128, 64, (1, 2), stride=(1, 2))
        self.dec2_conv = _res_stage(128, 64)
        self.up1 = nn.ConvTranspose2d(64, 32, (1, 2), stride=(1, 2))
        self.dec1_conv = _double_conv(64, 32)
        self.out = nn.Sequential(nn.Conv2d(32, 1, 1), nn.Sigmoid())

    def forward(self, x):
        skips = []
        for enc in (self.enc1_conv, self.enc2_conv, self.enc3_conv, self.enc4_conv):
            x = enc(x)
            skips.append(x)
            x = self.pool(x)
        x = self.bottleneck(x)
        for up, dec in ((self.up4, self.dec4_conv), (self.up3, self.dec3_conv), (self.up2, self.dec2_conv),
                        (self.up1, self.dec1_conv)):
            skip = skips.pop()
            x = dec(torch.cat([_fit(up(x), skip), skip], dim=1))
        return self.out(x).squeeze(1)
