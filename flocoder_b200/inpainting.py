"""Host side of the inpainting variant: what sits in FRONT of the mask-conditioned U-Net.

``MaskEncoder`` (``flocoder/inpainting.py:180-245``) turns a pixel-space mask ``[B,1,P,P]`` into the latent-shaped
conditioning tensor ``cond['mask_cond']`` ``[B,4,P/16,P/16]`` that :class:`flocoder_b200.unet.Unet` (``mask_cond=True``)
consumes; ``mask_blending`` (``inpainting.py:250-257``) forms the ODE start point.  Both run ONCE per batch, before the ODE
loop -- they are callers of the hot path, not part of it -- so they stay ordinary PyTorch modules here (same parameter
names / shapes / construction order as the reference, so its checkpoints and seeded inits load unchanged).  The mask
branches INSIDE the U-Net (the 5x5 / 3x3 fusion convolutions, the bilinear resizes, the four residual fusions,
``unet.py:214-235,298-305,336-340,360-364``) run in the CUDA library (``flo_unet_set_mask`` + the fp32 op program).
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn.functional as F
from torch import nn

__all__ = ["MaskEncoder", "mask_blending"]


class _DownsampleBlock(nn.Module):
    """``inpainting.py:159-177``: strided conv + 3x3 conv (SiLU after each) next to a hard (pooled) copy of channel 0."""

    def __init__(self, in_channels, out_channels, shrink_fac=4, mode="pool"):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, shrink_fac, stride=shrink_fac)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        if mode == "pool":
            self.hard_shrink = nn.AvgPool2d(kernel_size=shrink_fac, stride=shrink_fac)
        else:
            self.hard_shrink = partial(F.interpolate, scale_factor=1.0 / shrink_fac, mode="bilinear")

    def forward(self, x):
        skip = self.hard_shrink(x[:, 0:1])
        learned = F.silu(self.conv2(F.silu(self.conv1(x))))
        return torch.cat([skip, learned], dim=1)


class MaskEncoder(nn.Module):
    """Pixel-space mask -> mask latents (``inpainting.py:180-245``): channel 0 is the doubly-shrunk mask itself, the
    other ``output_channels - 1`` are learned features squashed by ``final_act`` (default sigmoid)."""

    def __init__(self, output_channels=4, shrink_fac=4, mode="pool", final_act=torch.sigmoid):
        super().__init__()
        self.layers = nn.Sequential(
            _DownsampleBlock(1, 16, shrink_fac, mode),       # 1 -> 17 channels
            _DownsampleBlock(17, 32, shrink_fac, mode),      # 17 -> 33 channels
            nn.Conv2d(33, output_channels - 1, 1),
        )
        self.final_act = final_act
        if mode == "pool":
            self.double_shrink = nn.AvgPool2d(kernel_size=shrink_fac ** 2, stride=shrink_fac ** 2)
        else:
            self.double_shrink = partial(F.interpolate, scale_factor=1.0 / (shrink_fac ** 2), mode="bilinear")

    def forward(self, mask_pixels):
        if mask_pixels.dtype in (torch.uint8, torch.int32, torch.int64, torch.bool):
            mask_pixels = mask_pixels.float()
        learned = self.layers(mask_pixels)
        if self.final_act is not None:
            learned = self.final_act(learned)
        return torch.cat([self.double_shrink(mask_pixels), learned], dim=1)


def mask_blending(source, mask, noise=None):
    """``inpainting.py:250-257``: noise where the mask is 1, the source elsewhere."""
    if noise is None:
        noise = torch.randn_like(source)
    return source + mask * (noise - source)
