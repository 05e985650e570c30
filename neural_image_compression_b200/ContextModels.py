"""Masked autoregressive context model (/root/reference/ContextModels.py).

``MaskedConv2d`` keeps the reference's ``weight`` / ``bias`` parameters and ``mask`` buffer
(:12-16).  The engine packs only the live taps of mask 'A' (12 of 25 for 5x5), so the masked
positions are never multiplied; to stay checkpoint-identical with the reference, whose forward
zeroes them in place (:19), the first forward zeroes them once here too.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine
from ._lib import EPI_BIAS


class MaskedConv2d(nn.Conv2d):
    def __init__(self, mask_type, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if mask_type not in ("A", "B"):
            raise AssertionError("mask_type must be 'A' or 'B'")
        kh, kw = self.kernel_size
        mask = torch.ones_like(self.weight.data)
        mask[:, :, kh // 2, kw // 2 + (mask_type == "B"):] = 0      # 'B' keeps the centre tap (ContextModels.py:15)
        mask[:, :, kh // 2 + 1:] = 0
        self.register_buffer("mask", mask)
        self.precision = None
        self._op = engine.ConvOp(self, EPI_BIAS, mask_a=1 if mask_type == "A" else 2)
        self._masked_version = None

    def apply_mask_(self):
        """The in-place side effect of ContextModels.py:19, done only when the weight changed."""
        key = (self.weight.data_ptr(), self.weight._version)
        if self._masked_version != key:
            with torch.no_grad():
                self.weight.data *= self.mask
            self._masked_version = (self.weight.data_ptr(), self.weight._version)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.apply_mask_()
        return engine.run_sequential_nchw([self._op], x, engine.resolve_precision(self.precision))


class ContextModel(nn.Module):
    def __init__(self, latent_channels: int = 192):
        super().__init__()
        self.masked = MaskedConv2d("A", in_channels=latent_channels, out_channels=2 * latent_channels,
                                   kernel_size=5, stride=1, padding=2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.masked(x)
