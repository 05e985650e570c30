"""Data-parallel evaluation: the images of a batch are sharded across ranks (one process per GPU).

The forward pass has no cross-image operation (no batch norm; GDN is per pixel), so there is no
data-path collective.  The only exchange is the one the rate-distortion terms need
(RateDistortionLoss.py:19-31 take means over the WHOLE batch, and PSNR is the log of the batch-mean
MSE): each rank contributes its per-image (bits_y, bits_z, mse) and one all-gather over
NCCL / NVLink rebuilds the global vectors, after which every rank evaluates the same formulas in
the same order.  Works with any torch.distributed backend (gloo on CPU in the tests).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist


def gather_per_image(per_image_local: torch.Tensor, group=None) -> torch.Tensor:
    """[3, B_local] on every rank -> [3, B_global] (rank-major image order) on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return per_image_local
    world = dist.get_world_size(group)
    local = per_image_local.contiguous()
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.reshape(-1), group=group)
    return out.reshape(world, local.shape[0], -1).permute(1, 0, 2).reshape(local.shape[0], -1)


def rd_terms_from_per_image(per_image: torch.Tensor, num_pixels: int, lambda_rd: float) -> dict:
    """RateDistortionLoss.py:19-34 from per-image (bits_y, bits_z, mse) rows; tensors stay on device."""
    bits_y, bits_z, mse_i = per_image[0].double(), per_image[1].double(), per_image[2].double()
    bpp_y, bpp_z = (bits_y / num_pixels).mean(), (bits_z / num_pixels).mean()
    mse = mse_i.mean()
    bpp_total = bpp_y + bpp_z
    return {
        "loss": (bpp_total + lambda_rd * (255 ** 2) * mse).float(),
        "bpp_y": bpp_y.float(), "bpp_z": bpp_z.float(), "bpp_total": bpp_total.float(), "mse": mse.float(),
        "psnr": (-10 * torch.log10(mse + 1e-8)).float(),
        "mse_per_image": per_image[2], "psnr_per_image": -10 * torch.log10(per_image[2] + 1e-8),
        "bits_y": bits_y.mean().float(), "bits_z": bits_z.mean().float(), "bits_total": (bits_y + bits_z).mean().float(),
    }


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous shard of the batch for `rank`; the batch must divide evenly ("replicas only" otherwise)."""
    b = x.shape[0]
    if b % world:
        raise ValueError(f"batch {b} does not divide over {world} ranks")
    per = b // world
    return x[rank * per:(rank + 1) * per]


class ShardedEvaluator:
    """model(x_local) + rd terms on this rank's images, then one all-gather for the global terms."""

    def __init__(self, model, lambda_rd: float, group=None, lean: bool = True):
        self.model, self.lambda_rd, self.group, self.lean = model, lambda_rd, group, lean

    @torch.no_grad()
    def step(self, x_local: torch.Tensor):
        from .RateDistortionLoss import rd_terms
        out = self.model(x_local, training=False, lean=self.lean)
        per_image, _ = rd_terms(out, x_local, self.lambda_rd)
        per_image = gather_per_image(per_image, self.group)
        return out, rd_terms_from_per_image(per_image, x_local.shape[2] * x_local.shape[3], self.lambda_rd)
