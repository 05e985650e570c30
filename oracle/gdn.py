"""Restatement of ``compressai.layers.gdn.GDN`` (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED AT THIS BOUNDARY.  The reference imports GDN from the PyPI
package ``compressai`` (call sites /root/reference/Components.py:2, 11, 13, 15,
40, 42, 44; requirements.txt:9 names it without a version) and that package is
neither vendored under /root/reference nor installed in this image, so no run of
the real third-party code is available to check against.  What follows restates
the arithmetic CompressAI has published unchanged through 1.1 - 1.2.x
(``compressai/layers/gdn.py`` and ``compressai/ops/parametrizers.py``):

  NonNegativeParametrizer(minimum, reparam_offset = 2**-18):
      pedestal = reparam_offset ** 2                      (= 2**-36)
      bound    = sqrt(minimum + pedestal)
      init(v)  = sqrt(max(v + pedestal, pedestal))
      forward(p) = max(p, bound) ** 2 - pedestal          (LowerBound autograd)
  GDN(C, inverse, beta_min = 1e-6, gamma_init = 0.1):
      beta  = beta_reparam.init(ones(C));  gamma = gamma_reparam.init(gamma_init * eye(C))
      norm  = conv2d(x ** 2, gamma_eff[C, C, 1, 1], beta_eff)
      out   = x * (sqrt(norm) if inverse else rsqrt(norm))

The state_dict names it registers (beta, gamma, {beta,gamma}_reparam.pedestal,
{beta,gamma}_reparam.lower_bound.bound) are the ones SURVEY.md §2.3 lists.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

REPARAM_OFFSET = 2.0 ** -18
PEDESTAL = REPARAM_OFFSET ** 2


class LowerBoundFunction(torch.autograd.Function):
    """``compressai.ops.bound_ops.LowerBoundFunction``: forward max(x, bound); the gradient passes where x >= bound OR where it
    would push x upwards (grad_output < 0), so a parameter sitting below the bound can still move back above it."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through = (x >= bound) | (grad_output < 0)
        return pass_through.to(grad_output.dtype) * grad_output, None


def lower_bound(x, bound: float):
    return LowerBoundFunction.apply(x, torch.tensor([float(bound)], dtype=x.dtype, device=x.device))


def gdn_effective(beta_raw, gamma_raw, beta_min: float = 1e-6):
    """(beta_eff, gamma_eff) from the stored (reparametrised) parameters (differentiable, LowerBound gradient rule)."""
    beta_bound = (beta_min + PEDESTAL) ** 0.5
    gamma_bound = (0.0 + PEDESTAL) ** 0.5
    beta = lower_bound(beta_raw, beta_bound) ** 2 - PEDESTAL
    gamma = lower_bound(gamma_raw, gamma_bound) ** 2 - PEDESTAL
    return beta, gamma


class _LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.tensor([float(bound)]))

    def forward(self, x):
        return LowerBoundFunction.apply(x, self.bound)


class _NonNegative(nn.Module):
    def __init__(self, minimum: float = 0.0):
        super().__init__()
        self.register_buffer("pedestal", torch.tensor([PEDESTAL]))
        self.lower_bound = _LowerBound((minimum + PEDESTAL) ** 0.5)

    def init(self, v):
        return torch.sqrt(torch.max(v + self.pedestal, self.pedestal))

    def forward(self, p):
        return self.lower_bound(p) ** 2 - self.pedestal


class GDN(nn.Module):
    """Stand-in installed as ``compressai.layers.gdn.GDN`` when the reference is imported by
    oracle/make_golden.py (so the reference's own Components.py builds and runs here)."""

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = _NonNegative(minimum=beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = _NonNegative()
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))

    def forward(self, x):
        C = x.size(1)
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm
