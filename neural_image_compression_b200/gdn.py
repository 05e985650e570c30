"""GDN / IGDN layer with the parameter and buffer names ``compressai.layers.gdn.GDN`` registers.

The reference imports this layer from the un-vendored ``compressai`` package
(/root/reference/Components.py:2; call sites :11, :13, :15, :40, :42, :44).  Checkpoints of the
reference therefore contain, per layer, ``beta``, ``gamma``, ``beta_reparam.pedestal``,
``beta_reparam.lower_bound.bound``, ``gamma_reparam.pedestal`` and
``gamma_reparam.lower_bound.bound``; this module keeps those keys so they load.
Inside the model the layer never runs on its own: its contraction and rsqrt / sqrt are fused into
the producing convolution (NIC_EPI_GDN / NIC_EPI_IGDN).  Called stand-alone it goes through
``nic_gdn_fwd``.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import LAYOUT_NCHW, PREC_FP32, check, current_stream, ptr

_REPARAM_OFFSET = 2.0 ** -18
_PEDESTAL = _REPARAM_OFFSET ** 2


class _Bound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.tensor([float(bound)]))


class _Reparam(nn.Module):
    """Holds the constants of compressai's NonNegativeParametrizer (pedestal, lower bound)."""

    def __init__(self, minimum: float = 0.0):
        super().__init__()
        self.register_buffer("pedestal", torch.tensor([_PEDESTAL]))
        self.lower_bound = _Bound((minimum + _PEDESTAL) ** 0.5)

    def init(self, v: torch.Tensor) -> torch.Tensor:
        return torch.sqrt(torch.max(v + self.pedestal, self.pedestal))


class GDN(nn.Module):
    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.in_channels = int(in_channels)
        self.inverse = bool(inverse)
        self.beta_min = float(beta_min)
        self.beta_reparam = _Reparam(minimum=beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = _Reparam()
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise _lib.NicError("GDN: expected a CUDA tensor on a B200; this package has no CPU path")
        lib = _lib.load()
        x = x.contiguous().float()
        n, c, h, w = x.shape
        with torch.cuda.device(x.device):
            gamma = torch.empty(c * c, dtype=torch.float32, device=x.device)
            beta = torch.empty(c, dtype=torch.float32, device=x.device)
            check(lib.nic_pack_gdn(c, self.beta_min, ptr(self.beta.detach().float().contiguous()),
                                   ptr(self.gamma.detach().float().contiguous()), ptr(beta), ptr(gamma), PREC_FP32,
                                   current_stream()), "nic_pack_gdn")
            y = torch.empty_like(x)
            check(lib.nic_gdn_fwd(ptr(x), n, c, h, w, LAYOUT_NCHW, int(self.inverse), ptr(gamma), ptr(beta), ptr(y),
                                  current_stream()), "nic_gdn_fwd")
        return y
