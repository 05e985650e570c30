"""Analysis / synthesis / hyper transforms: the 5x5 family of /root/reference/Components.py.

Same class names, constructor arguments and ``net.<i>`` parameter names as the reference
(Encoder5x5 :6-18, Decoder5x5 :35-47, HyperEncoder5x5 :65-75, HyperDecoder5x5 :94-105), so
``state_dict`` round-trips.  ``forward`` does not execute the ``nn.Sequential``: each
conv (+ the GDN / IGDN / LeakyReLU after it) is one ``nic_conv_fwd`` call with the activation fused
into the epilogue, activations staying NHWC between layers.

The 3x3 residual family (Encoder3x3 ... and Layers.py) is outside the hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine
from ._lib import EPI_BIAS, EPI_GDN, EPI_IGDN, EPI_LRELU
from .gdn import GDN


def _ops_from_sequential(net: nn.Sequential):
    """Pair every conv with the activation that follows it."""
    ops, mods, i = [], list(net), 0
    while i < len(mods):
        conv = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        if isinstance(nxt, GDN):
            ops.append(engine.ConvOp(conv, EPI_IGDN if nxt.inverse else EPI_GDN, gdn=nxt)); i += 2
        elif isinstance(nxt, nn.LeakyReLU):
            ops.append(engine.ConvOp(conv, EPI_LRELU)); i += 2
        else:
            ops.append(engine.ConvOp(conv, EPI_BIAS)); i += 1
    return ops


class _Transform(nn.Module):
    precision = None   # None -> engine.resolve_precision (fp32 for a stand-alone transform); set by the owning model

    def _build_ops(self):
        self._ops = _ops_from_sequential(self.net)

    @property
    def ops(self):
        return self._ops

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # stand-alone call of one transform: the fp32 arm unless the owner asked for a tensor-core arm explicitly
        return engine.run_sequential_nchw(self._ops, x, engine.resolve_precision(self.precision))


def _conv(cin, cout, k, s):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=k // 2)


def _deconv(cin, cout):
    return nn.ConvTranspose2d(cin, cout, kernel_size=5, stride=2, padding=2, output_padding=1)


class Encoder5x5(_Transform):
    """g_a: four 5x5 stride-2 convs, GDN after the first three (Components.py:9-17)."""

    def __init__(self, latent_channels: int = 192):
        super().__init__()
        m, layers, cin = latent_channels, [], 3
        for i in range(4):
            layers.append(_conv(cin, m, 5, 2))
            if i < 3:
                layers.append(GDN(m, beta_min=1e-6, gamma_init=.1))
            cin = m
        self.net = nn.Sequential(*layers)
        self._build_ops()


class Decoder5x5(_Transform):
    """g_s: four 5x5 stride-2 transposed convs, IGDN after the first three (Components.py:38-46)."""

    def __init__(self, latent_channels: int = 192):
        super().__init__()
        m, layers = latent_channels, []
        for i in range(4):
            layers.append(_deconv(m, m if i < 3 else 3))
            if i < 3:
                layers.append(GDN(m, inverse=True, beta_min=1e-6, gamma_init=.1))
        self.net = nn.Sequential(*layers)
        self._build_ops()


class HyperEncoder5x5(_Transform):
    """h_a: 3x3 s1, 5x5 s2, 5x5 s2 with LeakyReLU(0.01) between (Components.py:68-74)."""

    def __init__(self, latent_channels: int = 192):
        super().__init__()
        m = latent_channels
        self.net = nn.Sequential(_conv(m, m, 3, 1), nn.LeakyReLU(inplace=True), _conv(m, m, 5, 2),
                                 nn.LeakyReLU(inplace=True), _conv(m, m, 5, 2))
        self._build_ops()


class HyperDecoder5x5(_Transform):
    """h_s: two 5x5 s2 transposed convs and a 3x3, LeakyReLU(0.01) between (Components.py:98-104)."""

    def __init__(self, latent_channels: int = 192):
        super().__init__()
        m = latent_channels
        self.net = nn.Sequential(_deconv(m, m), nn.LeakyReLU(inplace=True), _deconv(m, int(1.5 * m)),
                                 nn.LeakyReLU(inplace=True), _conv(int(1.5 * m), 2 * m, 3, 1))
        self._build_ops()
        mid = int(1.5 * m)
        if mid % 64 and mid > 64:       # e.g. m = 192: 288 channels travel as 320 on the tensor-core arms (engine.ConvOp)
            pad = (mid + 63) // 64 * 64
            self._ops[1].tc_pad_cout, self._ops[2].tc_pad_cin = pad, pad
