"""Entropy models of /root/reference/EntropyModels.py behind the same classes.

``forward`` returns per-element likelihoods already clamped at ``likelihood_lower_bound`` (1e-9),
exactly as ``EntropyModel.forward`` does (:29-31); the arithmetic is in likelihood.cu.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib, engine
from ._lib import Q_PASSTHRU, check, current_stream, ptr

_KERNEL_BOUND = 1e-9     # the bound compiled into the kernels (EntropyModels.py:18)


def gm_likelihood(y: Tensor, raw: Tensor, M: int, K: int, qmode: int, noise: Tensor = None, full: bool = True,
                  want_y_in: bool = True, out: dict = None):
    """nic_gm_likelihood_fwd on NCHW f32 tensors.

    Returns dict(y_in, p, logp, partials[, weights, mus, sigmas | mu, sigma]).  `out`: a dict returned by an earlier call with
    the same arguments - its tensors are written again instead of fresh ones (timing loops: no allocator calls before the launch).
    """
    lib = _lib.load()
    engine.require_cuda(y, "y")
    y = y.contiguous().float()
    raw = raw.contiguous().float()
    b, m, h, w = y.shape
    if m != M or raw.shape[1] != (2 * M if K == 1 else 3 * K * M):
        raise ValueError(f"entropy parameters {tuple(raw.shape)} do not match y {tuple(y.shape)} with M={M}, K={K}")
    dev = y.device
    if noise is not None:
        noise = noise.contiguous().float()
    if out is not None:
        y_in, p, logp, parts = out.get("y_in"), out["p"], out["logp"], out["partials"]
        ws, mus, sgs = (None, out["mu"], out["sigma"]) if K == 1 and full else \
            (out["weights"], out["mus"], out["sigmas"]) if full else (None, None, None)
        if p.shape != y.shape or p.device != dev:
            raise ValueError("gm_likelihood: `out` does not match y")
        with torch.cuda.device(dev):
            check(lib.nic_gm_likelihood_fwd(ptr(y), ptr(raw), ptr(noise), b, m, h * w, K, qmode, ptr(y_in), ptr(p), ptr(logp),
                                            ptr(ws), ptr(mus), ptr(sgs), ptr(parts), current_stream()), "nic_gm_likelihood_fwd")
        return out
    y_in = torch.empty_like(y) if want_y_in else None
    p, logp = torch.empty_like(y), torch.empty_like(y)
    parts = engine.partials(b, dev)
    ws = mus = sgs = None
    if full:
        shape = (b, m, h, w) if K == 1 else (b, K, m, h, w)
        mus = torch.empty(shape, dtype=torch.float32, device=dev)
        sgs = torch.empty_like(mus)
        ws = torch.empty_like(mus) if K > 1 else None
    with torch.cuda.device(dev):
        check(lib.nic_gm_likelihood_fwd(ptr(y), ptr(raw), ptr(noise), b, m, h * w, K, qmode, ptr(y_in), ptr(p), ptr(logp),
                                        ptr(ws), ptr(mus), ptr(sgs), ptr(parts), current_stream()), "nic_gm_likelihood_fwd")
    out = {"y_in": y_in, "p": p, "logp": logp, "partials": parts}
    if full:
        if K == 1:
            out["mu"], out["sigma"] = mus, sgs
        else:
            out["weights"], out["mus"], out["sigmas"] = ws, mus, sgs
    return out


def split_entropy_parameters(raw: Tensor, M: int, K: int):
    """ParametersModels.py:43-64 for stand-alone EntropyParameters calls (cold path)."""
    y = torch.zeros((raw.shape[0], M) + tuple(raw.shape[2:]), dtype=torch.float32, device=raw.device)
    r = gm_likelihood(y, raw, M, K, Q_PASSTHRU, full=True, want_y_in=False)
    return (r["mu"], r["sigma"]) if K == 1 else (r["weights"], r["mus"], r["sigmas"])


class EntropyModel(nn.Module):
    """Interface of EntropyModels.py:11-46."""

    def __init__(self, likelihood_lower_bound: float = 1e-9):
        super().__init__()
        self.likelihood_lower_bound = likelihood_lower_bound

    def _check_bound(self):
        if abs(self.likelihood_lower_bound - _KERNEL_BOUND) > 1e-15:
            raise NotImplementedError("the sm_100a kernels are built for the reference's likelihood bound 1e-9")

    def _likelihood(self, inputs: Tensor, **kwargs) -> Tensor:
        raise NotImplementedError

    def forward(self, inputs: Tensor, **kwargs) -> Tensor:
        self._check_bound()
        return self._likelihood(inputs, **kwargs)      # the kernels clamp at the bound themselves

    def channel_cdf(self, ch: int, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def channel_pmf(self, ch: int, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @property
    def likelihood_bound(self) -> float:
        return self.likelihood_lower_bound


class FactorizedEntropyBottleneck(EntropyModel):
    """Per-channel 1-3-3-3-1 cumulative (EntropyModels.py:49-151); parameters as in :62-86."""

    def __init__(self, channels: int, init_scale: float = 10.0, hidden_dims: Tuple[int, ...] = (3, 3, 3),
                 likelihood_lower_bound: float = 1e-9):
        super().__init__(likelihood_lower_bound)
        self.channels = int(channels)
        self.init_scale = float(init_scale)
        self.filters = tuple(int(f) for f in hidden_dims)
        self.dtype = torch.float32
        dims = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1.0 / (len(self.filters) + 1))
        self.matrices, self.biases, self.factors = nn.ParameterList(), nn.ParameterList(), nn.ParameterList()
        for i in range(len(dims) - 1):
            fan_in, fan_out = dims[i], dims[i + 1]
            self.matrices.append(nn.Parameter(torch.full((self.channels, fan_out, fan_in),
                                                         math.log(math.expm1(1.0 / scale / fan_out)), dtype=self.dtype)))
            self.biases.append(nn.Parameter(torch.empty((self.channels, fan_out, 1), dtype=self.dtype).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.factors.append(nn.Parameter(torch.zeros((self.channels, fan_out, 1), dtype=self.dtype)))
        self._packed = None

    def packed(self) -> Tensor:
        """[C, 43] table of softplus(matrices) / biases / tanh(factors) (nic_pack_factorized)."""
        if self.filters != (3, 3, 3):
            raise NotImplementedError("the factorized kernel is built for hidden_dims=(3, 3, 3) (the reference default)")
        ts = list(self.matrices) + list(self.biases) + list(self.factors)
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if self._packed is None or self._packed[0] != key:
            lib = _lib.load()
            engine.require_cuda(ts[0], "factorized parameters")
            dev = ts[0].device
            out = torch.empty((self.channels, 43), dtype=torch.float32, device=dev)
            c = [t.detach().float().contiguous() for t in ts]
            m, b, f = c[0:4], c[4:8], c[8:11]
            with torch.cuda.device(dev):
                check(lib.nic_pack_factorized(self.channels, ptr(m[0]), ptr(b[0]), ptr(f[0]), ptr(m[1]), ptr(b[1]), ptr(f[1]),
                                              ptr(m[2]), ptr(b[2]), ptr(f[2]), ptr(m[3]), ptr(b[3]), ptr(out),
                                              current_stream()), "nic_pack_factorized")
            self._packed = (key, out)
        return self._packed[1]

    def likelihood(self, inputs: Tensor, qmode: int = Q_PASSTHRU, noise: Tensor = None, want_in: bool = False):
        """(z_in | None, p, logp, partials) from nic_factorized_likelihood_fwd."""
        if inputs.dim() < 2:
            raise ValueError("inputs must be at least 2D with channel axis")
        self._check_bound()
        lib = _lib.load()
        engine.require_cuda(inputs, "z")
        z = inputs.contiguous().float()
        b, c = z.shape[:2]
        hw = z[0, 0].numel()
        fp = self.packed()
        z_in = torch.empty_like(z) if want_in else None
        p, logp = torch.empty_like(z), torch.empty_like(z)
        parts = engine.partials(b, z.device)
        if noise is not None:
            noise = noise.contiguous().float()
        with torch.cuda.device(z.device):
            check(lib.nic_factorized_likelihood_fwd(ptr(z), ptr(fp), ptr(noise), b, c, hw, qmode, ptr(z_in), ptr(p), ptr(logp),
                                                    ptr(parts), current_stream()), "nic_factorized_likelihood_fwd")
        return z_in, p, logp, parts

    def _likelihood(self, inputs: Tensor) -> Tensor:
        return self.likelihood(inputs)[1]

    # ---- per-channel diagnostics used by Trainer._log_entropy_cdf (Trainer.py:255-345): cold path, plain torch ----
    @torch.no_grad()
    def channel_logits_cumulative(self, ch: int, x: torch.Tensor) -> torch.Tensor:
        h = x.reshape(1, -1)
        for i, (mat, bias) in enumerate(zip(self.matrices, self.biases)):
            h = F.softplus(mat[ch]) @ h + bias[ch]
            if i < len(self.factors):
                h = h + torch.tanh(self.factors[i][ch]) * torch.tanh(h)
        return h.reshape(-1)

    @torch.no_grad()
    def channel_cdf(self, ch: int, x: torch.Tensor) -> torch.Tensor:
        return torch.sigmoid(self.channel_logits_cumulative(ch, x))

    @torch.no_grad()
    def channel_pmf(self, ch: int, x: torch.Tensor) -> torch.Tensor:
        up = torch.sigmoid(self.channel_logits_cumulative(ch, x + 0.5))
        lo = torch.sigmoid(self.channel_logits_cumulative(ch, x - 0.5))
        return (up - lo).clamp_min(1e-12)


def _gm_pmf(x: Tensor, weights, mus: Tensor, sigmas: Tensor, K: int, clamp: bool = True) -> Tensor:
    lib = _lib.load()
    engine.require_cuda(x, "x")
    x = x.contiguous().float()
    b, m = x.shape[:2]
    hw = x[0, 0].numel()
    want = (b, m) + tuple(x.shape[2:]) if K == 1 else (b, K, m) + tuple(x.shape[2:])
    def prep(t):
        return None if t is None else t.float().expand(want).contiguous()
    weights, mus, sigmas = prep(weights), prep(mus), prep(sigmas)
    p = torch.empty_like(x)
    with torch.cuda.device(x.device):
        fn = lib.nic_gm_pmf_fwd if clamp else lib.nic_gm_pmf_mass_fwd
        check(fn(ptr(x), ptr(weights), ptr(mus), ptr(sigmas), b, m, hw, K, ptr(p), current_stream()), "nic_gm_pmf_fwd")
    return p


class GaussianConditional(EntropyModel):
    """Mean-scale Gaussian bin mass (EntropyModels.py:188-207)."""

    def discretized_gaussian_pmf(self, x: Tensor, mu: Tensor, sigma: Tensor) -> Tensor:
        """Phi((x + .5 - mu) / sigma) - Phi((x - .5 - mu) / sigma), unclamped (EntropyModels.py:192-204).  x [B, M, ...] with mu /
        sigma broadcastable to it; the reference's mixture code also calls it with x [B, 1, M, ...] against [B, K, M, ...]
        parameters - that form returns the per-component masses [B, K, M, ...]."""
        if x.dim() >= 3 and mu.dim() == x.dim() and x.shape[1] == 1 and mu.shape[1] > 1 and x.shape[2:] == mu.shape[2:]:
            K = int(mu.shape[1])
            xe = x.expand(mu.shape).reshape((x.shape[0], K * mu.shape[2]) + tuple(mu.shape[3:]))
            flat = lambda t: t.expand(mu.shape).reshape(xe.shape)                 # noqa: E731
            return _gm_pmf(xe, None, flat(mu), flat(sigma), 1, clamp=False).reshape(mu.shape)
        return _gm_pmf(x, None, mu, sigma, 1, clamp=False)

    def _likelihood(self, x: Tensor, mu: Tensor, sigma: Tensor) -> Tensor:
        return _gm_pmf(x, None, mu, sigma, 1)


class GaussianMixtureConditional(GaussianConditional):
    """K-component mixture of Gaussian bin masses (EntropyModels.py:210-233)."""

    def discretized_mixture_pmf(self, x: Tensor, weights: Tensor, mus: Tensor, sigmas: Tensor) -> Tensor:
        """sum_k w_k * pmf_k, unclamped (EntropyModels.py:214-230): x [B, M, H, W]; weights / mus / sigmas [B, K, M, H, W]."""
        return _gm_pmf(x, weights, mus, sigmas, int(mus.shape[1]), clamp=False)

    def _likelihood(self, x: Tensor, weights: Tensor, mus: Tensor, sigmas: Tensor) -> Tensor:
        return _gm_pmf(x, weights, mus, sigmas, int(mus.shape[1]))
