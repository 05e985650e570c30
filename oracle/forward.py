"""CPU restatement of the reference forward pass + rate-distortion terms.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional: every routine takes
a reference-layout ``state_dict`` and plain tensors, and
evaluates the same arithmetic the reference evaluates, in the same operation
order, with ``torch`` CPU ops (the reference itself is torch; ATen CPU fp32 is
its definition of the arithmetic).  ``dtype=torch.float64`` gives the shadow
run used to separate "our error" from the reference's own fp32 rounding.

Parity pin: ``oracle/make_golden.py`` imports the real reference classes from
``/root/reference`` (with a matplotlib stub and the GDN restatement of
``oracle/gdn.py`` standing in for the absent ``compressai``), runs them on
seeded weights/inputs and commits the outputs under ``tests/golden``;
``tests/test_oracle.py`` checks this file against those vectors.  The GDN
arithmetic itself is *unpinned by the reference* (third-party, un-vendored,
un-versioned ``compressai``); see oracle/gdn.py.

Reference lines followed (all under /root/reference):
  Models.py:49-106            forward orchestration, output dict
  Components.py:6-18          g_a   (4x conv 5x5 s2 p2, GDN after first three)
  Components.py:35-47         g_s   (4x convT 5x5 s2 p2 op1, IGDN after first three)
  Components.py:65-75         h_a   (3x3 s1, 5x5 s2, 5x5 s2, LeakyReLU(0.01))
  Components.py:94-105        h_s   (convT, convT, 3x3 s1, LeakyReLU(0.01))
  ContextModels.py:9-20       mask 'A' 5x5, weight *= mask
  ParametersModels.py:29-64   1x1 stack, chunk/view, softmax, softplus + 1e-6
  EntropyModels.py:29-31      clamp_min(1e-9)
  EntropyModels.py:88-151     factorized prior
  EntropyModels.py:192-230    discretized Gaussian / mixture pmf
  utils.py:6-8                Phi(u) = 0.5 * (1 + erf(u / sqrt(2)))
  RateDistortionLoss.py:5-49  rd_loss
  Layers.py:18-119, Components.py:20-122, Models.py:150-205   the 3x3 residual family (forward_residual)
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .gdn import gdn_effective

LIKELIHOOD_BOUND = 1e-9      # EntropyModels.py:18
SIGMA_FLOOR = 1e-6           # ParametersModels.py:47,62


DIFFERENTIABLE = False      # oracle/backward.py sets this while it builds an autograd graph over the parameters
DEVICE = "cpu"              # tools/library_baseline.py sets "cuda" to time the same torch ops on the GPU (cuDNN / ATen eager)


def _p(sd, key, dtype):
    t = sd[key] if DIFFERENTIABLE else sd[key].detach()
    return t.to(DEVICE, dtype)


def _gdn(sd, prefix, x, inverse, dtype):
    """compressai GDN forward, see oracle/gdn.py (call sites Components.py:11-15, 40-44)."""
    beta, gamma = gdn_effective(_p(sd, prefix + ".beta", dtype), _p(sd, prefix + ".gamma", dtype))
    C = beta.numel()
    norm = F.conv2d(x * x, gamma.reshape(C, C, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


def analysis(sd, x, dtype=torch.float32):
    """Encoder5x5, Components.py:9-18."""
    h = x
    for i in (0, 2, 4, 6):
        h = F.conv2d(h, _p(sd, f"encoder.net.{i}.weight", dtype), _p(sd, f"encoder.net.{i}.bias", dtype),
                     stride=2, padding=2)
        if i != 6:
            h = _gdn(sd, f"encoder.net.{i + 1}", h, False, dtype)
    return h


def synthesis(sd, y_in, dtype=torch.float32):
    """Decoder5x5, Components.py:38-47."""
    h = y_in
    for i in (0, 2, 4, 6):
        h = F.conv_transpose2d(h, _p(sd, f"decoder.net.{i}.weight", dtype), _p(sd, f"decoder.net.{i}.bias", dtype),
                               stride=2, padding=2, output_padding=1)
        if i != 6:
            h = _gdn(sd, f"decoder.net.{i + 1}", h, True, dtype)
    return h


def hyper_analysis(sd, y, dtype=torch.float32):
    """HyperEncoder5x5, Components.py:68-75 (input is the UNquantised y, Models.py:53)."""
    h = F.conv2d(y, _p(sd, "hyper_encoder.net.0.weight", dtype), _p(sd, "hyper_encoder.net.0.bias", dtype), padding=1)
    h = F.leaky_relu(h, 0.01)
    h = F.conv2d(h, _p(sd, "hyper_encoder.net.2.weight", dtype), _p(sd, "hyper_encoder.net.2.bias", dtype),
                 stride=2, padding=2)
    h = F.leaky_relu(h, 0.01)
    return F.conv2d(h, _p(sd, "hyper_encoder.net.4.weight", dtype), _p(sd, "hyper_encoder.net.4.bias", dtype),
                    stride=2, padding=2)


def hyper_synthesis(sd, z_in, dtype=torch.float32):
    """HyperDecoder5x5, Components.py:98-105."""
    h = F.conv_transpose2d(z_in, _p(sd, "hyper_decoder.net.0.weight", dtype), _p(sd, "hyper_decoder.net.0.bias", dtype),
                           stride=2, padding=2, output_padding=1)
    h = F.leaky_relu(h, 0.01)
    h = F.conv_transpose2d(h, _p(sd, "hyper_decoder.net.2.weight", dtype), _p(sd, "hyper_decoder.net.2.bias", dtype),
                           stride=2, padding=2, output_padding=1)
    h = F.leaky_relu(h, 0.01)
    return F.conv2d(h, _p(sd, "hyper_decoder.net.4.weight", dtype), _p(sd, "hyper_decoder.net.4.bias", dtype), padding=1)


def mask_a(weight):
    """Mask 'A' of ContextModels.py:13-16: rows above centre, and centre row left of centre."""
    kh, kw = weight.shape[-2:]
    m = torch.ones_like(weight)
    m[:, :, kh // 2, kw // 2:] = 0
    m[:, :, kh // 2 + 1:] = 0
    return m


def context(sd, y_in, dtype=torch.float32, prefix="context_model"):
    """ContextModel / MaskedConv2d('A'), ContextModels.py:18-20, 26-33."""
    w = _p(sd, f"{prefix}.masked.weight", dtype)
    if DIFFERENTIABLE:
        # the reference zeroes the masked taps of the PARAMETER in place (weight.data *= mask, ContextModels.py:19) and then
        # convolves with the parameter itself: the forward sees w * mask, the gradient reaches all 25 taps
        w = (w.detach() * mask_a(w) - w.detach()) + w
    else:
        w = w * mask_a(w)
    return F.conv2d(y_in, w, _p(sd, f"{prefix}.masked.bias", dtype), padding=2)


def entropy_parameters_raw(sd, combined, dtype=torch.float32, prefix="entropy_parameters"):
    """EntropyParameters.net, ParametersModels.py:21-35: three 1x1 convs, LeakyReLU(0.01) between."""
    h = combined
    for i in (0, 2, 4):
        h = F.conv2d(h, _p(sd, f"{prefix}.net.{i}.weight", dtype),
                     _p(sd, f"{prefix}.net.{i}.bias", dtype))
        if i != 4:
            h = F.leaky_relu(h, 0.01)
    return h


def split_parameters(raw, M, K):
    """ParametersModels.py:43-64.  K == 1 -> (mu, sigma); K > 1 -> (weights, mus, sigmas) [B,K,M,H,W]."""
    if K == 1:
        mu, sigma = raw.chunk(2, dim=1)
        return mu, F.softplus(sigma) + SIGMA_FLOOR
    w, mu, s = torch.chunk(raw, 3, dim=1)
    B, _, H, W = raw.shape
    w = F.softmax(w.reshape(B, K, M, H, W), dim=1)
    mu = mu.reshape(B, K, M, H, W)
    s = F.softplus(s.reshape(B, K, M, H, W)) + SIGMA_FLOOR
    return w, mu, s


def gaussian_cdf(u):
    """utils.py:6-8."""
    return 0.5 * (1.0 + torch.erf(u / math.sqrt(2.0)))


def gaussian_pmf(x, mu, sigma):
    """EntropyModels.py:199-204 (no abs, no tail-stable form)."""
    upper = (x + 0.5 - mu) / sigma
    lower = (x - 0.5 - mu) / sigma
    return gaussian_cdf(upper) - gaussian_cdf(lower)


def conditional_likelihood(y_in, params, K):
    """GaussianConditional / GaussianMixtureConditional forward incl. the 1e-9 clamp (EntropyModels.py:29-31, 206, 223-233)."""
    if K == 1:
        mu, sigma = params
        p = gaussian_pmf(y_in, mu, sigma)
    else:
        w, mus, sigmas = params
        p = torch.sum(w * gaussian_pmf(y_in.unsqueeze(1), mus, sigmas), dim=1)
    return p.clamp_min(LIKELIHOOD_BOUND)


def factorized_logits(sd, v, dtype=torch.float32, prefix="factorized_entropy_model"):
    """EntropyModels.py:88-111 on v of shape (C, 1, N)."""
    logits = v
    for i in range(4):
        m = F.softplus(_p(sd, f"{prefix}.matrices.{i}", dtype))
        logits = torch.matmul(m, logits) + _p(sd, f"{prefix}.biases.{i}", dtype)
        if i < 3:
            logits = logits + torch.tanh(_p(sd, f"{prefix}.factors.{i}", dtype)) * torch.tanh(logits)
    return logits


def factorized_likelihood(sd, z_in, dtype=torch.float32, prefix="factorized_entropy_model"):
    """FactorizedEntropyBottleneck forward incl. clamp (EntropyModels.py:113-151, 29-31)."""
    B, C = z_in.shape[:2]
    flat = z_in.transpose(0, 1).reshape(C, 1, -1)
    lower = factorized_logits(sd, flat - 0.5, dtype, prefix)
    upper = factorized_logits(sd, flat + 0.5, dtype, prefix)
    s = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(s * upper) - torch.sigmoid(s * lower))
    pmf = pmf.reshape(C, B, *z_in.shape[2:]).transpose(0, 1)
    return pmf.clamp_min(LIKELIHOOD_BOUND)


def forward(sd, x, M: int, K: int, training: bool = False,
            noise_z: Optional[torch.Tensor] = None, noise_y: Optional[torch.Tensor] = None,
            dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """JointAutoregressiveHierarchical.forward, Models.py:49-106.

    ``training=True`` needs the two U(-0.5, 0.5) tensors injected (the reference draws the
    z noise before the y noise, Models.py:57-58).
    """
    x = x.to(DEVICE, dtype)
    y = analysis(sd, x, dtype)
    z = hyper_analysis(sd, y, dtype)
    if training:
        z_in = z + noise_z.to(DEVICE, dtype)
        y_in = y + noise_y.to(DEVICE, dtype)
    else:
        z_in = torch.round(z)
        y_in = torch.round(y)
    psi = hyper_synthesis(sd, z_in, dtype)
    phi = context(sd, y_in, dtype)
    raw = entropy_parameters_raw(sd, torch.cat([phi, psi], dim=1), dtype)
    params = split_parameters(raw, M, K)
    p_z = factorized_likelihood(sd, z_in, dtype)
    p_y = conditional_likelihood(y_in, params, K)
    out = {
        "x_hat": synthesis(sd, y_in, dtype), "y": y, "y_in": y_in, "z": z, "z_in": z_in,
        "p_z": p_z, "logp_z": torch.log(p_z), "p_y": p_y, "logp_y": torch.log(p_y),
        "training": training,
        # not part of the reference dict; kept for kernel-level checks
        "_psi": psi, "_phi": phi, "_raw": raw,
    }
    if K == 1:
        out["mu"], out["sigma"] = params
    else:
        out["weights"], out["mus"], out["sigmas"] = params
    return out


def rd_loss(out, x, lambda_rd: float):
    """RateDistortionLoss.py:5-49 (same key set; scalars as Python floats)."""
    x = x.to(out["x_hat"].dtype)
    bits_y = -torch.sum(out["logp_y"], dim=(1, 2, 3)) / math.log(2.0)
    bits_z = -torch.sum(out["logp_z"], dim=(1, 2, 3)) / math.log(2.0)
    num_pixels = x.size(2) * x.size(3)
    bpp_y = (bits_y / num_pixels).mean()
    bpp_z = (bits_z / num_pixels).mean()
    bpp_total = bpp_y + bpp_z
    mse_per_image = torch.mean((out["x_hat"] - x) ** 2, dim=(1, 2, 3))
    mse = mse_per_image.mean()
    psnr = -10 * torch.log10(mse + 1e-8)
    psnr_per_image = -10 * torch.log10(mse_per_image + 1e-8)
    loss = bpp_total + lambda_rd * (255 ** 2) * mse
    return {
        "loss": loss, "bpp_y": bpp_y.item(), "bpp_z": bpp_z.item(), "bpp_total": bpp_total.item(),
        "mse": mse.item(), "psnr": psnr.item(), "mse_per_image": mse_per_image.detach(),
        "psnr_per_image": psnr_per_image.detach(), "bits_y": bits_y.mean().item(),
        "bits_z": bits_z.mean().item(), "bits_total": (bits_y + bits_z).mean().item(),
    }


# ---- the 3x3 residual family (Layers.py, Components.py:20-122, Models.py:109-205) -------------------------------------------------

def _c(sd, key, x, dtype, stride=1, padding=0):
    return F.conv2d(x, _p(sd, key + ".weight", dtype), _p(sd, key + ".bias", dtype), stride=stride, padding=padding)


def _d(sd, key, x, dtype):
    """TransposedDeconv3x3, Layers.py:18-24: ConvTranspose2d(3, stride 2, padding 1, output_padding 1)."""
    return F.conv_transpose2d(x, _p(sd, key + ".deconv.weight", dtype), _p(sd, key + ".deconv.bias", dtype), stride=2, padding=1,
                              output_padding=1)


def _res_block(sd, k, x, dtype):
    """ResidualBlock, Layers.py:91-119 (in_ch == out_ch on this path: identity skip)."""
    out = F.leaky_relu(_c(sd, k + ".conv1", x, dtype, padding=1), 0.01)
    out = F.leaky_relu(_c(sd, k + ".conv2", out, dtype, padding=1), 0.01)
    idn = _c(sd, k + ".skip", x, dtype) if (k + ".skip.weight") in sd else x
    return out + idn


def _res_stride(sd, k, x, dtype):
    """ResidualBlockWithStride, Layers.py:27-61: conv(s2) - LeakyReLU - conv - GDN, + 1x1 strided skip."""
    out = F.leaky_relu(_c(sd, k + ".conv1", x, dtype, stride=2, padding=1), 0.01)
    out = _gdn(sd, k + ".gdn", _c(sd, k + ".conv2", out, dtype, padding=1), False, dtype)
    return out + _c(sd, k + ".skip", x, dtype, stride=2)


def _res_up(sd, k, x, dtype):
    """ResidualBlockUpsample, Layers.py:64-88: deconv - LeakyReLU - conv - IGDN, + deconv skip."""
    out = F.leaky_relu(_d(sd, k + ".subpel_conv", x, dtype), 0.01)
    out = _gdn(sd, k + ".igdn", _c(sd, k + ".conv", out, dtype, padding=1), True, dtype)
    return out + _d(sd, k + ".upsample", x, dtype)


def analysis_3x3(sd, x, dtype=torch.float32):
    """Encoder3x3, Components.py:20-32."""
    h = x
    for i in (0, 2, 4):
        h = _res_block(sd, f"encoder.net.{i + 1}", _res_stride(sd, f"encoder.net.{i}", h, dtype), dtype)
    return _c(sd, "encoder.net.6", h, dtype, stride=2, padding=1)


def synthesis_3x3(sd, y_in, dtype=torch.float32):
    """Decoder3x3, Components.py:49-62."""
    h = y_in
    for i in (0, 2, 4):
        h = _res_up(sd, f"decoder.net.{i + 1}", _res_block(sd, f"decoder.net.{i}", h, dtype), dtype)
    return _d(sd, "decoder.net.7", _res_block(sd, "decoder.net.6", h, dtype), dtype)


def hyper_analysis_3x3(sd, y, dtype=torch.float32):
    """HyperEncoder3x3, Components.py:77-91."""
    h = y
    for i, s in zip((0, 2, 4, 6, 8), (1, 1, 2, 1, 2)):
        h = _c(sd, f"hyper_encoder.net.{i}", h, dtype, stride=s, padding=1)
        if i != 8:
            h = F.leaky_relu(h, 0.01)
    return h


def hyper_synthesis_3x3(sd, z_in, dtype=torch.float32):
    """HyperDecoder3x3, Components.py:107-122."""
    h = F.leaky_relu(_c(sd, "hyper_decoder.net.0", z_in, dtype, padding=1), 0.01)
    h = F.leaky_relu(_d(sd, "hyper_decoder.net.2", h, dtype), 0.01)
    h = F.leaky_relu(_c(sd, "hyper_decoder.net.4", h, dtype, padding=1), 0.01)
    h = F.leaky_relu(_d(sd, "hyper_decoder.net.6", h, dtype), 0.01)
    return _c(sd, "hyper_decoder.net.8", h, dtype, padding=1)


def forward_residual(sd, x, M: int, K: int, training: bool = False, noise_z=None, noise_y=None, dtype=torch.float32):
    """HierarchicalMixtureResidual.forward, Models.py:150-205 (same orchestration as `forward` with the 3x3 transforms)."""
    x = x.to(DEVICE, dtype)
    y = analysis_3x3(sd, x, dtype)
    z = hyper_analysis_3x3(sd, y, dtype)
    if training:
        z_in, y_in = z + noise_z.to(DEVICE, dtype), y + noise_y.to(DEVICE, dtype)
    else:
        z_in, y_in = torch.round(z), torch.round(y)
    psi = hyper_synthesis_3x3(sd, z_in, dtype)
    phi = context(sd, y_in, dtype)
    raw = entropy_parameters_raw(sd, torch.cat([phi, psi], dim=1), dtype)
    params = split_parameters(raw, M, K)
    p_z = factorized_likelihood(sd, z_in, dtype)
    p_y = conditional_likelihood(y_in, params, K)
    out = {"x_hat": synthesis_3x3(sd, y_in, dtype), "y": y, "y_in": y_in, "z": z, "z_in": z_in, "p_z": p_z, "logp_z": torch.log(p_z),
           "p_y": p_y, "logp_y": torch.log(p_y), "training": training}
    if K == 1:
        out["mu"], out["sigma"] = params
    else:
        out["weights"], out["mus"], out["sigmas"] = params
    return out


def forward_scalable(sd, x, M: int, M1: int, K: int, training: bool = False,
                     noise_z: Optional[torch.Tensor] = None, noise_y: Optional[torch.Tensor] = None,
                     dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """ScalableImageCoding.forward, Models.py:259-338, with the four repairs listed in SURVEY.md section 2.4 (the committed
    forward raises): factorized(z_in) without the stray argument, conditional(y_i, mu=, sigma=) keyword names, the second
    parameter dict bound to params2, and no LatentSpaceTransform / 'F_tilde'."""
    x = x.to("cpu", dtype)
    y = analysis(sd, x, dtype)
    z = hyper_analysis(sd, y, dtype)
    if training:
        z_in, y_in = z + noise_z.to(dtype), y + noise_y.to(dtype)
    else:
        z_in, y_in = torch.round(z), torch.round(y)
    y1, y2 = torch.split(y_in, [M1, M - M1], dim=1)                                   # Models.py:279
    psi = hyper_synthesis(sd, z_in, dtype)
    out = {"y": y, "y_in": y_in, "y1": y1, "y2": y2, "z": z, "z_in": z_in, "training": training}
    for i, (yi, mi) in enumerate(((y1, M1), (y2, M - M1)), start=1):
        phi = context(sd, yi, dtype, prefix=f"context_model_{i}")                    # :284-285
        raw = entropy_parameters_raw(sd, torch.cat([phi, psi], dim=1), dtype, prefix=f"entropy_parameters_{i}")   # :287-299
        params = split_parameters(raw, mi, K)
        p = conditional_likelihood(yi, params, K)                                     # :305-306
        out[f"p_y{i}"], out[f"logp_y{i}"] = p, torch.log(p)
        if K == 1:
            out[f"mu{i}"], out[f"sigma{i}"] = params
        else:
            out[f"weights{i}"], out[f"mus{i}"], out[f"sigmas{i}"] = params
    p_z = factorized_likelihood(sd, z_in, dtype)                                      # :302
    out["p_z"], out["logp_z"] = p_z, torch.log(p_z)
    out["x_hat"] = synthesis(sd, y_in, dtype)                                         # :316
    return out


def vision_rd_loss(out, x, lambda_rd: float):
    """RateDistortionLoss.py:52-121 with frozen_activation = V = None (no feature-space term)."""
    x = x.to(out["x_hat"].dtype)
    ln2 = math.log(2.0)
    b1 = -torch.sum(out["logp_y1"], dim=(1, 2, 3)) / ln2
    b2 = -torch.sum(out["logp_y2"], dim=(1, 2, 3)) / ln2
    bz = -torch.sum(out["logp_z"], dim=(1, 2, 3)) / ln2
    n = x.size(2) * x.size(3)
    bpp_y1, bpp_y2, bpp_z = (b1 / n).mean(), (b2 / n).mean(), (bz / n).mean()
    bpp_total = bpp_y1 + bpp_y2 + bpp_z
    mse_i = torch.mean((out["x_hat"] - x) ** 2, dim=(1, 2, 3))
    mse = mse_i.mean()
    return {"loss": bpp_total + lambda_rd * mse, "bpp_y1": bpp_y1.item(), "bpp_y2": bpp_y2.item(), "bpp_y": (bpp_y1 + bpp_y2).item(),
            "bpp_z": bpp_z.item(), "bpp_total": bpp_total.item(), "mse": mse.item(), "reconstruction_mse": mse.item(),
            "psnr": (-10 * torch.log10(mse + 1e-8)).item(), "vision_mse": 0.0, "mse_per_image": mse_i.detach(),
            "psnr_per_image": (-10 * torch.log10(mse_i + 1e-8)).detach(), "bits_y1": b1.mean().item(), "bits_y2": b2.mean().item(),
            "bits_y": (b1 + b2).mean().item(), "bits_z": bz.mean().item(), "bits_total": (b1 + b2 + bz).mean().item()}
