"""``JointAutoregressiveHierarchical`` of /root/reference/Models.py:10-106 on the sm_100a kernels.

Same constructor, attributes, ``state_dict`` keys and output dict.  ``forward`` issues one fused
chain of C-ABI calls on the current CUDA stream (activations NHWC between layers, nothing returns to
the host):

    g_a   4 x nic_conv_fwd (GDN fused)                              Models.py:52  Components.py:9-17
    y     nic_latent_handoff  -> 'y', 'y_in' (round | + noise)      Models.py:57-64
    h_a   3 x nic_conv_fwd (LeakyReLU fused)                        Models.py:53  Components.py:68-74
    z     nic_latent_handoff  -> 'z', 'z_in'
    h_s   3 x nic_conv_fwd, last one writes channels [2M, 4M)       Models.py:69  Components.py:98-104
    ctx   nic_conv_fwd (12 live taps), writes channels [0, 2M)      Models.py:71  ContextModels.py:18-20
          (so torch.cat([phi, psi]) of Models.py:73 never happens)
    ep    3 x nic_conv_fwd (1x1), last one writes NCHW raw params   Models.py:76-80 ParametersModels.py:29-35
    p_y   nic_gm_likelihood_fwd (+ weights/mus/sigmas, log, sums)   Models.py:86-87
    p_z   nic_factorized_likelihood_fwd (+ log, sums)               Models.py:83-84
    g_s   4 x nic_conv_fwd (IGDN fused), last one writes NCHW x_hat Models.py:90  Components.py:38-46
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import engine
from ._lib import LAYOUT_NCHW, LAYOUT_NHWC, MODEL_PRECISIONS, Q_NOISE, Q_PASSTHRU, Q_ROUND
from .Components import Decoder5x5, Encoder5x5, HyperDecoder5x5, HyperEncoder5x5
from .ContextModels import ContextModel
from .EntropyModels import FactorizedEntropyBottleneck, GaussianConditional, GaussianMixtureConditional, gm_likelihood
from .ParametersModels import EntropyParameters


class JointAutoregressiveHierarchical(nn.Module):
    """
    latent_channels : int, default=192, number of channels in the bottleneck y (M).
    K : int, default=1.  K == 1 -> mean-scale Gaussian; K > 1 -> mixture of K Gaussians.
    precision : arithmetic of the transforms (keyword-only extension; the reference has no such switch):
        None / "auto" (default, or $NIC_PRECISION): "bf16x3" when latent_channels == 128, else "fp32" - both parity grade;
        "fp32"  CUDA-core FFMA kernels, fp32 operands and accumulation - parity grade, any channel count;
        "bf16"  tcgen05 tensor-core kernels, bf16 operands, fp32 accumulation - the throughput arm;
        "mixed" g_a and h_a (everything upstream of the rounding, i.e. what decides the symbols) in fp32, the entropy
                path and g_s in bf16: symbols identical to the fp32 arm at ~2.5x its speed;
        "bf16x3" every transform on the tensor cores at fp32 grade: both operands of every contraction (convs and the GDN
                channel mixing) split into bf16 hi + lo and contracted as hi.hi + lo.hi + hi.lo in one fp32 TMEM accumulation
                (~1e-5 relative; 3x the MMA work), activations carried between layers as bf16 hi/lo pairs - the
                parity-grade tensor-core arm.
    """

    def __init__(self, latent_channels: int = 192, K: int = 1, *, precision: Optional[str] = None):
        super().__init__()
        if not isinstance(latent_channels, int) or latent_channels < 1:
            raise ValueError(f"latent_channels must be int >= 1, got {latent_channels}")
        if not isinstance(K, int) or K < 1:
            raise ValueError(f"K must be int >= 1, got {K}")
        self.M = latent_channels
        self.K = K
        self.H = latent_channels
        self.distribution = "Mean-Scale Gaussian" if K == 1 else "Mixture of Gaussians"
        self.conditional = GaussianConditional() if K == 1 else GaussianMixtureConditional()
        # construction order follows Models.py:34-46 so a seeded build draws the same initial weights
        self.encoder = Encoder5x5(latent_channels=self.M)
        self.decoder = Decoder5x5(latent_channels=self.M)
        self.hyper_encoder = HyperEncoder5x5(latent_channels=self.M)
        self.hyper_decoder = HyperDecoder5x5(latent_channels=self.M)
        self.factorized_entropy_model = FactorizedEntropyBottleneck(self.M)
        self.context_model = ContextModel(latent_channels=self.M)
        self.entropy_parameters = EntropyParameters(latent_channels=self.M, hyper_latent_channels=self.H, K=self.K)
        self.precision = engine.resolve_precision(precision, latent_channels)
        if self.precision not in MODEL_PRECISIONS:
            raise ValueError(f"precision must be one of {MODEL_PRECISIONS}, got {self.precision}")

    def forward(self, x: torch.Tensor, training: bool = True, *,
                noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, lean: bool = False):
        """Returns the reference's dict (Models.py:92-104).

        noise : optional (noise_z, noise_y) in U(-0.5, 0.5) to use instead of drawing them
                (training=True only; the reference draws z's noise first, Models.py:57-58).
        lean  : skip materialising weights/mus/sigmas (mu/sigma); those keys are then absent.
        """
        engine.require_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected x of shape [B, 3, H, W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 64 or W % 64:
            raise ValueError(f"H and W must be multiples of 64 (four stride-2 stages in g_a, two in h_a); got {H}x{W}")
        prec_up = {"fp32": "fp32", "mixed": "fp32", "bf16x3": "bf16x3", "bf16": "bf16"}[self.precision]   # g_a, h_a
        prec = {"fp32": "fp32", "mixed": "bf16", "bf16x3": "bf16x3", "bf16": "bf16"}[self.precision]    # h_s, context, entropy parameters, g_s
        adt = engine.act_dtype(prec)
        pair = prec == "bf16x3"                      # activations are bf16 hi/lo pairs: [hi(c) | lo(c)] channels
        cw = 2 if pair else 1
        M, K = self.M, self.K
        x = x.contiguous().float()
        hy, wy, hz, wz = H // 16, W // 16, H // 64, W // 64
        with torch.cuda.device(x.device), torch.no_grad():
            noise_z = noise_y = None
            if training:
                if noise is not None:
                    noise_z, noise_y = noise
                else:
                    noise_z = torch.rand((B, M, hz, wz), device=x.device) - 0.5
                    noise_y = torch.rand((B, M, hy, wy), device=x.device) - 0.5
            qmode = Q_NOISE if training else Q_ROUND

            # ---- g_a ------------------------------------------------------------------------------
            a, h, w, layout = x, H, W, LAYOUT_NCHW
            enc = self.encoder.ops
            for i, op in enumerate(enc):
                a = op.run(a, B, h, w, prec_up, in_layout=layout, out_layout=LAYOUT_NHWC,
                           out_dtype=torch.float32 if i == len(enc) - 1 else None)
                h, w = engine.conv_out_hw(op.conv, h, w)
                layout = LAYOUT_NHWC
            y_nhwc = a                                                     # f32 [B, hy, wy, M]
            lowp = prec_up != "fp32"
            y, y_in, y_in_nhwc, y_lowp = engine.latent_handoff(y_nhwc, qmode, noise_y, "bf16x2" if pair else adt, want_lowp=lowp,
                                                               lowp_pair=prec_up == "bf16x3")

            # ---- h_a (reads the unquantised y, Models.py:53) ------------------------------------
            a, h, w = (y_lowp if lowp else y_nhwc), hy, wy
            ha = self.hyper_encoder.ops
            for i, op in enumerate(ha):
                a = op.run(a, B, h, w, prec_up, out_dtype=torch.float32 if i == len(ha) - 1 else None)
                h, w = engine.conv_out_hw(op.conv, h, w)
            z, z_in, z_in_nhwc, _ = engine.latent_handoff(a, qmode, noise_z, "bf16x2" if pair else adt)

            # ---- h_s -> psi = combined[..., 2M:4M];  context -> phi = combined[..., 0:2M] -------
            combined = torch.empty((B, hy, wy, cw * 4 * M), dtype=adt, device=x.device)
            a, h, w = z_in_nhwc, hz, wz
            hs = self.hyper_decoder.ops
            for i, op in enumerate(hs):
                if i == len(hs) - 1:
                    op.run(a, B, h, w, prec, out=combined, out_c_total=4 * M, out_c_offset=2 * M)
                else:
                    a = op.run(a, B, h, w, prec)
                h, w = engine.conv_out_hw(op.conv, h, w)
            self.context_model.masked.apply_mask_()
            self.context_model.masked._op.run(y_in_nhwc, B, hy, wy, prec, out=combined, out_c_total=4 * M, out_c_offset=0)

            # ---- entropy parameters (1x1 stack) ---------------------------------------------------
            ep = self.entropy_parameters.ops
            a = ep[0].run(combined, B, hy, wy, prec)
            a = ep[1].run(a, B, hy, wy, prec)
            raw = ep[2].run(a, B, hy, wy, prec, out_layout=LAYOUT_NCHW, out_dtype=torch.float32)

            # ---- likelihoods -------------------------------------------------------------------------
            ly = gm_likelihood(y_in, raw, M, K, Q_PASSTHRU, full=not lean, want_y_in=False)
            _, p_z, logp_z, parts_z = self.factorized_entropy_model.likelihood(z_in, Q_PASSTHRU)
            p_y, logp_y = ly["p"], ly["logp"]

            # ---- g_s -----------------------------------------------------------------------------------
            a, h, w = y_in_nhwc, hy, wy
            dec = self.decoder.ops
            for i, op in enumerate(dec):
                last = i == len(dec) - 1
                a = op.run(a, B, h, w, prec, out_layout=LAYOUT_NCHW if last else LAYOUT_NHWC,
                           out_dtype=torch.float32 if last else None)
                h, w = engine.conv_out_hw(op.conv, h, w)
            x_hat = a

        # per-image partial sums of logp ride along for rd_loss (RateDistortionLoss.py:13-14)
        logp_y._nic_partials = ly["partials"]
        logp_z._nic_partials = parts_z
        out = {
            "x_hat": x_hat, "y": y, "y_in": y_in, "z": z, "z_in": z_in,
            "p_z": p_z, "logp_z": logp_z, "p_y": p_y, "logp_y": logp_y, "training": training,
        }
        if not lean:
            if K == 1:
                out.update({"mu": ly["mu"], "sigma": ly["sigma"]})
            else:
                out.update({"weights": ly["weights"], "mus": ly["mus"], "sigmas": ly["sigmas"]})
        return out
