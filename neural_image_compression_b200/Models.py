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

import contextlib
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import engine
from ._lib import LAYOUT_NCHW, LAYOUT_NHWC, MODEL_PRECISIONS, Q_NOISE, Q_PASSTHRU, Q_ROUND
from .Components import Decoder5x5, Encoder5x5, HyperDecoder5x5, HyperEncoder5x5
from .ContextModels import ContextModel
from .EntropyModels import FactorizedEntropyBottleneck, GaussianConditional, GaussianMixtureConditional, gm_likelihood
from .ParametersModels import EntropyParameters


def _pair_noise(model, x, training, noise):
    B, _, H, W = x.shape
    if not training:
        return None, None
    if noise is not None:
        return noise
    M = model.M
    return (torch.rand((B, M, H // 64, W // 64), device=x.device) - 0.5, torch.rand((B, M, H // 16, W // 16), device=x.device) - 0.5)


def _pair_g_a(model, x, training, noise_y, prec_up, prec):
    """g_a -> y hand-off of the fused pipeline (Models.py:52, 57-64): activations NHWC between layers (bf16 hi/lo pairs in the bf16x3
    arm).  Returns (y, y_in, y_in_nhwc, y_src) with y_src = the unquantised y in the engine's layout for h_a (Models.py:53)."""
    B, _, H, W = x.shape
    adt = engine.act_dtype(prec)
    pair = prec == "bf16x3"
    a, h, w, layout = x, H, W, LAYOUT_NCHW
    enc = model.encoder.ops
    for i, op in enumerate(enc):
        a = op.run(a, B, h, w, prec_up, in_layout=layout, out_layout=LAYOUT_NHWC,
                   out_dtype=torch.float32 if i == len(enc) - 1 else None)
        h, w = engine.conv_out_hw(op.conv, h, w)
        layout = LAYOUT_NHWC
    y_nhwc = a                                                     # f32 [B, hy, wy, M]
    lowp = prec_up != "fp32"
    y, y_in, y_in_nhwc, y_lowp = engine.latent_handoff(y_nhwc, Q_NOISE if training else Q_ROUND, noise_y, "bf16x2" if pair else adt,
                                                       want_lowp=lowp, lowp_pair=prec_up == "bf16x3")
    return y, y_in, y_in_nhwc, (y_lowp if lowp else y_nhwc)


def _pair_h_a(model, y_src, B, hy, wy, training, noise_z, prec_up, prec):
    """h_a -> z hand-off (Models.py:53, 57-64).  Returns (z, z_in, z_in_nhwc)."""
    adt = engine.act_dtype(prec)
    a, h, w = y_src, hy, wy
    ha = model.hyper_encoder.ops
    for i, op in enumerate(ha):
        a = op.run(a, B, h, w, prec_up, out_dtype=torch.float32 if i == len(ha) - 1 else None)
        h, w = engine.conv_out_hw(op.conv, h, w)
    z, z_in, z_in_nhwc, _ = engine.latent_handoff(a, Q_NOISE if training else Q_ROUND, noise_z, "bf16x2" if prec == "bf16x3" else adt)
    return z, z_in, z_in_nhwc


def _pair_h_s_into(model, z_in_nhwc, B, hz, wz, prec, combined, c_total, c_offset):
    """h_s (Models.py:69) with its last conv writing psi into channels [c_offset, c_offset + 2M) of `combined`."""
    z_flag = getattr(z_in_nhwc, "_nic_lo_flag", None)
    a, h, w = z_in_nhwc, hz, wz
    hs = model.hyper_decoder.ops
    for i, op in enumerate(hs):
        if i == len(hs) - 1:
            op.run(a, B, h, w, prec, out=combined, out_c_total=c_total, out_c_offset=c_offset)
        else:
            a = op.run(a, B, h, w, prec, in_lo_flag=z_flag if i == 0 else None)
        h, w = engine.conv_out_hw(op.conv, h, w)


_SIDE_STREAMS = {}


def _branch_stream(device, index: int = 0):
    """A second stream for g_s while the forward pass is being CAPTURED in a CUDA graph: g_s (y_in -> x_hat) and the entropy side
    (h_a, h_s, context, entropy parameters, likelihoods) only share y / y_in, so the graph gets two parallel branches and the small
    launches of the entropy side (32-128 CTAs on 148 SMs) and the tails of the big ones overlap the other branch.  Eager launches
    stay on one stream (they are host-bound; forking costs events).  NIC_EVAL_OVERLAP=0 switches it off."""
    if os.environ.get("NIC_EVAL_OVERLAP", "1") == "0" or not torch.cuda.is_current_stream_capturing():
        return None
    key = (str(device), index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def _pair_g_s(model, y_in_nhwc, B, hy, wy, prec):
    """g_s (Models.py:90): NHWC symbols -> x_hat NCHW f32."""
    y_flag = getattr(y_in_nhwc, "_nic_lo_flag", None)
    a, h, w = y_in_nhwc, hy, wy
    dec = model.decoder.ops
    for i, op in enumerate(dec):
        last = i == len(dec) - 1
        a = op.run(a, B, h, w, prec, out_layout=LAYOUT_NCHW if last else LAYOUT_NHWC,
                   out_dtype=torch.float32 if last else None, in_lo_flag=y_flag if i == 0 else None)
        h, w = engine.conv_out_hw(op.conv, h, w)
    return a


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
        if training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # the reference's training call (Trainer.py:82): autograd is on, so the step must be differentiable.  One autograd
            # node over the fp32 arm with hand-written backward kernels (training.py); under torch.no_grad() the call below
            # runs the forward-only path of self.precision instead.
            from . import training as _training
            return _training.train_forward(self, x, noise=noise, lean=lean)
        if self.precision == "bf16x3" and self.M not in (128, 192):
            # the fused pair-tensor pipeline (GDN kernels, first layer) is built for 128 and 192 channels (the reference's default);
            # other channel counts that are multiples of 64 run layer by layer: every conv and both GDN contractions on the
            # tcgen05 engine, fp32 NHWC tensors in between (training._forward_impl with rounding instead of noise)
            if self.M % 64:
                raise ValueError(f"precision='bf16x3' needs latent_channels % 64 == 0, got {self.M}")
            from . import training as _training
            from ._lib import Q_NOISE as _QN, Q_ROUND as _QR
            with torch.no_grad():
                nz = ny = None
                if training:
                    nz, ny = noise if noise is not None else (torch.rand((B, self.M, H // 64, W // 64), device=x.device) - 0.5,
                                                              torch.rand((B, self.M, H // 16, W // 16), device=x.device) - 0.5)
                _training.forget_pairs()
                res, _ = _training._forward_impl(self, x.contiguous().float(), nz, ny, lean, qmode=_QN if training else _QR, arm="bf16x3")
                _training.forget_pairs()
            x_hat, logp_y, logp_z, y, y_in, z, z_in, p_z, p_y, parts_y, parts_z = res[:11]
            engine.attach_partials(logp_y, parts_y)
            engine.attach_partials(logp_z, parts_z)
            out = {"x_hat": x_hat, "y": y, "y_in": y_in, "z": z, "z_in": z_in, "p_z": p_z, "logp_z": logp_z, "p_y": p_y,
                   "logp_y": logp_y, "training": training}
            if not lean:
                out.update(dict(zip(("mu", "sigma") if self.K == 1 else ("weights", "mus", "sigmas"), res[11:])))
            return out
        prec_up = {"fp32": "fp32", "mixed": "fp32", "bf16x3": "bf16x3", "bf16": "bf16"}[self.precision]   # g_a, h_a
        prec = {"fp32": "fp32", "mixed": "bf16", "bf16x3": "bf16x3", "bf16": "bf16"}[self.precision]    # h_s, context, entropy parameters, g_s
        adt = engine.act_dtype(prec)
        pair = prec == "bf16x3"                      # activations are bf16 hi/lo pairs: [hi(c) | lo(c)] channels
        cw = 2 if pair else 1
        M, K = self.M, self.K
        x = x.contiguous().float()
        hy, wy, hz, wz = H // 16, W // 16, H // 64, W // 64
        with torch.cuda.device(x.device), torch.no_grad():
            noise_z, noise_y = _pair_noise(self, x, training, noise)
            y, y_in, y_in_nhwc, y_src = _pair_g_a(self, x, training, noise_y, prec_up, prec)

            # ---- two parallel branches while the pass is captured in a CUDA graph: g_s (needs y_in only) and the context conv
            # (y_in only) run beside h_a -> h_s; the context conv joins before the 1x1 stack, g_s at the end
            branch, main_stream = _branch_stream(x.device), torch.cuda.current_stream(x.device)
            if branch is not None:
                branch.wait_stream(main_stream)
                with torch.cuda.stream(branch):
                    x_hat = _pair_g_s(self, y_in_nhwc, B, hy, wy, prec)
            # phi = combined[..., 0:2M] (context), psi = combined[..., 2M:4M] (h_s): torch.cat([phi, psi]) never happens.
            # (the quantised symbols split into bf16 pairs with an all-zero lo half - the hand-off kernel checks and flags it on the
            #  device: their first consumers - h_s layer 1, the context conv, g_s layer 1 - run 2 of the 3 MMA passes)
            combined = torch.empty((B, hy, wy, cw * 4 * M), dtype=adt, device=x.device)
            self.context_model.masked.apply_mask_()
            branch2 = _branch_stream(x.device, 1)
            if branch2 is not None:
                branch2.wait_stream(main_stream)
            # eager launches (no graph branches): context conv + 1x1 stack + likelihood kernel go out through ONE C-ABI call after
            # h_s (nic_ctx_ep_fwd; same kernels, same results).  While a graph is captured the context conv is its own branch instead.
            one_call = branch2 is None and os.environ.get("NIC_CTX_EP_ONE_CALL", "1") != "0"
            y_flag = getattr(y_in_nhwc, "_nic_lo_flag", None)
            if not one_call:
                with (torch.cuda.stream(branch2) if branch2 is not None else contextlib.nullcontext()):
                    self.context_model.masked._op.run(y_in_nhwc, B, hy, wy, prec, out=combined, out_c_total=4 * M, out_c_offset=0,
                                                      in_lo_flag=y_flag)
            z, z_in, z_in_nhwc = _pair_h_a(self, y_src, B, hy, wy, training, noise_z, prec_up, prec)
            _pair_h_s_into(self, z_in_nhwc, B, hz, wz, prec, combined, 4 * M, 2 * M)
            if branch2 is not None:
                main_stream.wait_stream(branch2)

            ep = self.entropy_parameters.ops
            if one_call:
                key = (B, hy, wy, prec, bool(lean), str(x.device))
                plans = self.__dict__.setdefault("_ctx_ep_plans", {})
                if key not in plans:
                    plans.clear()                                # one shape at a time: the plan owns the two hidden activations
                    plans[key] = engine.CtxEpPlan(self.context_model.masked._op, ep, B, hy, wy, prec, M, K, x.device, full=not lean)
                raw, ly = plans[key].run(y_in_nhwc, combined, y_in=y_in, qmode=Q_PASSTHRU, in_lo_flag=y_flag)
            else:
                # ---- entropy parameters (1x1 stack) ---------------------------------------------------
                a = ep[0].run(combined, B, hy, wy, prec)
                a = ep[1].run(a, B, hy, wy, prec)
                raw = ep[2].run(a, B, hy, wy, prec, out_layout=LAYOUT_NCHW, out_dtype=torch.float32)
                # ---- likelihoods -------------------------------------------------------------------------
                ly = gm_likelihood(y_in, raw, M, K, Q_PASSTHRU, full=not lean, want_y_in=False)
            _, p_z, logp_z, parts_z = self.factorized_entropy_model.likelihood(z_in, Q_PASSTHRU)
            p_y, logp_y = ly["p"], ly["logp"]

            # ---- g_s -----------------------------------------------------------------------------------
            if branch is not None:
                main_stream.wait_stream(branch)                 # join
            else:
                x_hat = _pair_g_s(self, y_in_nhwc, B, hy, wy, prec)

        # per-image partial sums of logp ride along for rd_loss (RateDistortionLoss.py:13-14)
        engine.attach_partials(logp_y, ly["partials"])
        engine.attach_partials(logp_z, parts_z)
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


class HierarchicalMixtureResidual(nn.Module):
    """``HierarchicalMixtureResidual`` of /root/reference/Models.py:109-205: the same entropy side and output dict as
    JointAutoregressiveHierarchical with the 3x3 residual transforms (Encoder3x3 / Decoder3x3 / HyperEncoder3x3 / HyperDecoder3x3,
    Components.py:20-122, blocks of Layers.py).  Same constructor, attributes and ``state_dict`` keys.

    The transforms run layer by layer on the conv engine (f32 NHWC tensors between layers; residual sums by nic_add_inplace),
    the entropy side on the same kernels as the 5x5 model.  With autograd enabled and training=True the call is the training
    step's forward (training.train_forward: one autograd node; backward = training.Tape over the same C-ABI backward kernels
    as the 5x5 model).  precision: "bf16x3" (default when latent_channels is a multiple of 64: tensor cores with hi/lo-
    split operands; the RGB-input convs of the first block on conv_smallcin_kernel in fp32) or "fp32"."""

    def __init__(self, latent_channels: int = 192, K: int = 1, *, precision: Optional[str] = None):
        super().__init__()
        if not isinstance(latent_channels, int) or latent_channels < 1:
            raise ValueError(f"latent_channels must be int >= 1, got {latent_channels}")
        if not isinstance(K, int) or K < 1:
            raise ValueError(f"K must be int >= 1, got {K}")
        from .Components import Decoder3x3, Encoder3x3, HyperDecoder3x3, HyperEncoder3x3
        self.M = latent_channels
        self.K = K
        self.H = latent_channels
        self.distribution = "Mean-Scale Gaussian" if K == 1 else "Mixture of Gaussians"
        self.conditional = GaussianConditional() if K == 1 else GaussianMixtureConditional()
        # construction order follows Models.py:133-146 so a seeded build draws the reference's initial weights
        self.encoder = Encoder3x3(latent_channels=self.M)
        self.decoder = Decoder3x3(latent_channels=self.M)
        self.hyper_encoder = HyperEncoder3x3(latent_channels=self.M)
        self.hyper_decoder = HyperDecoder3x3(latent_channels=self.M)
        self.factorized_entropy_model = FactorizedEntropyBottleneck(self.M)
        self.context_model = ContextModel(latent_channels=self.M)
        self.entropy_parameters = EntropyParameters(latent_channels=self.M, hyper_latent_channels=self.H, K=self.K)
        precision = precision or engine.DEFAULT_PRECISION
        self.precision = ("bf16x3" if self.M % 64 == 0 else "fp32") if precision == "auto" else precision
        if self.precision not in ("fp32", "bf16x3") or (self.precision == "bf16x3" and self.M % 64):
            raise ValueError("HierarchicalMixtureResidual: precision must be 'fp32', or 'bf16x3' with latent_channels % 64 == 0")

    def forward(self, x: torch.Tensor, training: bool = True, *, noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                lean: bool = False):
        engine.require_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected x of shape [B, 3, H, W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 64 or W % 64:
            raise ValueError(f"H and W must be multiples of 64; got {H}x{W}")
        from . import training as T
        if training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # the training step: ONE autograd node over the hand-written backward (training.Tape walks the residual graphs)
            return T.train_forward(self, x, noise=noise, lean=lean)
        arm, M, K = self.precision, self.M, self.K
        hy, wy, hz, wz = H // 16, W // 16, H // 64, W // 64
        x = x.contiguous().float()
        with torch.cuda.device(x.device), torch.no_grad():
            noise_z = noise_y = None
            if training:
                noise_z, noise_y = noise if noise is not None else (torch.rand((B, M, hz, wz), device=x.device) - 0.5,
                                                                    torch.rand((B, M, hy, wy), device=x.device) - 0.5)
            qmode = Q_NOISE if training else Q_ROUND
            y_nhwc, _, _ = self.encoder.run_nhwc(x, B, H, W, arm, in_layout=LAYOUT_NCHW)
            y, y_in, y_in_nhwc, _ = engine.latent_handoff(y_nhwc, qmode, noise_y, torch.float32)
            z_nhwc, _, _ = self.hyper_encoder.run_nhwc(y_nhwc, B, hy, wy, arm)
            z, z_in, z_in_nhwc, _ = engine.latent_handoff(z_nhwc, qmode, noise_z, torch.float32)
            psi, _, _ = self.hyper_decoder.run_nhwc(z_in_nhwc, B, hz, wz, arm)
            combined = torch.empty((B, hy, wy, 4 * M), dtype=torch.float32, device=x.device)
            combined[..., 2 * M:] = psi                      # torch.cat([phi, psi]) of Models.py:171: phi is written in place below
            masked = self.context_model.masked
            masked.apply_mask_()
            T.conv_forward(arm, masked, engine.EPI_BIAS, y_in_nhwc, B, hy, wy, out=combined, out_c_total=4 * M, out_c_offset=0,
                           mask_a=masked._op.mask_a)
            ep = self.entropy_parameters.ops
            a = T.conv_forward(arm, ep[0].conv, ep[0].epilogue, combined, B, hy, wy)
            a = T.conv_forward(arm, ep[1].conv, ep[1].epilogue, a, B, hy, wy)
            raw = T.conv_forward(arm, ep[2].conv, ep[2].epilogue, a, B, hy, wy, out_layout=LAYOUT_NCHW)
            T.forget_pairs()
            ly = gm_likelihood(y_in, raw, M, K, Q_PASSTHRU, full=not lean, want_y_in=False)
            _, p_z, logp_z, parts_z = self.factorized_entropy_model.likelihood(z_in, Q_PASSTHRU)
            x_nhwc, _, _ = self.decoder.run_nhwc(y_in_nhwc, B, hy, wy, arm)
            x_hat = x_nhwc.permute(0, 3, 1, 2).contiguous()
        p_y, logp_y = ly["p"], ly["logp"]
        engine.attach_partials(logp_y, ly["partials"])
        engine.attach_partials(logp_z, parts_z)
        out = {"x_hat": x_hat, "y": y, "y_in": y_in, "z": z, "z_in": z_in, "p_z": p_z, "logp_z": logp_z, "p_y": p_y, "logp_y": logp_y,
               "training": training}
        if not lean:
            out.update({"mu": ly["mu"], "sigma": ly["sigma"]} if K == 1 else
                       {"weights": ly["weights"], "mus": ly["mus"], "sigmas": ly["sigmas"]})
        return out


class ScalableImageCoding(nn.Module):
    """The scalable-coding variant of /root/reference/Models.py:208-338: the JointAutoregressiveHierarchical trunk at M
    channels with y split into a base part y1 (M1 channels) and an enhancement part y2 (M - M1), each with its own
    (context model, entropy-parameter net, conditional), both heads sharing the hyper features psi.

    The reference's forward raises as committed; this class implements the evident intent,
    i.e. the reference with these four repairs, and nothing else:
      * ``self.factorized_entropy_model(z_in, debug)`` (Models.py:302)  ->  ``self.factorized_entropy_model(z_in)``
      * K == 1: ``conditional(y1, mu1=, sigma1=)`` (:293-294, :305-306)  ->  ``conditional(y1, mu=mu1, sigma=sigma1)``
      * K > 1: ``params1`` assigned twice (:298-299)                      ->  ``params2`` holds the second head
      * ``LatentSpaceTransform`` (:256, :319; inconsistent channel counts, Components.py:129-132) is NOT built: the output
        dict has no 'F_tilde', and reference checkpoints load with their ``LST.*`` entries ignored.
    Parity for this class is against the reference's own sub-modules called in that repaired order
    (oracle/make_golden.py: scalable cases) - "parity unpinned" by the reference itself, which has no runnable forward.
    Arithmetic: precision="bf16x3" (default for channel counts that are multiples of 64) runs the convolutions and the GDN
    contractions on the tensor cores, "fp32" on the CUDA cores; both meet the parity bar (tests/test_gpu_scalable.py).
    """

    def __init__(self, latent_channels: int = 192, base_channels: int = 128, K: int = 1, *, precision: Optional[str] = None):
        super().__init__()
        if not isinstance(latent_channels, int) or latent_channels < 1:
            raise ValueError(f"latent_channels must be int >= 1, got {latent_channels}")
        if not isinstance(K, int) or K < 1:
            raise ValueError(f"K must be int >= 1, got {K}")
        if not isinstance(base_channels, int) or not 0 < base_channels < latent_channels:
            raise ValueError(f"base_channels must be an int in (0, latent_channels), got {base_channels}")
        self.M, self.M1, self.M2 = latent_channels, base_channels, latent_channels - base_channels
        self.H, self.K = latent_channels, K
        self.distribution = "Mean-Scale Gaussian" if K == 1 else "Mixture of Gaussians"
        self.conditional = GaussianConditional() if K == 1 else GaussianMixtureConditional()
        # construction order follows Models.py:237-254 so a seeded build draws the reference's initial weights
        self.encoder = Encoder5x5(latent_channels=self.M)
        self.decoder = Decoder5x5(latent_channels=self.M)
        self.hyper_encoder = HyperEncoder5x5(latent_channels=self.M)
        self.hyper_decoder = HyperDecoder5x5(latent_channels=self.M)
        self.factorized_entropy_model = FactorizedEntropyBottleneck(self.M)
        self.context_model_1 = ContextModel(latent_channels=self.M1)
        self.context_model_2 = ContextModel(latent_channels=self.M2)
        self.entropy_parameters_1 = EntropyParameters(latent_channels=self.M1, hyper_latent_channels=self.H, K=self.K)
        self.entropy_parameters_2 = EntropyParameters(latent_channels=self.M2, hyper_latent_channels=self.H, K=self.K)
        # "bf16x3" (default when every channel count is a multiple of 64, e.g. the reference's 192 / 128): every convolution
        # and the GDN channel contractions on the tcgen05 engine with hi/lo-split operands (fp32 grade), fp32 NHWC tensors
        # between layers split on the fly - the layer-by-layer form the training step uses (training.conv_forward / gdn_forward);
        # "fp32": CUDA cores.
        tc_ok = all(c % 64 == 0 for c in (self.M, self.M1, self.M2))
        precision = precision or engine.DEFAULT_PRECISION
        self.precision = ("bf16x3" if tc_ok else "fp32") if precision == "auto" else precision
        if self.precision not in ("fp32", "bf16x3") or (self.precision == "bf16x3" and not tc_ok):
            raise ValueError("ScalableImageCoding: precision must be 'fp32', or 'bf16x3' with channel counts that are multiples of 64")

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        return super().load_state_dict({k: v for k, v in state_dict.items() if not k.startswith("LST.")}, strict=strict, **kwargs)

    def _forward_pairs(self, x, training, noise):
        """The bf16x3 arm on the fused pipeline of JointAutoregressiveHierarchical (conv -> GDN kernel, activations as bf16 hi/lo
        pairs, no fp32 round trips): the shared trunk, then the two (context, entropy-parameter, likelihood) heads over psi."""
        B, _, H, W = x.shape
        M, M1, M2, K = self.M, self.M1, self.M2, self.K
        prec = "bf16x3"
        x = x.contiguous().float()
        hy, wy, hz, wz = H // 16, W // 16, H // 64, W // 64
        with torch.cuda.device(x.device), torch.no_grad():
            noise_z, noise_y = _pair_noise(self, x, training, noise)
            y, y_in, y_in_nhwc, y_src = _pair_g_a(self, x, training, noise_y, prec, prec)
            branch, main_stream = _branch_stream(x.device), torch.cuda.current_stream(x.device)
            if branch is not None:                                        # g_s as a parallel branch of a captured graph
                branch.wait_stream(main_stream)
                with torch.cuda.stream(branch):
                    x_hat = _pair_g_s(self, y_in_nhwc, B, hy, wy, prec)
            z, z_in, z_in_nhwc = _pair_h_a(self, y_src, B, hy, wy, training, noise_z, prec, prec)
            y_flag = getattr(y_in_nhwc, "_nic_lo_flag", None)
            c1, c2 = 2 * M1 + 2 * M, 2 * M2 + 2 * M                       # [phi_i (2 M_i) | psi (2 M)] channels of the two heads
            comb1 = torch.empty((B, hy, wy, 2 * c1), dtype=torch.bfloat16, device=x.device)      # pair tensors: [hi(c) | lo(c)]
            comb2 = torch.empty((B, hy, wy, 2 * c2), dtype=torch.bfloat16, device=x.device)
            _pair_h_s_into(self, z_in_nhwc, B, hz, wz, prec, comb1, c1, 2 * M1)                  # psi once, into head 1's buffer
            comb2[..., 2 * M2:c2] = comb1[..., 2 * M1:c1]                                        # ... head 2 gets a copy: hi half,
            comb2[..., c2 + 2 * M2:] = comb1[..., c1 + 2 * M1:]                                  # lo half
            y1, y2 = torch.split(y_in, [M1, M2], dim=1)                   # Models.py:279 (views, as in the reference)
            heads = []
            for ctx, ep, comb, ct, lo, hi_, yi in ((self.context_model_1, self.entropy_parameters_1, comb1, c1, 0, M1, y1),
                                                   (self.context_model_2, self.entropy_parameters_2, comb2, c2, M1, M, y2)):
                mi = hi_ - lo
                yi_pair = torch.cat([y_in_nhwc[..., lo:hi_], y_in_nhwc[..., M + lo:M + hi_]], dim=-1).contiguous()
                ctx.masked.apply_mask_()
                ctx.masked._op.run(yi_pair, B, hy, wy, prec, out=comb, out_c_total=ct, out_c_offset=0, in_lo_flag=y_flag)
                a = ep.ops[0].run(comb, B, hy, wy, prec)
                a = ep.ops[1].run(a, B, hy, wy, prec)
                raw = ep.ops[2].run(a, B, hy, wy, prec, out_layout=LAYOUT_NCHW, out_dtype=torch.float32)
                heads.append(gm_likelihood(yi.contiguous(), raw, mi, K, Q_PASSTHRU, full=True, want_y_in=False))
            _, p_z, logp_z, parts_z = self.factorized_entropy_model.likelihood(z_in, Q_PASSTHRU)
            if branch is not None:
                main_stream.wait_stream(branch)
            else:
                x_hat = _pair_g_s(self, y_in_nhwc, B, hy, wy, prec)
        l1, l2 = heads
        engine.attach_partials(l1["logp"], l1["partials"])
        engine.attach_partials(l2["logp"], l2["partials"])
        engine.attach_partials(logp_z, parts_z)
        out = {
            "x_hat": x_hat, "y": y, "y_in": y_in, "y1": y1, "y2": y2, "z": z, "z_in": z_in, "p_z": p_z, "logp_z": logp_z,
            "p_y1": l1["p"], "logp_y1": l1["logp"], "p_y2": l2["p"], "logp_y2": l2["logp"], "training": training,
        }
        if K == 1:
            out.update({"mu1": l1["mu"], "sigma1": l1["sigma"], "mu2": l2["mu"], "sigma2": l2["sigma"]})
        else:
            out.update({"weights1": l1["weights"], "mus1": l1["mus"], "sigmas1": l1["sigmas"],
                        "weights2": l2["weights"], "mus2": l2["mus"], "sigmas2": l2["sigmas"]})
        return out

    def forward(self, x: torch.Tensor, training: bool = True, debug=False, *,
                noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        engine.require_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected x of shape [B, 3, H, W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 64 or W % 64:
            raise ValueError(f"H and W must be multiples of 64; got {H}x{W}")
        if training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # the training call with autograd on: one autograd node over the layer-by-layer forward (training_scalable.py)
            from . import training_scalable as _ts
            return _ts.train_forward(self, x, noise=noise)
        arm, M, M1, M2, K = self.precision, self.M, self.M1, self.M2, self.K
        if arm == "bf16x3" and M in (128, 192) and M1 % 64 == 0 and M2 % 64 == 0:
            return self._forward_pairs(x, training, noise)
        from . import training as T

        def layer(op, a, h, w, **kw):
            """One conv (+ GDN / LeakyReLU) with f32 tensors at both ends."""
            if arm == "fp32":
                return op.run(a, B, h, w, "fp32", **kw)
            u = T.conv_forward(arm, op.conv, engine.EPI_BIAS if op.gdn is not None else op.epilogue, a, B, h, w,
                               mask_a=op.mask_a, **kw)
            if op.gdn is not None:
                ho, wo = engine.conv_out_hw(op.conv, h, w)
                u = T.gdn_forward(arm, op.gdn, u, B, ho, wo)[0]
            T.forget_pairs()                     # the split copies are per layer here (the training step keeps them for its backward)
            return u
        x = x.contiguous().float()
        hy, wy, hz, wz = H // 16, W // 16, H // 64, W // 64
        with torch.cuda.device(x.device), torch.no_grad():
            noise_z = noise_y = None
            if training:
                noise_z, noise_y = noise if noise is not None else (torch.rand((B, M, hz, wz), device=x.device) - 0.5,
                                                                    torch.rand((B, M, hy, wy), device=x.device) - 0.5)
            qmode = Q_NOISE if training else Q_ROUND
            a, h, w, layout = x, H, W, LAYOUT_NCHW
            for op in self.encoder.ops:
                a = layer(op, a, h, w, in_layout=layout, out_layout=LAYOUT_NHWC)
                h, w = engine.conv_out_hw(op.conv, h, w)
                layout = LAYOUT_NHWC
            y_nhwc = a
            y, y_in, y_in_nhwc, _ = engine.latent_handoff(y_nhwc, qmode, noise_y, torch.float32)
            a, h, w = y_nhwc, hy, wy
            for op in self.hyper_encoder.ops:
                a = layer(op, a, h, w)
                h, w = engine.conv_out_hw(op.conv, h, w)
            z, z_in, z_in_nhwc, _ = engine.latent_handoff(a, qmode, noise_z, torch.float32)
            # psi once, into head 1's concat buffer [phi1 (2 M1) | psi (2 M)]; head 2's buffer gets a copy of the window
            comb1 = torch.empty((B, hy, wy, 2 * M1 + 2 * M), dtype=torch.float32, device=x.device)
            comb2 = torch.empty((B, hy, wy, 2 * M2 + 2 * M), dtype=torch.float32, device=x.device)
            a, h, w = z_in_nhwc, hz, wz
            hs = self.hyper_decoder.ops
            for i, op in enumerate(hs):
                if i == len(hs) - 1:
                    layer(op, a, h, w, out=comb1, out_c_total=2 * M1 + 2 * M, out_c_offset=2 * M1)
                else:
                    a = layer(op, a, h, w)
                h, w = engine.conv_out_hw(op.conv, h, w)
            comb2[..., 2 * M2:] = comb1[..., 2 * M1:]
            y1, y2 = torch.split(y_in, [M1, M2], dim=1)                   # Models.py:279 (views, as in the reference)
            y1c, y2c = y1.contiguous(), y2.contiguous()
            heads = []
            for ctx, ep, comb, mi, yi_nhwc, yi in ((self.context_model_1, self.entropy_parameters_1, comb1, M1,
                                                    y_in_nhwc[..., :M1].contiguous(), y1c),
                                                   (self.context_model_2, self.entropy_parameters_2, comb2, M2,
                                                    y_in_nhwc[..., M1:].contiguous(), y2c)):
                ctx.masked.apply_mask_()
                layer(ctx.masked._op, yi_nhwc, hy, wy, out=comb, out_c_total=comb.shape[-1], out_c_offset=0)
                a = layer(ep.ops[0], comb, hy, wy)
                a = layer(ep.ops[1], a, hy, wy)
                raw = layer(ep.ops[2], a, hy, wy, out_layout=LAYOUT_NCHW)
                heads.append(gm_likelihood(yi, raw, mi, K, Q_PASSTHRU, full=True, want_y_in=False))
            _, p_z, logp_z, parts_z = self.factorized_entropy_model.likelihood(z_in, Q_PASSTHRU)
            a, h, w = y_in_nhwc, hy, wy
            dec = self.decoder.ops
            for i, op in enumerate(dec):
                last = i == len(dec) - 1
                a = layer(op, a, h, w, out_layout=LAYOUT_NCHW if last else LAYOUT_NHWC)
                h, w = engine.conv_out_hw(op.conv, h, w)
            x_hat = a
        l1, l2 = heads
        engine.attach_partials(l1["logp"], l1["partials"])
        engine.attach_partials(l2["logp"], l2["partials"])
        engine.attach_partials(logp_z, parts_z)
        out = {
            "x_hat": x_hat, "y": y, "y_in": y_in, "y1": y1, "y2": y2, "z": z, "z_in": z_in, "p_z": p_z, "logp_z": logp_z,
            "p_y1": l1["p"], "logp_y1": l1["logp"], "p_y2": l2["p"], "logp_y2": l2["logp"], "training": training,
        }
        if K == 1:
            out.update({"mu1": l1["mu"], "sigma1": l1["sigma"], "mu2": l2["mu"], "sigma2": l2["sigma"]})
        else:
            out.update({"weights1": l1["weights"], "mus1": l1["mus"], "sigmas1": l1["sigmas"],
                        "weights2": l2["weights"], "mus2": l2["mus"], "sigmas2": l2["sigmas"]})
        return out
