"""CPU oracle of the TRAINING STEP (forward with injected noise, rd_loss, backward, one Adam update).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference has no backward code of its own: its gradients are whatever
torch autograd derives from the forward of Models.py:49-106 + RateDistortionLoss.py:5-49, with two hand-written rules on
the way: compressai's ``LowerBound`` gradient (oracle/gdn.py) and the masked conv's in-place ``weight.data *= mask``
(ContextModels.py:19: the gradient reaches all 25 taps).  So the oracle is torch autograd over oracle/forward.py run in
differentiable mode.  The optimizer is ``torch.optim.Adam(model.parameters(), lr=1e-4)`` (Main.ipynb:133, Trainer.py:85-86).

Parity pin: ``oracle/make_golden.py train`` runs the REAL reference model (training=True, torch.rand_like noise reproduced
from the seed), its own rd_loss, ``loss.backward()`` and one Adam step, and commits per-parameter gradient norms + sampled
entries + the updated parameters' samples under tests/golden/c4_train_*.npz; tests/test_oracle.py checks this file against them.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import forward as O

BUFFER_SUFFIXES = (".mask", ".pedestal", ".bound")


def parameter_keys(sd):
    """state_dict keys that are nn.Parameters in the reference (buffers: masks and the GDN reparam constants)."""
    return [k for k in sd.keys() if not k.endswith(BUFFER_SUFFIXES)]


def loss_and_grads(sd, x, M: int, K: int, noise_z: torch.Tensor, noise_y: torch.Tensor, lambda_rd: float,
                   dtype=torch.float32) -> Tuple[Dict[str, float], Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Returns (rd dict with 'loss' as float, {param key: dLoss/dparam}, forward dict (detached))."""
    leaves = {k: (v.detach().to("cpu", dtype).clone().requires_grad_(True) if k in set(parameter_keys(sd)) else v.detach().clone())
              for k, v in sd.items()}
    prev = O.DIFFERENTIABLE
    O.DIFFERENTIABLE = True
    try:
        out = O.forward(leaves, x, M, K, training=True, noise_z=noise_z, noise_y=noise_y, dtype=dtype)
        rd = O.rd_loss(out, x, lambda_rd)
        rd["loss"].backward()
    finally:
        O.DIFFERENTIABLE = prev
    grads = {k: leaves[k].grad.detach() for k in parameter_keys(sd) if leaves[k].grad is not None}
    rd = dict(rd)
    rd["loss"] = float(rd["loss"].detach())
    return rd, grads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}


def loss_and_grads_scalable(sd, x, M: int, M1: int, K: int, noise_z: torch.Tensor, noise_y: torch.Tensor, lambda_rd: float,
                            dtype=torch.float32):
    """The same for ScalableImageCoding: autograd over forward_scalable (the reference's forward with the repairs of SURVEY.md
    section 2.4) + vision_rd_loss (RateDistortionLoss.py:52-121, V = None).  Returns (rd dict, grads)."""
    keys = set(parameter_keys(sd))
    leaves = {k: (v.detach().to("cpu", dtype).clone().requires_grad_(True) if k in keys else v.detach().clone()) for k, v in sd.items()}
    prev = O.DIFFERENTIABLE
    O.DIFFERENTIABLE = True
    try:
        out = O.forward_scalable(leaves, x, M, M1, K, training=True, noise_z=noise_z, noise_y=noise_y, dtype=dtype)
        rd = O.vision_rd_loss(out, x, lambda_rd)
        rd["loss"].backward()
    finally:
        O.DIFFERENTIABLE = prev
    grads = {k: leaves[k].grad.detach() for k in keys if leaves[k].grad is not None}
    rd = dict(rd)
    rd["loss"] = float(rd["loss"].detach())
    return rd, grads


def loss_and_grads_residual(sd, x, M: int, K: int, noise_z: torch.Tensor, noise_y: torch.Tensor, lambda_rd: float, dtype=torch.float32):
    """The same for HierarchicalMixtureResidual (Models.py:109-205, the 3x3 residual transforms of Layers.py / Components.py:20-122):
    autograd over forward_residual + rd_loss.  Pinned by tests/golden/c6_train_*.npz (``oracle/make_golden.py residual-train``:
    the reference's own unmodified class, its rd_loss, loss.backward() and one Adam step).  Returns (rd dict, grads, forward dict)."""
    keys = set(parameter_keys(sd))
    leaves = {k: (v.detach().to("cpu", dtype).clone().requires_grad_(True) if k in keys else v.detach().clone()) for k, v in sd.items()}
    prev = O.DIFFERENTIABLE
    O.DIFFERENTIABLE = True
    try:
        out = O.forward_residual(leaves, x, M, K, training=True, noise_z=noise_z, noise_y=noise_y, dtype=dtype)
        rd = O.rd_loss(out, x, lambda_rd)
        rd["loss"].backward()
    finally:
        O.DIFFERENTIABLE = prev
    grads = {k: leaves[k].grad.detach() for k in parameter_keys(sd) if leaves[k].grad is not None}
    rd = dict(rd)
    rd["loss"] = float(rd["loss"].detach())
    return rd, grads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}


def kink_margin(sd, x, M: int, K: int, noise_z: torch.Tensor, noise_y: torch.Tensor) -> float:
    """Smallest |pre-activation| over the six LeakyReLU layers of the step (h_a 0/2, h_s 0/2, entropy-parameter 0/2), in float64.
    LeakyReLU's derivative jumps from 0.01 to 1 at 0, so an element whose pre-activation is below the fp32 accumulation noise
    (~1e-7 here) gets either slope depending on summation order - the reference's own CPU and CUDA runs would disagree there,
    and every gradient upstream of that layer moves by ~1e-3 relative.  Parity cases are chosen with a margin (>= 1e-6)."""
    import torch.nn.functional as F
    dt = torch.float64
    P = lambda k: sd[k].detach().to("cpu", dt)     # noqa: E731
    x = x.to("cpu", dt)
    y = O.analysis(sd, x, dt)
    margins = []

    def act(h):
        margins.append(float(h.abs().min()))
        return F.leaky_relu(h, 0.01)
    h = act(F.conv2d(y, P("hyper_encoder.net.0.weight"), P("hyper_encoder.net.0.bias"), padding=1))
    h = act(F.conv2d(h, P("hyper_encoder.net.2.weight"), P("hyper_encoder.net.2.bias"), stride=2, padding=2))
    z = F.conv2d(h, P("hyper_encoder.net.4.weight"), P("hyper_encoder.net.4.bias"), stride=2, padding=2)
    z_in, y_in = z + noise_z.to(dt), y + noise_y.to(dt)
    h = act(F.conv_transpose2d(z_in, P("hyper_decoder.net.0.weight"), P("hyper_decoder.net.0.bias"), stride=2, padding=2, output_padding=1))
    h = act(F.conv_transpose2d(h, P("hyper_decoder.net.2.weight"), P("hyper_decoder.net.2.bias"), stride=2, padding=2, output_padding=1))
    psi = F.conv2d(h, P("hyper_decoder.net.4.weight"), P("hyper_decoder.net.4.bias"), padding=1)
    h = torch.cat([O.context(sd, y_in, dt), psi], dim=1)
    h = act(F.conv2d(h, P("entropy_parameters.net.0.weight"), P("entropy_parameters.net.0.bias")))
    act(F.conv2d(h, P("entropy_parameters.net.2.weight"), P("entropy_parameters.net.2.bias")))
    return min(margins)


def noise_with_margin(sd, x, M: int, K: int, first_seed: int, margin: float = 1e-6, tries: int = 64):
    """(seed, noise_z, noise_y) of the first seed >= first_seed whose step keeps every LeakyReLU pre-activation `margin` away from 0.
    Noise is drawn as the reference draws it: z's first, then y's, from the global generator (Models.py:57-58)."""
    B, _, H, W = x.shape
    for seed in range(first_seed, first_seed + tries):
        torch.manual_seed(seed)
        nz = torch.rand(B, M, H // 64, W // 64) - 0.5
        ny = torch.rand(B, M, H // 16, W // 16) - 0.5
        if kink_margin(sd, x, M, K, nz, ny) >= margin:
            return seed, nz, ny
    raise RuntimeError("no noise draw with the requested LeakyReLU margin")


def adam_step(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: Optional[dict] = None, lr: float = 1e-4,
              betas=(0.9, 0.999), eps: float = 1e-8):
    """One torch.optim.Adam update (defaults of Main.ipynb:133: no weight decay, no amsgrad), restated:
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)."""
    state = state if state is not None else {"step": 0, "m": {}, "v": {}}
    state["step"] += 1
    t = state["step"]
    out = {}
    for k, g in grads.items():
        p = params[k].detach().to(g.dtype)
        m = state["m"].get(k, torch.zeros_like(p)).mul(betas[0]).add(g, alpha=1 - betas[0])
        v = state["v"].get(k, torch.zeros_like(p)).mul(betas[1]).addcmul(g, g, value=1 - betas[1])
        state["m"][k], state["v"][k] = m, v
        denom = v.sqrt() / (1 - betas[1] ** t) ** 0.5 + eps
        out[k] = p - (lr / (1 - betas[0] ** t)) * m / denom
    return out, state
