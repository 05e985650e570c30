"""Shared pieces of the parity tests: seeded weights, golden loading, tolerance definitions."""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# seeded weight sets, see oracle/make_golden.py: (gain on last g_a conv, gain on last h_a conv, + on sigma biases)
INITS = {"plain": (1.0, 1.0, 0.0), "gain": (136.2, 3.86, 0.0), "calib": (34.0, 3.86, 3.0),
         # M = 192: the 128-channel calib gains leave 1.5 % of p_y at the 1e-9 clamp (gradients there are rounding noise over 1e-9:
         # the reference's own fp32 and fp64 gradients differ by 30 %); this set keeps min p_y > 1e-5 (fp32-vs-fp64 spread 5e-5)
         "calib192": (16.0, 3.86, 4.0)}


def golden_cases():
    """JointAutoregressiveHierarchical cases (configs 1-3)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "c[1-3]*.npz")))


def train_cases():
    """Training-step cases (config 4: the reference's forward + rd_loss + backward + one Adam step, oracle/make_golden.py train)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "c4_train_*.npz")))


def sample_index(numel: int, n: int = 48):
    return np.unique(np.linspace(0, numel - 1, num=min(n, numel)).astype(np.int64))


def scalable_cases():
    """ScalableImageCoding cases (config 5; the reference's sub-modules in the repaired order, see oracle/make_golden.py)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "c5_*.npz")))


def seeded_scalable_model(M, M1, K, init="calib", precision="fp32"):
    from neural_image_compression_b200.Models import ScalableImageCoding
    torch.manual_seed(0)
    model = ScalableImageCoding(M, M1, K=K, precision=precision)
    if init != "plain":
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        gy, gz, sigma_bias = INITS[init]
        for k in ("encoder.net.6.weight", "encoder.net.6.bias"):
            sd[k] = sd[k] * gy
        for k in ("hyper_encoder.net.4.weight", "hyper_encoder.net.4.bias"):
            sd[k] = sd[k] * gz
        if sigma_bias:
            for head in ("entropy_parameters_1", "entropy_parameters_2"):
                b = sd[f"{head}.net.4.bias"].clone()
                n = b.numel()
                b[(n // 2 if n % 3 else 2 * n // 3):] += sigma_bias
                sd[f"{head}.net.4.bias"] = b
        model.load_state_dict(sd)
    return model


def residual_cases():
    """HierarchicalMixtureResidual cases (the 3x3 residual family; oracle/make_golden.py residual)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "c6_res3x3_*.npz")))


def residual_train_cases():
    """Training step of HierarchicalMixtureResidual (the reference's own class: forward, rd_loss, backward, Adam;
    oracle/make_golden.py residual-train)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "c6_train_*.npz")))


def seeded_residual_model(M, K, gain_y, gain_z, precision="fp32", sigma_bias=3.0):
    from neural_image_compression_b200.Models import HierarchicalMixtureResidual
    torch.manual_seed(0)
    model = HierarchicalMixtureResidual(M, K=K, precision=precision)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k in ("encoder.net.6.weight", "encoder.net.6.bias"):
        sd[k] = sd[k] * gain_y
    for k in ("hyper_encoder.net.8.weight", "hyper_encoder.net.8.bias"):
        sd[k] = sd[k] * gain_z
    b = sd["entropy_parameters.net.4.bias"].clone()
    n = b.numel()
    b[(n // 2 if n % 3 else 2 * n // 3):] += sigma_bias
    sd["entropy_parameters.net.4.bias"] = b
    model.load_state_dict(sd)
    return model


def residual_state_dict(M, K, gain_y, gain_z, sigma_bias=3.0):
    m = seeded_residual_model(M, K, gain_y, gain_z, sigma_bias=sigma_bias)
    return {k: v.clone() for k, v in m.state_dict().items()}


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode()); h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def apply_init(sd, init):
    gy, gz, sigma_bias = INITS[init]
    for k in ("encoder.net.6.weight", "encoder.net.6.bias"):
        sd[k] = sd[k] * gy
    for k in ("hyper_encoder.net.4.weight", "hyper_encoder.net.4.bias"):
        sd[k] = sd[k] * gz
    if sigma_bias:
        b = sd["entropy_parameters.net.4.bias"].clone()
        n = b.numel()
        start = n // 2 if n % 3 else 2 * n // 3          # [mu | sigma] (K = 1) or [w | mu | sigma]
        b[start:] += sigma_bias
        sd["entropy_parameters.net.4.bias"] = b
    return sd


def seeded_model(M, K, init="plain", precision="fp32"):
    """Product model with the weights the reference draws under torch.manual_seed(0), re-scaled per `init`."""
    from neural_image_compression_b200.Models import JointAutoregressiveHierarchical
    if isinstance(init, bool):
        init = "gain" if init else "plain"
    torch.manual_seed(0)
    model = JointAutoregressiveHierarchical(M, K=K, precision=precision)
    if init != "plain":
        model.load_state_dict(apply_init({k: v.clone() for k, v in model.state_dict().items()}, init))
    return model


def seeded_input(shape):
    torch.manual_seed(1)
    return torch.rand(*shape)


# ---- tolerances (north_star: symbols bit-exact, likelihoods 1e-4 relative, bpp / PSNR 1e-3) ----------------------
#
# Likelihoods: the reference evaluates Phi(u) - Phi(l) with Phi = 0.5 (1 + erf) in fp32, so every p carries an
# absolute rounding noise of a few ulp(1) = 1.2e-7 that no implementation can reproduce bit for bit (the reference's
# own CPU and CUDA runs differ by it).  The check is therefore |dp| <= 1e-4 * p + 4 ulp(1).
P_RTOL = 1e-4
P_ATOL = 4 * 1.1920929e-07
BPP_TOL = 1e-3
PSNR_TOL = 1e-3


def likelihood_close(p, p_ref):
    p, p_ref = np.asarray(p, np.float64), np.asarray(p_ref, np.float64)
    err = np.abs(p - p_ref)
    bad = err > P_RTOL * p_ref + P_ATOL
    return int(bad.sum()), float(err.max())


# ---- end-to-end comparison in the presence of rounding-tie flips --------------------------------------------------
#
# A symbol whose pre-rounding value sits within a few 1e-4 of a half-integer may legitimately round the other way in a pipeline
# with a different (but fp32-grade) accumulation order; the reference's own fp32 and fp64 runs flip 3 of 393 216 (SURVEY.md
# fact 7).  A flipped symbol changes every tensor downstream of it, so the per-element likelihood / x_hat comparison against
# the reference's vectors is made on the elements that are NOT a function of a flipped symbol (masks below), and a second
# comparison - the oracle's entropy path and synthesis transform evaluated on THIS run's symbols - covers 100 % of the elements.
TIE_TAU = 2e-3            # |frac(y_ref) - 0.5| below which a flip counts as a tie
# fp32-grade bound on max |y - y_ref| / max |y_ref| before the rounding (4 convs + 3 GDN deep; bf16x3 measures 4-5e-5, the fp32
# arm 3e-6), and on z likewise (3 more convs: bf16x3 measures up to 1.1e-4 on the 192-channel model)
PRE_RTOL = {"y": 1e-4, "z": 2e-4}


def tie_flip_bound(pre, pre_ref):
    """How many rounding-tie flips the measured pre-rounding error explains.  The fractional part of y is uniformly distributed, so a
    value perturbed by d rounds the other way with probability 2 |d|: the expected number of flips is 2 sum |y - y_ref|
    (calib weights, std(y) = 2: ~7e-5 of the symbols in the bf16x3 arm; gain weights, std(y) = 8: ~3e-4).  Bound: 3x that + 3
    (Poisson slack).  Together with PRE_RTOL this ties the flip count to an fp32-grade error on y instead of exempting a window."""
    return 6.0 * float(np.abs(np.asarray(pre, np.float64) - np.asarray(pre_ref, np.float64)).sum()) + 3.0


def _dilate(m, kh, kw):
    """binary [B,1,H,W] map: 1 where any 1 lies within the (kh x kw) window centred on the pixel"""
    return torch.nn.functional.max_pool2d(m, (kh, kw), stride=1, padding=(kh // 2, kw // 2))


def flip_masks(y_in, y_ref, z_in, z_ref, x_shape):
    """Boolean maps of the elements that do NOT depend on a flipped symbol:
    ok_y [B,1,hy,wy] for p_y (causal 5x5 footprint of the masked context conv, ContextModels.py:13-16, and the h_s receptive
    field of a flipped z: ConvT5 s2 -> ConvT5 s2 -> 3x3 = +-7 around 4 i), ok_z [B,C,hz,wz] for p_z (elementwise),
    ok_x [B,1,H,W] for x_hat (g_s: four ConvT 5x5 s2 = +-30 around 16 i)."""
    y_in, y_ref, z_in, z_ref = (torch.as_tensor(np.asarray(a)) for a in (y_in, y_ref, z_in, z_ref))
    fy = (y_in != y_ref).any(dim=1, keepdim=True).float()
    fz = (z_in != z_ref)
    k = torch.zeros(1, 1, 5, 5)
    k[0, 0, :2, :] = 1; k[0, 0, 2, :3] = 1          # taps the context conv reads + the symbol itself
    # output pixel (i', j') reads flip[i' + di, j' + dj] over the live taps: cross-correlation with the live-tap kernel
    bad_y = torch.nn.functional.conv2d(fy, k, padding=2) > 0
    fzp = fz.any(dim=1, keepdim=True).float()
    if fzp.sum() > 0:
        up = torch.zeros_like(fy)
        up[:, :, ::4, ::4] = fzp
        bad_y |= _dilate(up, 15, 15) > 0
    B, _, Hx, Wx = x_shape
    upx = torch.zeros(B, 1, Hx, Wx)
    upx[:, :, ::16, ::16] = fy
    bad_x = _dilate(upx, 61, 61) > 0 if fy.sum() > 0 else torch.zeros(B, 1, Hx, Wx, dtype=torch.bool)
    return (~bad_y).numpy(), (~fz).numpy(), (~bad_x).numpy()


def masked_likelihood_close(p, p_ref, ok):
    """(outliers, worst abs err, compared elements, compared fraction) over the elements where ok (broadcast over channels)"""
    p, p_ref = np.asarray(p, np.float64), np.asarray(p_ref, np.float64)
    ok = np.broadcast_to(ok, p.shape)
    err = np.abs(p - p_ref)
    bad = (err > P_RTOL * p_ref + P_ATOL) & ok
    n = int(ok.sum())
    return int(bad.sum()), float(err[ok].max()) if n else 0.0, n, n / p.size


def oracle_given_symbols(sd, y_in, z_in, M, K):
    """The reference's entropy path and synthesis transform (oracle restatement) evaluated on GIVEN symbols: what p_y, p_z and
    x_hat must be for this run's y_in / z_in (Models.py:69-90)."""
    from oracle import forward as O
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    y_in, z_in = torch.as_tensor(np.asarray(y_in)), torch.as_tensor(np.asarray(z_in))
    with torch.no_grad():
        psi = O.hyper_synthesis(sd, z_in)
        phi = O.context(sd, y_in)
        raw = O.entropy_parameters_raw(sd, torch.cat([phi, psi], dim=1))
        p_y = O.conditional_likelihood(y_in, O.split_parameters(raw, M, K), K)
        p_z = O.factorized_likelihood(sd, z_in)
        x_hat = O.synthesis(sd, y_in)
    return p_y.numpy(), p_z.numpy(), x_hat.numpy()


def oracle_scalable_given_symbols(sd, y_in, z_in, M, M1, K):
    """oracle_given_symbols for ScalableImageCoding (two entropy heads over the shared psi; oracle.forward_scalable's order)."""
    from oracle import forward as O
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    y_in, z_in = torch.as_tensor(np.asarray(y_in)), torch.as_tensor(np.asarray(z_in))
    out = {}
    with torch.no_grad():
        psi = O.hyper_synthesis(sd, z_in)
        for i, (yi, mi) in enumerate(zip(torch.split(y_in, [M1, M - M1], dim=1), (M1, M - M1)), start=1):
            phi = O.context(sd, yi, prefix=f"context_model_{i}")
            raw = O.entropy_parameters_raw(sd, torch.cat([phi, psi], dim=1), prefix=f"entropy_parameters_{i}")
            out[f"p_y{i}"] = O.conditional_likelihood(yi, O.split_parameters(raw, mi, K), K).numpy()
        out["p_z"] = O.factorized_likelihood(sd, z_in).numpy()
        out["x_hat"] = O.synthesis(sd, y_in).numpy()
    return out


def record_report(name, report):
    """Append one line to gpurun_out/parity_report.jsonl (copied to profiles/ per round: the judged evidence of what the parity
    tests compared)."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: (list(v) if isinstance(v, tuple) else v) for k, v in report.items()}}) + "\n")
    except OSError:
        pass


def symbol_mismatches(sym, sym_ref, pre_ref, tau):
    """Mismatching symbols, split into near-tie ones (|frac(pre_ref)| within tau of .5) and real ones."""
    sym, sym_ref, pre_ref = (np.asarray(a) for a in (sym, sym_ref, pre_ref))
    diff = sym != sym_ref
    frac = np.abs(pre_ref - np.floor(pre_ref) - 0.5)
    tie = frac < tau
    return int((diff & ~tie).sum()), int((diff & tie).sum())
