"""Generate tests/golden/*.npz from the REAL reference (TEST INFRASTRUCTURE ONLY).

Runs only in the authoring container, where /root/reference exists.  The
reference classes are imported unmodified with two stand-ins placed in
``sys.modules`` first:
  * ``matplotlib`` / ``matplotlib.pyplot``: empty modules (utils.py:2 imports
    pyplot at module top; nothing on the forward path uses it);
  * ``compressai.layers.gdn.GDN``: oracle/gdn.py (third-party arithmetic that
    is absent from /root/reference and from this image - parity unpinned there).

Usage:  python -m oracle.make_golden [scalable | train | residual | residual-train]  (from the repo root)
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# Three seeded weight sets (all start from the reference's own default init under torch.manual_seed(0)):
#   plain : as drawn.  |y| < 0.25, every symbol is 0 - checks plumbing only.
#   gain  : last g_a conv x 136.2, last h_a conv x 3.86 (std(y) ~ 8, std(z) ~ 3).  Non-trivial symbols,
#           but the random entropy model is badly matched: > 50 % of p_y sit at the 1e-9 clamp, where the reference's
#           erf-difference form is rounding noise (its own fp32 and fp64 runs differ by 2e-2 bpp) - the stress case.
#   calib : last g_a conv x 34 (std(y) ~ 2), last h_a conv x 3.86, + 3 on the sigma biases of the entropy-parameter
#           head (sigma ~ 3).  Non-trivial symbols AND well-conditioned likelihoods, like a trained model - the case the
#           1e-3 bpp / PSNR criterion is checked on.
INITS = {"plain": (1.0, 1.0, 0.0), "gain": (136.2, 3.86, 0.0), "calib": (34.0, 3.86, 3.0)}

CASES = {
    # name: (M, K, input shape, init)
    "c1_k1_256_plain": (128, 1, (1, 3, 256, 256), "plain"),   # BASELINE.json configs[0]
    "c1_k1_128_gain": (128, 1, (1, 3, 128, 128), "gain"),
    "c1_k1_128_calib": (128, 1, (1, 3, 128, 128), "calib"),
    "c2_k3_128x192_plain": (128, 3, (2, 3, 128, 192), "plain"),
    "c2_k3_128x192_gain": (128, 3, (2, 3, 128, 192), "gain"),
    "c2_k3_128x192_calib": (128, 3, (2, 3, 128, 192), "calib"),
}


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from oracle import gdn as _gdn
    pkg = types.ModuleType("compressai"); layers = types.ModuleType("compressai.layers")
    gmod = types.ModuleType("compressai.layers.gdn"); gmod.GDN = _gdn.GDN
    pkg.layers = layers; layers.gdn = gmod
    sys.modules.update({"compressai": pkg, "compressai.layers": layers, "compressai.layers.gdn": gmod})
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import Models, RateDistortionLoss  # noqa: E401  (the reference's own modules)
    return Models, RateDistortionLoss


def apply_init(sd, init: str):
    """Re-scale a default-init state_dict into one of the INITS weight sets (in place; returns sd)."""
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


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode()); h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_reference_model(Models, M, K, init):
    torch.manual_seed(0)
    model = Models.JointAutoregressiveHierarchical(M, K=K)
    if init != "plain":
        model.load_state_dict(apply_init({k: v.clone() for k, v in model.state_dict().items()}, init))
    return model


SCALABLE_CASES = {
    # name: (M, M1, K, input shape, init) - BASELINE.json configs[4]'s model at parity-test size
    "c5_scalable_k1_128_calib": (192, 128, 1, (1, 3, 128, 128), "calib"),
    "c5_scalable_k3_64x128_calib": (192, 128, 3, (2, 3, 64, 128), "calib"),
}


def apply_init_scalable(sd, init: str):
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
    return sd


def reference_scalable_forward(model, x):
    """The reference's OWN sub-modules of ScalableImageCoding called in the order of Models.py:259-338 with the four repairs
    of SURVEY.md section 2.4 (the committed forward raises): no stray `debug` argument, keyword names mu= / sigma=, params2
    bound, LatentSpaceTransform skipped."""
    y = model.encoder(x)
    z = model.hyper_encoder(y)
    z_in, y_in = torch.round(z), torch.round(y)
    y1, y2 = torch.split(y_in, [model.M1, model.M2], dim=1)
    psi = model.hyper_decoder(z_in)
    phi1, phi2 = model.context_model_1(y1), model.context_model_2(y2)
    par1 = model.entropy_parameters_1(torch.cat([phi1, psi], dim=1))
    par2 = model.entropy_parameters_2(torch.cat([phi2, psi], dim=1))
    names = ("mu", "sigma") if model.K == 1 else ("weights", "mus", "sigmas")
    p_z = model.factorized_entropy_model(z_in)
    p_y1 = model.conditional(y1, **dict(zip(names, par1)))
    p_y2 = model.conditional(y2, **dict(zip(names, par2)))
    out = {"x_hat": model.decoder(y_in), "y": y, "y_in": y_in, "y1": y1, "y2": y2, "z": z, "z_in": z_in, "p_z": p_z,
           "logp_z": torch.log(p_z), "p_y1": p_y1, "logp_y1": torch.log(p_y1), "p_y2": p_y2, "logp_y2": torch.log(p_y2),
           "training": False}
    for i, par in ((1, par1), (2, par2)):
        for n, v in zip(names, par):
            out[f"{n}{i}"] = v
    return out


def main_scalable():
    Models, RDL = import_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, (M, M1, K, shape, init) in SCALABLE_CASES.items():
        torch.manual_seed(0)
        model = Models.ScalableImageCoding(M, M1, K=K)
        sd = apply_init_scalable({k: v.clone() for k, v in model.state_dict().items()}, init)
        model.load_state_dict(sd)
        digest = state_digest({k: v for k, v in sd.items() if not k.startswith("LST.")})
        torch.manual_seed(1)
        x = torch.rand(*shape)
        with torch.no_grad():
            out = reference_scalable_forward(model, x)
            rd = RDL.vision_rd_loss(out, x, 0.005, 0.0)           # the reference's own loss, V = None branch
        blob = {"x": x.numpy(), "state_digest": np.array(digest), "M": np.array(M), "M1": np.array(M1), "K": np.array(K),
                "init": np.array(init)}
        for k, v in out.items():
            if torch.is_tensor(v):
                blob["out_" + k] = v.numpy()
        for k, v in rd.items():
            blob["rd_" + k] = v.detach().numpy() if torch.is_tensor(v) else np.array(v, dtype=np.float64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: bpp_y1 {rd['bpp_y1']:.6f} bpp_y2 {rd['bpp_y2']:.6f} bpp_z {rd['bpp_z']:.6f} psnr {rd['psnr']:.6f} "
              f"nonzero y_in {int((out['y_in'] != 0).sum())}/{out['y_in'].numel()} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def main():
    Models, RDL = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    from oracle import forward as O
    for name, (M, K, shape, init) in CASES.items():
        model = build_reference_model(Models, M, K, init)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        digest = state_digest(sd0)                       # before the masked conv zeroes its taps in place
        torch.manual_seed(1)
        x = torch.rand(*shape)
        with torch.no_grad():
            out = model(x, training=False)
            rd = RDL.rd_loss(out, x, 0.005)
            # fp64 shadow of the same arithmetic (oracle restatement): how far the reference's own fp32 rounding
            # moves bpp / PSNR on this weight set = the conditioning band of the case
            o64 = O.forward(sd0, x, M, K, dtype=torch.float64)
            rd64 = O.rd_loss(o64, x, 0.005)
        blob = {"x": x.numpy(), "state_digest": np.array(digest), "M": np.array(M), "K": np.array(K),
                "init": np.array(init), "fp64_bpp_total": np.array(rd64["bpp_total"]),
                "fp64_psnr": np.array(rd64["psnr"]),
                "fp64_symbol_flips": np.array(int((o64["y_in"] != out["y_in"].double()).sum()))}
        for k, v in out.items():
            if torch.is_tensor(v):
                blob["out_" + k] = v.numpy()
        for k, v in rd.items():
            blob["rd_" + k] = v.detach().numpy() if torch.is_tensor(v) else np.array(v, dtype=np.float64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: bpp_y {rd['bpp_y']:.6f} bpp_z {rd['bpp_z']:.6f} (fp64 total {rd64['bpp_total']:.6f}) psnr {rd['psnr']:.6f} "
              f"nonzero y_in {int((out['y_in'] != 0).sum())}/{out['y_in'].numel()} -> {path} "
              f"({os.path.getsize(path) / 1e6:.2f} MB)")


RESIDUAL_CASES = {
    # name: (M, K, input shape) - HierarchicalMixtureResidual (Models.py:109-205), the 3x3 residual family
    "c6_res3x3_k3_128_calib": (128, 3, (1, 3, 128, 192)),
    "c6_res3x3_k1_128_calib": (128, 1, (2, 3, 64, 128)),
}


def residual_init(sd, gy, gz, sigma_bias=3.0):
    """The 'calib' recipe for the residual model: gains on the bottleneck conv of g_a (net.6) and the last conv of h_a (net.8),
    + sigma_bias on the sigma biases of the entropy-parameter head."""
    for k in ("encoder.net.6.weight", "encoder.net.6.bias"):
        sd[k] = sd[k] * gy
    for k in ("hyper_encoder.net.8.weight", "hyper_encoder.net.8.bias"):
        sd[k] = sd[k] * gz
    b = sd["entropy_parameters.net.4.bias"].clone()
    n = b.numel()
    b[(n // 2 if n % 3 else 2 * n // 3):] += sigma_bias
    sd["entropy_parameters.net.4.bias"] = b
    return sd


def main_residual():
    """The reference's own HierarchicalMixtureResidual (unmodified classes) on seeded weights.  The gains are chosen per case so that
    std(y) = 2 and std(z) = 3 on the case's input (non-trivial symbols, well-conditioned likelihoods) and stored in the file."""
    Models, RDL = import_reference()
    os.makedirs(OUT, exist_ok=True)
    from oracle import forward as O
    for name, (M, K, shape) in RESIDUAL_CASES.items():
        torch.manual_seed(0)
        model = Models.HierarchicalMixtureResidual(M, K=K)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        torch.manual_seed(1)
        x = torch.rand(*shape)
        with torch.no_grad():
            o = model(x, training=False)
            gy = float(np.float32(2.0 / float(o["y"].std())))
            sd1 = residual_init({k: v.clone() for k, v in sd0.items()}, gy, 1.0, 0.0)
            model.load_state_dict(sd1)
            gz = float(np.float32(3.0 / float(model(x, training=False)["z"].std())))
        sd = residual_init({k: v.clone() for k, v in sd0.items()}, gy, gz)
        model.load_state_dict(sd)
        digest = state_digest(sd)
        with torch.no_grad():
            out = model(x, training=False)
            rd = RDL.rd_loss(out, x, 0.005)
            rd64 = O.rd_loss(O.forward_residual(sd, x, M, K, dtype=torch.float64), x, 0.005)
        blob = {"x": x.numpy(), "state_digest": np.array(digest), "M": np.array(M), "K": np.array(K), "gain_y": np.array(gy, np.float32),
                "gain_z": np.array(gz, np.float32), "fp64_bpp_total": np.array(rd64["bpp_total"]), "fp64_psnr": np.array(rd64["psnr"])}
        for k, v in out.items():
            if torch.is_tensor(v):
                blob["out_" + k] = v.numpy()
        for k, v in rd.items():
            blob["rd_" + k] = v.detach().numpy() if torch.is_tensor(v) else np.array(v, dtype=np.float64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: gains {gy:.4f} / {gz:.4f}, bpp_y {rd['bpp_y']:.6f} bpp_z {rd['bpp_z']:.6f} (fp64 total {rd64['bpp_total']:.6f}) psnr {rd['psnr']:.6f} "
              f"nonzero y_in {int((out['y_in'] != 0).sum())}/{out['y_in'].numel()} min p_y {float(out['p_y'].min()):.2e} -> {path} "
              f"({os.path.getsize(path) / 1e6:.2f} MB)")


TRAIN_CASES = {
    # name: (M, K, input shape, init, noise seed) - BASELINE.json configs[3] (training step) at parity-test size
    "c4_train_k3_128_calib": (128, 3, (2, 3, 128, 128), "calib", 11),
    "c4_train_k1_64x128_calib": (128, 1, (2, 3, 64, 128), "calib", 12),
}
TRAIN_SAMPLES = 48        # entries kept per gradient tensor


def sample_index(numel: int, n: int = TRAIN_SAMPLES):
    return np.unique(np.linspace(0, numel - 1, num=min(n, numel)).astype(np.int64))


def main_train():
    """The reference's training step (Trainer.py:81-86): model(imgs) with its own torch.rand_like noise, rd_loss,
    loss.backward(), torch.optim.Adam(lr=1e-4).step().  Gradients are kept as per-tensor L2 norms + sampled entries."""
    Models, RDL = import_reference()
    os.makedirs(OUT, exist_ok=True)
    from oracle import backward as OB
    for name, (M, K, shape, init, nseed) in TRAIN_CASES.items():
        model = build_reference_model(Models, M, K, init)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        digest = state_digest(sd0)
        torch.manual_seed(1)
        x = torch.rand(*shape)
        B, _, H, W = shape
        # the reference draws z's noise first, then y's (Models.py:57-58), from the global generator; the seed is the first one
        # from `nseed` on that keeps every LeakyReLU pre-activation >= 1e-6 away from its kink (oracle/backward.py: kink_margin)
        nseed, noise_z, noise_y = OB.noise_with_margin(sd0, x, M, K, nseed)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        opt.zero_grad()
        torch.manual_seed(nseed)
        out = model(x)                                   # training=True is the default (Trainer.py:82)
        assert torch.equal(out["z_in"] - out["z"], (out["z"] + noise_z) - out["z"]), "noise stream not reproduced"
        rd = RDL.rd_loss(out, x, 0.005)
        rd["loss"].backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        opt.step()
        after = {k: p.detach().clone() for k, p in model.named_parameters()}
        # fp64 shadow of the same step through the oracle: the conditioning band of each gradient
        _, g64, _ = OB.loss_and_grads(sd0, x, M, K, noise_z, noise_y, 0.005, dtype=torch.float64)
        blob = {"x": x.numpy(), "noise_z": noise_z.numpy(), "noise_y": noise_y.numpy(), "state_digest": np.array(digest),
                "M": np.array(M), "K": np.array(K), "init": np.array(init), "loss": np.array(float(rd["loss"].detach())),
                "bpp_total": np.array(rd["bpp_total"]), "mse": np.array(rd["mse"]), "noise_seed": np.array(nseed),
                "kink_margin": np.array(OB.kink_margin(sd0, x, M, K, noise_z, noise_y))}
        for k, g in grads.items():
            idx = sample_index(g.numel())
            blob["gnorm_" + k] = np.array(float(g.double().norm()))
            blob["gnorm64_" + k] = np.array(float(g64[k].norm()))
            blob["gerr64_" + k] = np.array(float((g.double() - g64[k]).norm()))
            blob["gsamp_" + k] = g.reshape(-1)[idx].numpy()
            blob["psamp_" + k] = after[k].reshape(-1)[idx].numpy()
            blob["dnorm_" + k] = np.array(float((after[k].double() - sd0[k].double()).norm()))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        worst = max(float(blob["gerr64_" + k] / max(blob["gnorm64_" + k], 1e-30)) for k in grads)
        print(f"{name}: noise seed {nseed}, LeakyReLU margin {float(blob['kink_margin']):.2e}, loss {float(rd['loss'].detach()):.6f} bpp {rd['bpp_total']:.6f} mse {rd['mse']:.6f}; {len(grads)} gradients, "
              f"worst fp32-vs-fp64 relative gradient error {worst:.2e} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


RESIDUAL_TRAIN_CASES = {
    # name: (M, K, input shape, noise seed) - the training step of HierarchicalMixtureResidual at parity-test size
    "c6_train_res3x3_k3_128_calib": (128, 3, (2, 3, 64, 128), 21),
    "c6_train_res3x3_k1_128x64_calib": (128, 1, (1, 3, 128, 64), 22),
}


def main_residual_train():
    """The reference's training step (Trainer.py:81-86) on its own HierarchicalMixtureResidual: model(imgs) with its torch.rand_like
    noise, rd_loss, loss.backward(), torch.optim.Adam(lr=1e-4).step().  Weights: the calib recipe of main_residual (gains from the
    evaluation forward of the same input).  Gradients are kept as per-tensor L2 norms + sampled entries, as in main_train."""
    Models, RDL = import_reference()
    os.makedirs(OUT, exist_ok=True)
    from oracle import backward as OB
    for name, (M, K, shape, nseed) in RESIDUAL_TRAIN_CASES.items():
        torch.manual_seed(0)
        model = Models.HierarchicalMixtureResidual(M, K=K)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        torch.manual_seed(1)
        x = torch.rand(*shape)
        with torch.no_grad():
            gy = float(np.float32(2.0 / float(model(x, training=False)["y"].std())))
            model.load_state_dict(residual_init({k: v.clone() for k, v in sd0.items()}, gy, 1.0, 0.0))
            gz = float(np.float32(3.0 / float(model(x, training=False)["z"].std())))
        sd = residual_init({k: v.clone() for k, v in sd0.items()}, gy, gz)
        model.load_state_dict(sd)
        digest = state_digest(sd)
        B, _, H, W = shape
        torch.manual_seed(nseed)                         # the reference draws z's noise first, then y's (Models.py:158-159)
        noise_z = torch.rand(B, M, H // 64, W // 64) - 0.5
        noise_y = torch.rand(B, M, H // 16, W // 16) - 0.5
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        opt.zero_grad()
        torch.manual_seed(nseed)
        out = model(x)                                   # training=True is the default
        assert torch.equal(out["z_in"] - out["z"], (out["z"] + noise_z) - out["z"]), "noise stream not reproduced"
        assert torch.equal(out["y_in"] - out["y"], (out["y"] + noise_y) - out["y"]), "noise stream not reproduced"
        rd = RDL.rd_loss(out, x, 0.005)
        rd["loss"].backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        opt.step()
        after = {k: p.detach().clone() for k, p in model.named_parameters()}
        _, g64, _ = OB.loss_and_grads_residual(sd, x, M, K, noise_z, noise_y, 0.005, dtype=torch.float64)
        blob = {"x": x.numpy(), "noise_z": noise_z.numpy(), "noise_y": noise_y.numpy(), "state_digest": np.array(digest),
                "M": np.array(M), "K": np.array(K), "gain_y": np.array(gy, np.float32), "gain_z": np.array(gz, np.float32),
                "loss": np.array(float(rd["loss"].detach())), "bpp_total": np.array(rd["bpp_total"]), "mse": np.array(rd["mse"]),
                "noise_seed": np.array(nseed)}
        for k, g in grads.items():
            idx = sample_index(g.numel())
            blob["gnorm_" + k] = np.array(float(g.double().norm()))
            blob["gnorm64_" + k] = np.array(float(g64[k].norm()))
            blob["gerr64_" + k] = np.array(float((g.double() - g64[k]).norm()))
            blob["gsamp_" + k] = g.reshape(-1)[idx].numpy()
            blob["psamp_" + k] = after[k].reshape(-1)[idx].numpy()
            blob["dnorm_" + k] = np.array(float((after[k].double() - sd[k].double()).norm()))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        worst = max(float(blob["gerr64_" + k] / max(blob["gnorm64_" + k], 1e-30)) for k in grads)
        print(f"{name}: gains {gy:.4f} / {gz:.4f}, loss {float(rd['loss'].detach()):.6f} bpp {rd['bpp_total']:.6f} mse {rd['mse']:.6f}; {len(grads)} gradients, "
              f"worst fp32-vs-fp64 relative gradient error {worst:.2e} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        main_train()               # only the c4_train_* files
    elif len(sys.argv) > 1 and sys.argv[1] == "residual-train":
        main_residual_train()      # only the c6_train_* file
    elif len(sys.argv) > 1 and sys.argv[1] == "residual":
        main_residual()            # only the c6_* files
    elif len(sys.argv) > 1 and sys.argv[1] == "scalable":
        main_scalable()            # only the c5_* files; the other vectors stay byte-identical
    else:
        main()
