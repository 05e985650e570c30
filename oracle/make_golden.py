"""Generate tests/golden/*.npz from the REAL reference (TEST INFRASTRUCTURE ONLY).

Runs only in the authoring container, where /root/reference exists.  The
reference classes are imported unmodified with two stand-ins placed in
``sys.modules`` first:
  * ``matplotlib`` / ``matplotlib.pyplot``: empty modules (utils.py:2 imports
    pyplot at module top; nothing on the forward path uses it);
  * ``compressai.layers.gdn.GDN``: oracle/gdn.py (third-party arithmetic that
    is absent from /root/reference and from this image - parity unpinned there).

Usage:  python -m oracle.make_golden          (from the repo root)
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

# frozen gain-init constants: std(y) ~ 8, std(z) ~ 3 on the config-2 input
GAIN_Y = 136.2
GAIN_Z = 3.86

CASES = {
    # name: (M, K, input shape, gain-init?)
    "c1_k1_256_plain": (128, 1, (1, 3, 256, 256), False),   # BASELINE.json configs[0]
    "c1_k1_128_gain": (128, 1, (1, 3, 128, 128), True),
    "c2_k3_128x192_plain": (128, 3, (2, 3, 128, 192), False),
    "c2_k3_128x192_gain": (128, 3, (2, 3, 128, 192), True),
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


def apply_gain(sd):
    """gain-init: scale the last g_a and h_a convs so the symbols are non-trivial."""
    for k in ("encoder.net.6.weight", "encoder.net.6.bias"):
        sd[k] = sd[k] * GAIN_Y
    for k in ("hyper_encoder.net.4.weight", "hyper_encoder.net.4.bias"):
        sd[k] = sd[k] * GAIN_Z
    return sd


def state_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode()); h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_reference_model(Models, M, K, gain):
    torch.manual_seed(0)
    model = Models.JointAutoregressiveHierarchical(M, K=K)
    if gain:
        model.load_state_dict(apply_gain({k: v.clone() for k, v in model.state_dict().items()}))
    return model


def main():
    Models, RDL = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, (M, K, shape, gain) in CASES.items():
        model = build_reference_model(Models, M, K, gain)
        digest = state_digest(model.state_dict())        # before the masked conv zeroes its taps in place
        torch.manual_seed(1)
        x = torch.rand(*shape)
        with torch.no_grad():
            out = model(x, training=False)
            rd = RDL.rd_loss(out, x, 0.005)
        blob = {"x": x.numpy(), "state_digest": np.array(digest), "M": np.array(M), "K": np.array(K),
                "gain": np.array(gain)}
        for k, v in out.items():
            if torch.is_tensor(v):
                blob["out_" + k] = v.numpy()
        for k, v in rd.items():
            blob["rd_" + k] = v.detach().numpy() if torch.is_tensor(v) else np.array(v, dtype=np.float64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: bpp_y {rd['bpp_y']:.6f} bpp_z {rd['bpp_z']:.6f} psnr {rd['psnr']:.6f} "
              f"nonzero y_in {int((out['y_in'] != 0).sum())}/{out['y_in'].numel()} -> {path} "
              f"({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
