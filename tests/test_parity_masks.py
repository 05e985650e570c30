"""CPU: the masked end-to-end comparison of tests/test_gpu_model.py::check_against, exercised with the oracle itself.

"Ours" is the oracle with y perturbed by ~5e-5 before rounding, so that a few symbols flip at rounding ties and everything
downstream is consistently computed from the flipped symbols - the situation of an fp32-grade GPU arm.  The check must pass
(flips recognised as ties, likelihoods / x_hat compared off the flips' footprints and, given the symbols, everywhere), and it
must FAIL when a likelihood away from any flip is wrong."""
import numpy as np
import pytest
import torch

from oracle import forward as O
from tests import helpers as H
from tests.test_gpu_model import check_against


def _perturbed_run(sd, x, M, K, eps):
    with torch.no_grad():
        y = O.analysis(sd, x)
        z = O.hyper_analysis(sd, y)
        torch.manual_seed(5)
        y = y + eps * torch.randn_like(y)                      # "our" y: the reference's + an fp32-grade error
        y_in, z_in = torch.round(y), torch.round(z)
        psi, phi = O.hyper_synthesis(sd, z_in), O.context(sd, y_in)
        params = O.split_parameters(O.entropy_parameters_raw(sd, torch.cat([phi, psi], 1)), M, K)
        p_z, p_y = O.factorized_likelihood(sd, z_in), O.conditional_likelihood(y_in, params, K)
        out = {"x_hat": O.synthesis(sd, y_in), "y": y, "y_in": y_in, "z": z, "z_in": z_in, "p_z": p_z, "logp_z": torch.log(p_z),
               "p_y": p_y, "logp_y": torch.log(p_y), "training": False}
        out["weights"], out["mus"], out["sigmas"] = params
    return out


def test_masked_check_accepts_tie_flips_and_rejects_real_errors():
    from neural_image_compression_b200.Models import JointAutoregressiveHierarchical
    torch.manual_seed(0)
    model = JointAutoregressiveHierarchical(128, K=3, precision="fp32")
    sd = H.apply_init({k: v.clone() for k, v in model.state_dict().items()}, "calib")
    x = H.seeded_input((1, 3, 256, 384))
    with torch.no_grad():
        ref_t = O.forward(sd, x, 128, 3)
    ref_rd = O.rd_loss(ref_t, x, 0.005)
    ref = {k: v.numpy() for k, v in ref_t.items() if torch.is_tensor(v) and not k.startswith("_")}
    out = _perturbed_run(sd, x, 128, 3, 5e-5)
    flips = int((out["y_in"] != ref_t["y_in"]).sum())
    assert flips > 0, "the perturbation was meant to flip a few symbols"
    rd = O.rd_loss(out, x, 0.005)
    rep = check_against(out, rd, ref, ref_rd, 3, precision="fp32", sd=sd, M=128, min_frac=0.5)
    assert rep["y_in_flips_real_ties"][1] == flips and rep["p_y"]["fraction"] < 1.0 and "p_y_given_symbols" in rep
    # a wrong likelihood far from every flip must be caught
    ok_y, _, _ = H.flip_masks(out["y_in"].numpy(), ref["y_in"], out["z_in"].numpy(), ref["z_in"], ref["x_hat"].shape)
    b, _, i, j = np.argwhere(ok_y)[0]
    bad = dict(out)
    bad["p_y"] = out["p_y"].clone()
    bad["p_y"][b, :, i, j] *= 1.01
    with pytest.raises(AssertionError):
        check_against(bad, rd, ref, ref_rd, 3, precision="fp32", sd=sd, M=128, min_frac=0.5)
    # and a flip that is NOT at a rounding tie is a real mismatch
    bad = dict(out)
    bad["y_in"] = out["y_in"].clone()
    k = int(torch.argmin((ref_t["y"] - torch.round(ref_t["y"])).abs().flatten()))      # a value sitting on an integer
    bad["y_in"].view(-1)[k] += 1
    with pytest.raises(AssertionError):
        check_against(bad, rd, ref, ref_rd, 3, precision="fp32", sd=sd, M=128, min_frac=0.5)
