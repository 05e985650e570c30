"""GPU: the reference's DEFAULT channel count (latent_channels = 192, Models.py:17) on the tensor-core arm.

M = 128 runs the fused pair-tensor pipeline; other multiples of 64 run layer by layer (every conv and both GDN contractions on
the tcgen05 engine with hi/lo-split operands, fp32 NHWC tensors in between).  Same parity bar as the other arms."""
import numpy as np
import pytest
import torch

from oracle import backward as OB
from oracle import forward as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [1, 3])
def test_m192_eval_forward_matches_the_oracle_on_the_tensor_core_arm(K):
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = H.seeded_model(192, K, "calib", precision=None)
    assert model.precision == "bf16x3"                      # the default arm for channel counts that are multiples of 64
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = H.seeded_input((2, 3, 128, 192))
    ref = O.forward(sd, x, 192, K)
    ref_rd = O.rd_loss(ref, x, 0.005)
    # the weight set is tuned for M = 128: at M = 192 about 1.5 % of p_y sit at the 1e-9 clamp, where the reference's erf difference
    # is rounding noise - its own fp32 and fp64 runs differ by ~5e-3 bpp.  As for the 'gain' cases: the total must land in that
    # band, and the bits over the well-conditioned elements (p >= 1e-6) within 1e-3.
    band = abs(ref_rd["bpp_total"] - O.rd_loss(O.forward(sd, x, 192, K, dtype=torch.float64), x, 0.005)["bpp_total"])
    model = model.cuda()
    out = model(x.cuda(), training=False)
    rd = rd_loss(out, x.cuda(), 0.005)
    for name, pre in (("y_in", "y"), ("z_in", "z")):
        real, ties = H.symbol_mismatches(out[name].cpu().numpy(), ref[name].numpy(), ref[pre].numpy(), 2e-3)
        assert real == 0, (name, real, ties)
    same = (out["y_in"].cpu() == ref["y_in"]).all() and (out["z_in"].cpu() == ref["z_in"]).all()
    if same:
        bad, worst = H.likelihood_close(out["p_y"].cpu().numpy(), ref["p_y"].numpy())
        assert bad <= 2e-3 * ref["p_y"].numel(), (bad, worst)
    assert abs(rd["bpp_total"] - ref_rd["bpp_total"]) <= H.BPP_TOL + band and abs(rd["psnr"] - ref_rd["psnr"]) <= H.PSNR_TOL, (rd, ref_rd, band)
    if same:
        good = ref["p_y"].numpy() >= 1e-6
        npix = x.shape[0] * x.shape[2] * x.shape[3]
        bits = -(np.log2(out["p_y"].cpu().numpy().astype(np.float64)) * good).sum() / npix
        ref_bits = -(np.log2(ref["p_y"].numpy().astype(np.float64)) * good).sum() / npix
        assert abs(bits - ref_bits) <= H.BPP_TOL, (bits, ref_bits)
    assert set(out) == {k for k in ref if not k.startswith("_")}


def test_m192_training_step_gradients():
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = H.seeded_model(192, 1, "calib192", precision=None)          # a weight set that is well conditioned at 192 channels
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = H.seeded_input((1, 3, 128, 128))
    _, nz, ny = OB.noise_with_margin(sd, x, 192, 1, 31)
    ref_rd, ref_g, _ = OB.loss_and_grads(sd, x, 192, 1, nz, ny, 0.005)
    model = model.cuda()
    rd = rd_loss(model(x.cuda(), noise=(nz.cuda(), ny.cuda())), x.cuda(), 0.005)
    assert abs(float(rd["loss"].detach()) - ref_rd["loss"]) <= 5e-5 * abs(ref_rd["loss"])
    rd["loss"].backward()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        exposed = k.startswith(("hyper_encoder.", "hyper_decoder.", "context_model.", "entropy_parameters.net.0", "entropy_parameters.net.2"))
        err = float((p.grad.double().cpu() - ref_g[k].double()).norm() / ref_g[k].double().norm())
        assert err < (3e-2 if exposed else 1e-3), (k, err)        # see tests/test_gpu_train.py: grad_tol (fp32-vs-fp64 spread here: 5e-5)
