"""GPU: the 3x3 residual family (/root/reference/Layers.py, Components.py:20-122,
Models.py:109-205) through the module API -> C ABI: every new layer shape against a float64 CPU convolution, every block
against the oracle, and HierarchicalMixtureResidual against vectors produced by the reference's own class."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import forward as O
from tests import helpers as H
from tests.test_gpu_model import check_against

pytestmark = pytest.mark.gpu
ARMS = ("fp32", "bf16x3")


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("kind,hw", [("c3s2", (16, 24)), ("c3s2", (10, 14)), ("c1s2", (16, 24)), ("t3s2", (8, 12)), ("t3s2", (5, 7)),
                                     ("t3s2_rgb", (8, 12)), ("c3s1_288", (8, 12)), ("t3s2_288", (4, 6))])
def test_new_layer_shapes(kind, hw, arm):
    """Conv2d 3x3 stride 2 / 1x1 stride 2, ConvTranspose2d 3x3 stride 2 (pad 1, output_padding 1), incl. the 3-channel output
    and the 288-channel h_s layers of M = 192 (not a multiple of 64: CUDA-core kernels in both arms)."""
    from neural_image_compression_b200 import Layers as L
    from neural_image_compression_b200._lib import EPI_BIAS
    torch.manual_seed(60)
    conv = {"c3s2": lambda: nn.Conv2d(128, 128, 3, 2, 1), "c1s2": lambda: nn.Conv2d(128, 128, 1, 2),
            "t3s2": lambda: nn.ConvTranspose2d(128, 128, 3, 2, 1, output_padding=1),
            "t3s2_rgb": lambda: nn.ConvTranspose2d(128, 3, 3, 2, 1, output_padding=1),
            "c3s1_288": lambda: nn.Conv2d(192, 288, 3, 1, 1), "t3s2_288": lambda: nn.ConvTranspose2d(288, 288, 3, 2, 1, output_padding=1)}[kind]()
    x = torch.randn(2, conv.in_channels, *hw)
    if isinstance(conv, nn.ConvTranspose2d):
        ref = F.conv_transpose2d(x.double(), conv.weight.double(), conv.bias.double(), stride=2, padding=1, output_padding=1)
    else:
        ref = F.conv2d(x.double(), conv.weight.double(), conv.bias.double(), stride=conv.stride, padding=conv.padding)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    y, ho, wo = L._conv(arm, conv.cuda(), EPI_BIAS, xn, 2, hw[0], hw[1])
    torch.cuda.synchronize()
    got = y.cpu().permute(0, 3, 1, 2)
    assert got.shape == ref.shape and _rel(got, ref) < (2e-5 if arm == "fp32" else 1e-4), (kind, _rel(got, ref))


@pytest.mark.parametrize("arm", ARMS)
def test_blocks_match_oracle(arm):
    from neural_image_compression_b200 import Layers as L
    torch.manual_seed(61)
    for blk, fn, shape in ((L.ResidualBlock(128, 128), O._res_block, (2, 128, 8, 12)),
                           (L.ResidualBlockWithStride(128, 128, 2), O._res_stride, (2, 128, 16, 24)),
                           (L.ResidualBlockWithStride(3, 128, 2), O._res_stride, (2, 3, 32, 32)),
                           (L.ResidualBlockUpsample(128, 128, 2), O._res_up, (2, 128, 8, 12))):
        blk.precision = arm
        with torch.no_grad():
            for g in ("gdn", "igdn"):
                if hasattr(blk, g):
                    getattr(blk, g).gamma.add_(0.02 * torch.rand_like(getattr(blk, g).gamma))
        x = torch.randn(*shape) if shape[1] > 3 else torch.rand(*shape)
        sd = {"b." + k: v.detach().clone().double() for k, v in blk.state_dict().items()}
        with torch.no_grad():
            ref = fn(sd, "b", x.double(), torch.float64)
        got = blk.cuda()(x.cuda()).cpu()
        assert got.shape == ref.shape and _rel(got, ref) < (3e-5 if arm == "fp32" else 1.5e-4), (type(blk).__name__, _rel(got, ref))


def test_subpel_conv_matches_torch():
    """SubpelConv3x3 (Layers.py:6-16) is public in the reference although its models build TransposedDeconv3x3 instead."""
    from neural_image_compression_b200 import Layers as L
    torch.manual_seed(62)
    m = L.SubpelConv3x3(128, 64, 2)
    x = torch.randn(2, 128, 8, 12)
    with torch.no_grad():
        ref = F.pixel_shuffle(F.conv2d(x.double(), m.conv.weight.double(), m.conv.bias.double(), padding=1), 2)
    m.precision = "fp32"
    got = m.cuda()(x.cuda()).cpu()
    assert got.shape == ref.shape and _rel(got, ref) < 2e-5


@pytest.mark.parametrize("precision", ARMS)
@pytest.mark.parametrize("case", H.residual_cases())
def test_residual_model_matches_reference_vectors(case, precision):
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    g = H.load_golden(case)
    M, K = int(g["M"]), int(g["K"])
    model = H.seeded_residual_model(M, K, float(g["gain_y"]), float(g["gain_z"]), precision=precision)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    x = torch.from_numpy(g["x"]).cuda()
    out = model(x, training=False)
    rd = rd_loss(out, x, 0.005)
    ref = {k[4:]: g[k] for k in g.files if k.startswith("out_")}
    ref_rd = {k[3:]: float(g[k]) for k in g.files if k.startswith("rd_") and g[k].ndim == 0}
    o = {k: out[k].cpu().numpy() for k in ("y", "y_in", "z", "z_in", "p_y", "p_z", "x_hat")}
    rep = {"precision": precision}
    # symbols: bit exact away from rounding ties; the tie flips are bounded by what the measured error on y / z explains.  This
    # family is 16 convs + 3 GDN deep before the rounding (the 5x5 model: 4 + 3), so the pre-rounding bound is 3e-4 here.
    for name, pre in (("y_in", "y"), ("z_in", "z")):
        real, ties = H.symbol_mismatches(o[name], ref[name], ref[pre], H.TIE_TAU)
        rel = float(np.abs(o[pre] - ref[pre]).max() / np.abs(ref[pre]).max())
        rep[name + "_flips_real_ties"], rep[pre + "_rel_err"] = [real, ties], rel
        assert real == 0 and ties <= H.tie_flip_bound(o[pre], ref[pre]) and rel <= 3e-4, (name, real, ties, rel)
    ok_y, ok_z, ok_x = H.flip_masks(o["y_in"], ref["y_in"], o["z_in"], ref["z_in"], ref["x_hat"].shape)
    for name, ok in (("p_y", ok_y), ("p_z", ok_z)):
        bad, worst, n, frac = H.masked_likelihood_close(o[name], ref[name], ok)
        rep[name] = {"outliers": bad, "max_abs_err": worst, "compared": n, "fraction": frac}
        assert bad <= 1e-5 * n + 1, (name, bad, n, worst)
    okx = np.broadcast_to(ok_x, ref["x_hat"].shape)
    if okx.any():
        xe = float(np.abs(o["x_hat"] - ref["x_hat"])[okx].max() / np.abs(ref["x_hat"]).max())
        rep["x_hat"] = {"rel_err": xe, "fraction": float(okx.mean())}
        assert xe < 5e-4, xe
    if not (ok_y.all() and ok_z.all()):                       # flips: everything else against the oracle on this run's symbols
        with torch.no_grad():
            yi, zi = torch.from_numpy(o["y_in"]), torch.from_numpy(o["z_in"])
            raw = O.entropy_parameters_raw(sd, torch.cat([O.context(sd, yi), O.hyper_synthesis_3x3(sd, zi)], dim=1))
            p_y = O.conditional_likelihood(yi, O.split_parameters(raw, M, K), K).numpy()
            x_hat = O.synthesis_3x3(sd, yi).numpy()
        bad, worst = H.likelihood_close(o["p_y"], p_y)
        rep["p_y_given_symbols"] = {"outliers": bad, "max_abs_err": worst}
        assert bad <= 1e-5 * p_y.size + 1, (bad, worst)
        xe2 = float(np.abs(o["x_hat"] - x_hat).max() / np.abs(x_hat).max())
        rep["x_hat_given_symbols_rel_err"] = xe2
        assert xe2 < 5e-4, xe2
    assert abs(rd["bpp_total"] - ref_rd["bpp_total"]) <= H.BPP_TOL and abs(rd["psnr"] - ref_rd["psnr"]) <= H.PSNR_TOL, (rd, ref_rd)
    H.record_report(f"residual/{case}/{precision}", rep)
    assert set(out) == set(ref) | {"training"}


def test_residual_model_bad_arguments_and_training_calls():
    from neural_image_compression_b200.Models import HierarchicalMixtureResidual
    with pytest.raises(ValueError):
        HierarchicalMixtureResidual(0)
    with pytest.raises(ValueError):
        HierarchicalMixtureResidual(128, K=0)
    m = HierarchicalMixtureResidual(128, K=1).cuda()
    x = torch.rand(1, 3, 64, 64).cuda()
    out = m(x)                                                # training=True with autograd on: the differentiable training forward
    assert out["training"] is True and out["x_hat"].requires_grad and out["logp_y"].requires_grad and not out["y_in"].requires_grad
    assert float((out["y_in"] - out["y"]).abs().max()) <= 0.5
    with torch.no_grad():
        out = m(x, training=True)
    assert out["training"] is True and float((out["y_in"] - out["y"]).abs().max()) <= 0.5 and not out["x_hat"].requires_grad


def test_graphed_forward_of_the_main_and_residual_models():
    """parallel.GraphedForward on JointAutoregressiveHierarchical (captured with its parallel branches: g_s and the context conv
    beside h_a -> h_s) and on HierarchicalMixtureResidual: bit-identical to the eager calls."""
    from neural_image_compression_b200 import parallel
    x = H.seeded_input((2, 3, 128, 192)).cuda()
    for model in (H.seeded_model(128, 3, "calib", precision="bf16x3").cuda(),
                  H.seeded_residual_model(128, 1, 17.98, 141.28, precision="bf16x3").cuda()):
        eager = model(x, training=False)
        gf = parallel.GraphedForward(model)
        for _ in range(3):
            out = gf(x)
        torch.cuda.synchronize()
        for k in ("x_hat", "y_in", "z_in", "p_y", "p_z", "logp_y"):
            assert torch.equal(out[k], eager[k]), (type(model).__name__, k)


# ---------------------------------------------------------------------------------------------------
# training step of the residual family (training.Tape): forward with noise, rd_loss, backward, Adam
# ---------------------------------------------------------------------------------------------------
def _relnorm(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


def _res_grad_tol(arm, key):
    """The reference's own fp32-vs-fp64 gradient spread on this case is 1.5e-4 of a tensor's norm (the model is 16 convs deep with a
    LeakyReLU in every residual block: an element within rounding of a kink takes the other slope).  Measured against the oracle's
    autograd: fp32 arm <= 4.1e-6 (allowed 2e-4), bf16x3 arm <= 2.7e-4 (allowed 5e-3: a kink flip moves the tensors upstream of it)."""
    return 2e-4 if arm == "fp32" else 5e-3


@pytest.mark.parametrize("arm", ["fp32", "bf16x3"])
@pytest.mark.parametrize("case", H.residual_train_cases())
def test_residual_training_step_matches_reference_golden(case, arm):
    """model(x) with autograd on -> rd_loss -> loss.backward() -> Adam.step() of HierarchicalMixtureResidual against the committed
    vectors of the REAL reference class's step (sampled entries + norms of all 119 gradient tensors) and, tensor by tensor in full,
    against the oracle's autograd on the same weights, input and noise."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    from neural_image_compression_b200.training import Adam
    from oracle import backward as OB
    g = H.load_golden(case)
    M, K = int(g["M"]), int(g["K"])
    model = H.seeded_residual_model(M, K, float(g["gain_y"]), float(g["gain_z"]), precision="fp32")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x, nz, ny = (torch.from_numpy(g[k]) for k in ("x", "noise_z", "noise_y"))
    ref_rd, ref_g, ref_out = OB.loss_and_grads_residual(sd, x, M, K, nz, ny, 0.005)
    model = model.cuda()
    model.train_precision = arm
    opt = Adam(model.parameters(), lr=1e-4)
    opt.zero_grad()
    out = model(x.cuda(), noise=(nz.cuda(), ny.cuda()))              # training=True is the default, as in Trainer.py:82
    rd = rd_loss(out, x.cuda(), 0.005)
    ltol = 1e-5 if arm == "fp32" else 1e-4
    assert abs(float(rd["loss"].detach()) - float(g["loss"])) <= ltol * abs(float(g["loss"]))
    assert _relnorm(out["x_hat"].detach().cpu(), ref_out["x_hat"]) < (1e-5 if arm == "fp32" else 3e-4)
    assert _relnorm(out["y"].detach().cpu(), ref_out["y"]) < (1e-5 if arm == "fp32" else 3e-4)
    rd["loss"].backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    torch.cuda.synchronize()
    assert len(grads) == 119 and set(grads) == set(ref_g)
    worst = {k: float((grads[k].double().cpu() - ref_g[k].double()).norm() / max(float(ref_g[k].double().norm()), 1e-30)) for k in grads}
    print(arm, "worst gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    bad = {k: v for k, v in worst.items() if v > _res_grad_tol(arm, k)}
    assert not bad, bad
    want, _ = OB.adam_step({k: v.cpu() for k, v in before.items()}, {k: v.cpu() for k, v in grads.items()})
    for k, p in model.named_parameters():
        gr = grads[k]
        idx = H.sample_index(gr.numel())
        gn = float(g["gnorm_" + k])
        tol = _res_grad_tol(arm, k)
        assert abs(float(gr.double().norm()) - gn) <= tol * gn + 1e-12, (k, float(gr.double().norm()), gn)
        np.testing.assert_allclose(gr.reshape(-1)[idx].cpu().numpy(), g["gsamp_" + k], rtol=5 * tol, atol=10 * tol * gn / gr.numel() ** 0.5,
                                   err_msg=k)
        np.testing.assert_allclose(p.detach().cpu().numpy(), want[k].numpy(), rtol=1e-6, atol=1e-7, err_msg=k)      # the Adam kernel


def test_residual_step_gradients_equals_autograd_backward():
    """training.step_gradients (what parallel.ShardedTrainer runs and captures) on the residual model = the autograd route."""
    from neural_image_compression_b200 import training as T
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    g = H.load_golden(H.residual_train_cases()[0])
    model = H.seeded_residual_model(int(g["M"]), int(g["K"]), float(g["gain_y"]), float(g["gain_z"]), precision="fp32").cuda()
    x, nz, ny = (torch.from_numpy(g[k]).cuda() for k in ("x", "noise_z", "noise_y"))
    out = model(x, noise=(nz, ny))
    rd_loss(out, x, 0.005)["loss"].backward()
    ref = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = None
    res = T.step_gradients(model, x, 0.005, noise=(nz, ny))
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.allclose(p.grad, ref[k], rtol=1e-5, atol=1e-8 + 1e-6 * float(ref[k].abs().max())), k
    assert res is not None


def test_residual_graphed_trainer_steps_like_the_eager_one():
    """parallel.ShardedTrainer on HierarchicalMixtureResidual: eager and CUDA-graph replays of forward + loss + backward + Adam
    (fresh noise per step: statistical check, as tests/test_gpu_train.py does for the 5x5 model)."""
    from neural_image_compression_b200 import parallel
    g = H.load_golden(H.residual_train_cases()[0])
    x = torch.from_numpy(g["x"]).cuda()
    losses = {}
    for mode in (False, True):
        model = H.seeded_residual_model(int(g["M"]), int(g["K"]), float(g["gain_y"]), float(g["gain_z"]), precision="bf16x3").cuda()
        w0 = model.decoder.net[6].conv1.weight.detach().clone()
        tr = parallel.ShardedTrainer(model, 0.005, lr=1e-4, graph=mode)
        ls = [float(tr.step(x)["loss"].detach()) for _ in range(20)]
        torch.cuda.synchronize()
        losses[mode] = ls
        assert tr.optimizer.t == 20 and int(tr.optimizer._t_dev) == 20
        moved = float((model.decoder.net[6].conv1.weight.detach() - w0).abs().max())
        assert 3e-4 < moved <= 20 * 3.2e-4, moved
    assert abs(losses[True][0] - losses[False][0]) < 0.02 * losses[False][0], (losses[True][0], losses[False][0])
    assert losses[True][-1] < losses[True][0] and abs(losses[True][-1] - losses[False][-1]) < 0.05 * losses[False][-1], losses
