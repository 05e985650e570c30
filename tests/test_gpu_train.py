"""GPU: the training step (BASELINE configs[3]; reference Trainer.py:79-86) against the oracle's autograd.

model(x) with autograd on -> rd_loss -> loss.backward() -> Adam step, all through the C-ABI backward kernels, compared with
oracle/backward.py (torch autograd over the restated forward + compressai's LowerBound rule + restated Adam), which is itself
pinned to the REAL reference's gradients by tests/test_oracle.py (tests/golden/c4_train_*.npz).
Cases keep every LeakyReLU pre-activation >= 1e-6 away from 0 (the derivative jumps there; see oracle/backward.py).
Tolerances: every gradient tensor within 2e-4 of its own norm (the reference's fp32-vs-fp64 gradient spread on these cases is
4e-6 relative; ours adds the accumulation order of split-K sums), loss within 1e-5 relative, Adam-updated parameters within 2e-6.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import backward as OB
from oracle import forward as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

GRAD_RTOL = 2e-4
ARMS = ("fp32", "bf16x3")       # arithmetic of the step's convolutions (training.py): CUDA cores / tensor cores with hi/lo-split operands


def grad_tol(arm, key):
    """fp32 arm: 2e-4 of the tensor's norm everywhere (measured: <= 5e-6).  bf16x3 arm: the split-operand convs carry ~1e-5 per
    layer (5e-4 allowed), and their pre-activation error (~5e-7 absolute) exceeds the 1e-6 kink margin of the cases now and then:
    one LeakyReLU element taking the other slope moves every gradient UPSTREAM of that layer by up to ~1e-2 of its norm
    (oracle/backward.py: kink_margin), so the tensors behind a LeakyReLU get 3e-2 in that arm."""
    if arm == "fp32":
        return GRAD_RTOL
    exposed = key.startswith(("hyper_encoder.", "hyper_decoder.", "context_model.", "entropy_parameters.net.0", "entropy_parameters.net.2"))
    return 3e-2 if exposed else 5e-4


def rel_err(got, ref):
    ref = ref.double()
    return float((got.double().cpu() - ref).norm() / max(float(ref.norm()), 1e-30))


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("case", H.train_cases())
def test_training_step_matches_reference_golden(case, arm):
    """The committed vectors of the REAL reference's backward + Adam step (sampled entries + norms)."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    from neural_image_compression_b200.training import Adam
    g = H.load_golden(case)
    M, K, init = int(g["M"]), int(g["K"]), str(g["init"])
    model = H.seeded_model(M, K, init, precision="fp32").cuda()
    model.train_precision = arm
    x, nz, ny = (torch.from_numpy(g[k]).cuda() for k in ("x", "noise_z", "noise_y"))
    opt = Adam(model.parameters(), lr=1e-4)
    opt.zero_grad()
    out = model(x, noise=(nz, ny))                      # training=True is the default, as in Trainer.py:82
    rd = rd_loss(out, x, 0.005)
    assert abs(float(rd["loss"].detach()) - float(g["loss"])) <= (1e-5 if arm == "fp32" else 5e-5) * abs(float(g["loss"]))
    rd["loss"].backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    torch.cuda.synchronize()
    assert len(grads) == 59
    # the Adam kernel itself: restated torch.optim.Adam applied to OUR gradients
    want, _ = OB.adam_step({k: v.cpu() for k, v in before.items()}, {k: v.cpu() for k, v in grads.items()})
    for k, p in model.named_parameters():
        gr = grads[k]
        idx = H.sample_index(gr.numel())
        gn = float(g["gnorm_" + k])
        tol = grad_tol(arm, k)
        assert abs(float(gr.double().norm()) - gn) <= tol * gn + 1e-12, (k, float(gr.double().norm()), gn)
        g_atol = 10 * tol * gn / gr.numel() ** 0.5
        np.testing.assert_allclose(gr.reshape(-1)[idx].cpu().numpy(), g["gsamp_" + k], rtol=5 * tol, atol=g_atol, err_msg=k)
        np.testing.assert_allclose(p.detach().cpu().numpy(), want[k].numpy(), rtol=1e-6, atol=1e-7, err_msg=k)
        # against the reference's updated parameters: the first Adam step is lr * g / (|g| + eps), whose sensitivity to an
        # absolute gradient error dg is lr * eps * dg / (|g| + eps)^2 - large only where the gradient itself is ~eps = 1e-8
        gs = np.abs(g["gsamp_" + k].astype(np.float64))
        p_atol = 2e-6 + 1e-4 * 1e-8 * g_atol / (gs + 1e-8) ** 2
        err = np.abs(p.detach().reshape(-1)[idx].cpu().numpy().astype(np.float64) - g["psamp_" + k])
        bad = err > np.minimum(p_atol, 2.1e-4)
        assert not bad.any(), (k, err[bad], p_atol[bad], gs[bad], g_atol)


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("K,shape", [(3, (2, 3, 128, 192)), (1, (3, 3, 64, 64)), (2, (1, 3, 192, 64))])
def test_every_gradient_tensor_against_the_oracle(K, shape, arm):
    """All 59 gradient tensors in full against the oracle's autograd on the same weights, input and noise."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = H.seeded_model(128, K, "calib", precision="fp32")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = H.seeded_input(shape)
    B, _, Hh, W = shape
    # a noise draw that keeps every LeakyReLU pre-activation >= 1e-6 from its kink (oracle/backward.py: kink_margin) - closer than
    # the fp32 accumulation noise, the slope an element gets depends on summation order, in the reference's own runs too
    _, nz, ny = OB.noise_with_margin(sd, x, 128, K, 21)
    ref_rd, ref_g, ref_out = OB.loss_and_grads(sd, x, 128, K, nz, ny, 0.005)
    model = model.cuda()
    model.train_precision = arm
    out = model(x.cuda(), training=True, noise=(nz.cuda(), ny.cuda()))
    rd = rd_loss(out, x.cuda(), 0.005)
    assert abs(float(rd["loss"].detach()) - ref_rd["loss"]) <= (1e-5 if arm == "fp32" else 5e-5) * abs(ref_rd["loss"])
    assert rel_err(out["x_hat"].detach(), ref_out["x_hat"]) < (1e-5 if arm == "fp32" else 1e-4)
    rd["loss"].backward()
    torch.cuda.synchronize()
    worst = {}
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        worst[k] = rel_err(p.grad, ref_g[k])
    bad = {k: v for k, v in worst.items() if v > grad_tol(arm, k)}
    print(arm, "worst gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    assert not bad, bad
    # the dict entries the reference returns are all there, the non-differentiable ones detached
    assert out["logp_y"].requires_grad and out["x_hat"].requires_grad and not out["y_in"].requires_grad
    assert set(out) >= {"x_hat", "y", "y_in", "z", "z_in", "p_z", "logp_z", "p_y", "logp_y", "training"}


def test_gradient_accumulates_and_is_deterministic():
    """Two backward passes accumulate into .grad like torch; a bf16x3 model's training call takes the differentiable path
    (default step arm for M = 128: tensor-core convs)."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = H.seeded_model(128, 3, "calib", precision="bf16x3").cuda()
    x = H.seeded_input((1, 3, 64, 64)).cuda()
    torch.manual_seed(5)
    noise = (torch.rand(1, 128, 1, 1).cuda() - 0.5, torch.rand(1, 128, 4, 4).cuda() - 0.5)
    rd_loss(model(x, noise=noise), x, 0.005)["loss"].backward()
    g1 = model.encoder.net[6].weight.grad.clone()
    rd_loss(model(x, noise=noise), x, 0.005)["loss"].backward()
    torch.cuda.synchronize()
    assert torch.allclose(model.encoder.net[6].weight.grad, 2 * g1, rtol=1e-6, atol=0)
    # determinism: the split-K folds have a fixed order
    model.zero_grad()
    rd_loss(model(x, noise=noise), x, 0.005)["loss"].backward()
    assert torch.equal(model.encoder.net[6].weight.grad, g1)


def test_wgrad_kernel_against_torch_autograd_shapes():
    """nic_conv_wgrad on the layer shapes of the path (incl. the 3-channel NCHW sides, 192 / 640 / 1152 channels, 1x1 and 3x3)."""
    import torch.nn as nn
    import torch.nn.functional as F
    from neural_image_compression_b200 import _lib
    from neural_image_compression_b200.training import conv_dgrad, conv_wgrad
    torch.manual_seed(3)
    cases = [(nn.Conv2d(3, 128, 5, 2, 2), 2, 32, 48, True), (nn.Conv2d(128, 128, 5, 2, 2), 2, 20, 24, False),
             (nn.ConvTranspose2d(128, 192, 5, 2, 2, 1), 1, 6, 10, False), (nn.ConvTranspose2d(128, 3, 5, 2, 2, 1), 2, 8, 12, False),
             (nn.Conv2d(192, 256, 3, 1, 1), 1, 9, 7, False), (nn.Conv2d(640, 1152, 1), 2, 5, 6, False),
             (nn.Conv2d(128, 256, 5, 1, 2), 1, 8, 8, False)]
    for conv, n, h, w, x_nchw in cases:
        conv = conv.double()                               # float64 CPU autograd is the reference of this kernel-level check
        x = torch.randn(n, conv.in_channels, h, w, dtype=torch.float64, requires_grad=True)
        y = conv(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        ref_dw, ref_db, ref_dx = conv.weight.grad.clone(), conv.bias.grad.clone(), x.grad.clone()
        conv = conv.float().cuda()
        x, gy = x.detach().float().cuda(), gy.float().cuda()
        out_nchw = conv.out_channels == 3
        xi = x.detach() if x_nchw else x.detach().permute(0, 2, 3, 1).contiguous()
        gi = gy if out_nchw else gy.permute(0, 2, 3, 1).contiguous()
        dw, db = conv_wgrad(conv, xi, gi, n, h, w, _lib.LAYOUT_NCHW if x_nchw else _lib.LAYOUT_NHWC,
                            _lib.LAYOUT_NCHW if out_nchw else _lib.LAYOUT_NHWC)
        assert rel_err(dw, ref_dw) < 5e-6, (conv, rel_err(dw, ref_dw))
        assert rel_err(db, ref_db) < 5e-6, conv
        if conv.in_channels >= 16:
            dx = conv_dgrad(conv, gi, n, h, w, _lib.LAYOUT_NCHW if out_nchw else _lib.LAYOUT_NHWC)
            assert rel_err(dx.permute(0, 3, 1, 2), ref_dx) < 5e-6, (conv, rel_err(dx.permute(0, 3, 1, 2), ref_dx))


def test_wgrad_tensor_core_kernel_against_float64_autograd():
    """nic_conv_wgrad_tc (tcgen05, MN-major operands straight from the NHWC tensors, hi/lo-split): every layer type it is built for."""
    import torch.nn as nn
    from neural_image_compression_b200 import _lib
    from neural_image_compression_b200.training import conv_wgrad
    torch.manual_seed(4)
    cases = [(nn.Conv2d(128, 128, 1), 2, 16, 16), (nn.Conv2d(128, 128, 3, 1, 1), 2, 16, 24), (nn.Conv2d(128, 256, 5, 1, 2), 1, 16, 16),
             (nn.Conv2d(128, 128, 5, 2, 2), 2, 32, 48), (nn.ConvTranspose2d(128, 128, 5, 2, 2, 1), 2, 8, 16),
             (nn.Conv2d(640, 1152, 1), 2, 8, 16), (nn.Conv2d(128, 128, 5, 2, 2), 3, 40, 24),
             (nn.ConvTranspose2d(128, 192, 5, 2, 2, 1), 1, 8, 16), (nn.Conv2d(192, 256, 3, 1, 1), 2, 16, 8), (nn.Conv2d(192, 192, 5, 2, 2), 1, 32, 32),
             (nn.Conv2d(64, 128, 5, 1, 2), 1, 8, 8)]
    results = []
    for conv, n, h, w in cases:
        conv = conv.double()
        x = torch.randn(n, conv.in_channels, h, w, dtype=torch.float64, requires_grad=True)
        y = conv(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        ref_dw, ref_db = conv.weight.grad.clone(), conv.bias.grad.clone()
        conv = conv.float().cuda()
        xi = x.detach().float().cuda().permute(0, 2, 3, 1).contiguous()
        gi = gy.float().cuda().permute(0, 2, 3, 1).contiguous()
        d = __import__("neural_image_compression_b200.engine", fromlist=["ConvOp"]).ConvOp(conv).desc(
            n, h, w, "fp32", _lib.LAYOUT_NHWC, _lib.LAYOUT_NHWC, _lib.DT_F32, _lib.DT_F32)
        assert _lib.load().nic_conv_wgrad_tc_workspace_bytes(ctypes_byref(d)) > 0, conv
        dw, db = conv_wgrad(conv, xi, gi, n, h, w, _lib.LAYOUT_NHWC, _lib.LAYOUT_NHWC, arm="bf16x3")
        torch.cuda.synchronize()
        results.append((str(conv), rel_err(dw, ref_dw), rel_err(db, ref_db)))
    print("\n".join(f"{e1:.2e} {e2:.2e} {c}" for c, e1, e2 in results))
    assert all(e1 < 5e-5 and e2 < 5e-6 for _, e1, e2 in results), results


def ctypes_byref(d):
    import ctypes
    return ctypes.byref(d)


def test_graphed_trainer_steps_like_the_eager_one():
    """ShardedTrainer(graph=True): forward + loss + backward + Adam replayed from one CUDA graph.  The noise is drawn inside the
    graph (a new draw per replay), so the check is statistical: same first loss within the noise spread, the step counter
    advances on the device, parameters move by ~lr per step, and the loss goes down over 30 steps like the eager trainer's."""
    from neural_image_compression_b200 import parallel
    x = H.seeded_input((2, 3, 128, 128)).cuda()
    losses = {}
    for mode in (False, True):
        model = H.seeded_model(128, 3, "calib", precision="bf16x3").cuda()
        w0 = model.decoder.net[6].weight.detach().clone()
        tr = parallel.ShardedTrainer(model, 0.005, lr=1e-4, graph=mode)
        ls = []
        for _ in range(30):
            rd = tr.step(x)
            ls.append(float(rd["loss"].detach()))
        torch.cuda.synchronize()
        losses[mode] = ls
        assert tr.optimizer.t == 30 and int(tr.optimizer._t_dev) == 30
        moved = float((model.decoder.net[6].weight.detach() - w0).abs().max())
        assert 5e-4 < moved <= 30 * 3.2e-4, moved            # Adam: a few lr per step at most, and the weights did move
    assert abs(losses[True][0] - losses[False][0]) < 0.02 * losses[False][0], (losses[True][0], losses[False][0])
    assert losses[True][-1] < losses[True][0] and abs(losses[True][-1] - losses[False][-1]) < 0.05 * losses[False][-1], losses


def test_step_gradients_equals_autograd_backward():
    """training.step_gradients (the autograd-free path the graphed trainer captures) deposits the same .grad as loss.backward()."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    from neural_image_compression_b200.training import step_gradients
    model = H.seeded_model(128, 3, "calib", precision="bf16x3").cuda()
    x = H.seeded_input((2, 3, 64, 128)).cuda()
    torch.manual_seed(9)
    noise = (torch.rand(2, 128, 1, 2).cuda() - 0.5, torch.rand(2, 128, 4, 8).cuda() - 0.5)
    rd = rd_loss(model(x, noise=noise), x, 0.005)
    rd["loss"].backward()
    ref = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad()
    loss, per_image, scalars = step_gradients(model, x, 0.005, noise=noise)
    torch.cuda.synchronize()
    assert float(loss) == float(rd["loss"].detach())
    for k, p in model.named_parameters():
        assert torch.equal(p.grad, ref[k]), k


@pytest.mark.parametrize("graph", [False, True])
def test_evaluation_after_training_sees_the_updated_weights(graph):
    """The optimizer writes the parameters behind torch's back (nic_adam_multi_step, also from inside a graph replay): every derived
    cache (packed conv weights of all arms, GDN tables, the factorized table, the masked taps) must follow.  After a few steps the
    fused evaluation path must agree with the oracle run on the model's CURRENT state_dict."""
    from neural_image_compression_b200 import parallel
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = H.seeded_model(128, 3, "calib", precision="bf16x3").cuda()
    x = H.seeded_input((2, 3, 128, 128))
    with torch.no_grad():
        before = rd_loss(model(x.cuda(), training=False), x.cuda(), 0.005)["bpp_total"]       # fills the evaluation arm's caches
    tr = parallel.ShardedTrainer(model, 0.005, lr=1e-4, graph=graph)
    for _ in range(6):
        tr.step(x.cuda())
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    assert float(sd["context_model.masked.weight"][:, :, 3:].abs().max()) > 0      # Adam moved the masked taps (as in the reference) ...
    ref = O.forward(sd, x, 128, 3)
    ref_rd = O.rd_loss(ref, x, 0.005)
    band = abs(ref_rd["bpp_total"] - O.rd_loss(O.forward(sd, x, 128, 3, dtype=torch.float64), x, 0.005)["bpp_total"])   # conditioning of the moved weights
    with torch.no_grad():
        out = model(x.cuda(), training=False)
        rd = rd_loss(out, x.cuda(), 0.005)
    assert float(model.context_model.masked.weight.detach()[:, :, 3:].abs().max()) == 0     # ... and the forward zeroes them again (ContextModels.py:19)
    real, ties = H.symbol_mismatches(out["y_in"].cpu().numpy(), ref["y_in"].numpy(), ref["y"].numpy(), 2e-3)
    assert real == 0, (real, ties)
    assert abs(rd["bpp_total"] - ref_rd["bpp_total"]) <= H.BPP_TOL + band and abs(rd["psnr"] - ref_rd["psnr"]) <= H.PSNR_TOL, (rd, ref_rd, band)
    assert abs(rd["bpp_total"] - before) > 1e-3, "six steps at lr 1e-4 must have changed the rate"


def test_adam_kernel_over_several_steps():
    """training.Adam (nic_adam_multi_step with the device-side step counter) against the restated torch.optim.Adam for t = 1..6,
    on tensors of odd sizes (one launch for all of them)."""
    from neural_image_compression_b200.training import Adam
    torch.manual_seed(4)
    shapes = [(7, 5), (11,), (4099,), (3, 128, 5, 5), (1,)]
    params = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    cur = {str(i): p.detach().cpu().clone() for i, p in enumerate(params)}
    opt, state = Adam(params, lr=1e-4), None
    for t in range(6):
        grads = {str(i): torch.randn(*s) * (10.0 ** (t - 3)) for i, s in enumerate(shapes)}
        for i, p in enumerate(params):
            p.grad = grads[str(i)].cuda()
        opt.step()
        cur, state = OB.adam_step(cur, grads, state)
        torch.cuda.synchronize()
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.detach().cpu().numpy(), cur[str(i)].numpy(), rtol=2e-6, atol=1e-8, err_msg=f"tensor {i} step {t + 1}")
    assert opt.t == 6 and int(opt._t_dev) == 6


@pytest.mark.parametrize("K", [1, 2, 3, 5])
def test_likelihood_backward_kernel_against_autograd(K):
    """nic_gm_likelihood_bwd on random latents / raw parameters (incl. elements at the 1e-9 clamp, where the reference's clamp_min
    passes no gradient) against torch autograd over the oracle's split_parameters + conditional_likelihood + log."""
    from neural_image_compression_b200 import _lib
    from neural_image_compression_b200._lib import check, current_stream, ptr
    torch.manual_seed(40 + K)
    b, m, h, w = 2, 16, 6, 10
    y = torch.round(4 * torch.randn(b, m, h, w))
    y[0, 0, 0, :4] = 60.0                                          # far tails: mass below the clamp
    raw = torch.randn(b, (2 if K == 1 else 3 * K) * m, h, w)
    gl = torch.randn(b, m, h, w)
    yr, rr = y.double().requires_grad_(True), raw.double().requires_grad_(True)
    p = O.conditional_likelihood(yr, O.split_parameters(rr, m, K), K)
    (torch.log(p) * gl.double()).sum().backward()
    clamped = int((p.detach() <= 1e-9).sum())
    assert clamped >= 4
    dy = torch.empty_like(y).cuda()
    draw = torch.empty_like(raw).cuda()
    yc, rc, gc = y.cuda(), raw.cuda(), gl.cuda()                   # named: a temporary's memory would be recycled before the launch
    check(_lib.load().nic_gm_likelihood_bwd(ptr(yc), ptr(rc), ptr(gc), 0.0, b, m, h * w, K, ptr(dy), ptr(draw), current_stream()),
          "nic_gm_likelihood_bwd")
    torch.cuda.synchronize()
    # the kernel evaluates the reference's fp32 erf difference: p carries ~1e-7 of absolute rounding noise, so d log p = dp / p is
    # conditioned like 1e-7 / p.  Compare tightly where p > 1e-3, by norm everywhere (the float64 autograd is the yardstick).
    good = (p.detach() > 1e-3)
    e_y = (dy.cpu().double() - yr.grad)[good].abs().max() / yr.grad[good].abs().max()
    assert float(e_y) < 5e-4, float(e_y)
    planes = raw.shape[1] // m
    gmask = good.unsqueeze(1).expand(b, planes, m, h, w)
    dr, rr_g = draw.cpu().double().view(b, planes, m, h, w)[gmask], rr.grad.view(b, planes, m, h, w)[gmask]
    e_r = (dr - rr_g).abs().max() / rr_g.abs().max()
    assert float(e_r) < 5e-4, float(e_r)                            # in the tails (p << 1e-3) fp32 and float64 gradients differ by O(1): not compared
    assert float(dy.cpu()[0, 0, 0, :4].abs().max()) == 0.0 and float(draw.cpu()[0, :, 0, :4][::m].abs().max()) == 0.0     # clamped: no gradient


def test_factorized_backward_kernel_against_autograd():
    """nic_factorized_likelihood_bwd (per-channel 1-3-3-3-1 MLP, softplus / tanh chains, 43 block-summed parameter gradients)."""
    from neural_image_compression_b200 import _lib
    from neural_image_compression_b200._lib import check, current_stream, ptr
    from neural_image_compression_b200.EntropyModels import FactorizedEntropyBottleneck
    from neural_image_compression_b200.training import _FACT_SLICES
    torch.manual_seed(50)
    c, b, h, w = 12, 3, 5, 7
    fe = FactorizedEntropyBottleneck(c)
    with torch.no_grad():
        for f in fe.factors:
            f.uniform_(-0.5, 0.5)                                  # the default init (zeros) would hide the tanh chain
        for mtx in fe.matrices:
            mtx.add_(0.3 * torch.randn_like(mtx))
    sd = {"factorized_entropy_model." + k: v.detach().double().requires_grad_(True) for k, v in fe.state_dict().items()}
    z = (3 * torch.randn(b, c, h, w) + torch.rand(b, c, h, w) - 0.5)
    gl = torch.randn(b, c, h, w)
    zr = z.double().requires_grad_(True)
    prev, O.DIFFERENTIABLE = O.DIFFERENTIABLE, True
    try:
        p = O.factorized_likelihood(sd, zr, dtype=torch.float64)
        (torch.log(p) * gl.double()).sum().backward()
    finally:
        O.DIFFERENTIABLE = prev
    fe = fe.cuda()
    dz = torch.empty_like(z).cuda()
    dpar = torch.empty(c, 43, device="cuda")
    zc, gc, fp = z.cuda(), gl.cuda(), fe.packed()
    check(_lib.load().nic_factorized_likelihood_bwd(ptr(zc), ptr(fp), ptr(gc), 0.0, b, c, h * w, ptr(dz), ptr(dpar),
                                                    current_stream()), "nic_factorized_likelihood_bwd")
    torch.cuda.synchronize()
    assert rel_err(dz, zr.grad) < 1e-4, rel_err(dz, zr.grad)
    for name, idx, lo, hi in _FACT_SLICES:
        ref = sd[f"factorized_entropy_model.{name}.{idx}"].grad
        assert rel_err(dpar[:, lo:hi].reshape(ref.shape), ref) < 1e-4, (name, idx, rel_err(dpar[:, lo:hi].reshape(ref.shape), ref))


def test_adam_is_a_torch_optimizer_with_reference_checkpoint_format():
    """training.Adam against torch.optim.Adam itself: identical updates over 5 steps with an lr schedule (CosineAnnealingLR writes
    param_groups[0]['lr'], Trainer.py:33-36), and state_dict() / load_state_dict() interchange in both directions mid-run
    (the reference's checkpoint carries optimizer.state_dict(), Trainer.py:52-68)."""
    from neural_image_compression_b200.training import Adam
    torch.manual_seed(21)
    shapes = [(33, 7), (129,), (2, 64, 5, 5)]
    ours = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    a, t = Adam(ours, lr=1e-3), torch.optim.Adam(theirs, lr=1e-3)
    sa, st = torch.optim.lr_scheduler.CosineAnnealingLR(a, T_max=8), torch.optim.lr_scheduler.CosineAnnealingLR(t, T_max=8)

    def one_step(a, t):
        for p, q in zip(a.param_groups[0]["params"], t.param_groups[0]["params"]):
            g = torch.randn_like(p)
            p.grad, q.grad = g, g.clone()
        a.step(); t.step()

    for _ in range(3):
        one_step(a, t); sa.step(); st.step()
    assert abs(a.param_groups[0]["lr"] - t.param_groups[0]["lr"]) < 1e-12 and a.param_groups[0]["lr"] < 1e-3
    for p, q in zip(ours, theirs):
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=3e-6, atol=1e-7)
    # ours -> torch: a fresh torch.optim.Adam resumes from our checkpoint;  torch -> ours: and the other way round
    t2 = torch.optim.Adam(theirs, lr=5e-2); t2.load_state_dict(a.state_dict())
    a2 = Adam(ours, lr=5e-2); a2.load_state_dict(t.state_dict())
    assert a2.t == 3 and abs(a2.lr - t.param_groups[0]["lr"]) < 1e-12 and abs(t2.param_groups[0]["lr"] - a.lr) < 1e-12
    for _ in range(2):
        one_step(a2, t2)
    torch.cuda.synchronize()
    assert a2.t == 5 and int(a2._t_dev) == 5 and float(t2.state[theirs[0]]["step"]) == 5
    for p, q in zip(ours, theirs):
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=3e-6, atol=1e-7)


def test_graphed_trainer_follows_a_learning_rate_change():
    """The learning rate of a captured Adam launch lives on the device: lr = 0 between replays freezes the weights, and restoring
    it moves them again - without re-capturing (a by-value lr would be baked into the graph)."""
    from neural_image_compression_b200 import parallel
    x = H.seeded_input((2, 3, 64, 64)).cuda()
    model = H.seeded_model(128, 3, "calib", precision="bf16x3").cuda()
    tr = parallel.ShardedTrainer(model, 0.005, lr=1e-4, graph=True)
    tr.step(x); tr.step(x)
    torch.cuda.synchronize()
    ngraphs = len(tr._graphs)
    w1 = model.decoder.net[6].weight.detach().clone()
    tr.optimizer.param_groups[0]["lr"] = 0.0
    tr.step(x); torch.cuda.synchronize()
    assert torch.equal(model.decoder.net[6].weight.detach(), w1)
    tr.optimizer.param_groups[0]["lr"] = 1e-4
    tr.step(x); torch.cuda.synchronize()
    assert not torch.equal(model.decoder.net[6].weight.detach(), w1) and len(tr._graphs) == ngraphs


def test_graphed_evaluator_recaptures_after_a_weight_change():
    """ShardedEvaluator(graph=True) held across optimizer steps / load_state_dict must not replay against stale packed weights."""
    from neural_image_compression_b200 import parallel
    x = H.seeded_input((2, 3, 64, 128)).cuda()
    model = H.seeded_model(128, 3, "calib", precision="bf16x3").cuda()
    ev = parallel.ShardedEvaluator(model, 0.005, lean=True, graph=True)
    _, t0 = ev.step(x)
    _, t0b = ev.step(x)
    assert float(t0["bpp_total"]) == float(t0b["bpp_total"])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["entropy_parameters.net.4.bias"] = sd["entropy_parameters.net.4.bias"] + 0.5
    model.load_state_dict(sd)
    _, t1 = ev.step(x)
    fresh = parallel.ShardedEvaluator(model, 0.005, lean=True, graph=False)
    _, t2 = fresh.step(x)
    assert float(t1["bpp_total"]) == float(t2["bpp_total"]) and abs(float(t1["bpp_total"]) - float(t0["bpp_total"])) > 1e-4
