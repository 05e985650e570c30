"""GPU: each C-ABI kernel against the CPU oracle on seeded inputs (bit-level where the op is exact)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import forward as O
from oracle.gdn import gdn_effective
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from neural_image_compression_b200 import _lib
    _lib.check(_lib.load().nic_check_device(), "nic_check_device")
    return torch.device("cuda:0")


# ---------------------------------------------------------------------------------------------------
# likelihood kernels
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,shape", [(3, (2, 128, 8, 12)), (1, (2, 128, 8, 12)), (3, (1, 16, 3, 5)), (2, (3, 8, 4, 4))])
def test_gm_likelihood_matches_oracle(dev, K, shape):
    from neural_image_compression_b200.EntropyModels import gm_likelihood
    from neural_image_compression_b200._lib import Q_ROUND
    B, M, Hh, Ww = shape
    torch.manual_seed(2)
    y = 5 * torch.randn(shape)
    raw = torch.randn(B, (2 if K == 1 else 3 * K) * M, Hh, Ww)
    r = gm_likelihood(y.to(dev), raw.to(dev), M, K, Q_ROUND)
    y_in = torch.round(y)
    params = O.split_parameters(raw, M, K)
    p_ref = O.conditional_likelihood(y_in, params, K)
    assert torch.equal(r["y_in"].cpu(), y_in)                       # rint == torch.round, bit exact incl. -0.0
    assert np.array_equal(np.signbit(r["y_in"].cpu().numpy()), np.signbit(y_in.numpy()))
    bad, worst = H.likelihood_close(r["p"].cpu().numpy(), p_ref.numpy())
    assert bad == 0, f"{bad} likelihoods outside tolerance (max abs err {worst:.3e})"
    big = p_ref > 1e-6
    np.testing.assert_allclose(r["logp"].cpu().numpy()[big], torch.log(p_ref).numpy()[big], rtol=0, atol=0.5)
    names = ("mu", "sigma") if K == 1 else ("weights", "mus", "sigmas")
    for name, ref in zip(names, params):
        np.testing.assert_allclose(r[name].cpu().numpy(), ref.numpy(), rtol=2e-6, atol=1e-7, err_msg=name)
    # per-image sums of logp (deterministic partials)
    s = r["partials"].double().sum(dim=1).cpu()
    ref_s = torch.log(torch.from_numpy(r["p"].cpu().numpy())).double().sum(dim=(1, 2, 3))
    np.testing.assert_allclose(s.numpy(), ref_s.numpy(), rtol=1e-5)


@pytest.mark.parametrize("K,shape,full", [(3, (5, 128, 8, 12), True), (3, (16, 128, 32, 48), True), (1, (3, 192, 17, 20), True),
                                          (2, (7, 16, 3, 5), False), (5, (2, 8, 4, 4), True), (3, (33, 8, 4, 4), False)])
def test_gm_likelihood_kernel_forms_agree(dev, K, shape, full, monkeypatch):
    """The flat balanced form (default: one list of chunks cut into equal ranges, a block folds its sum at every image
    boundary it crosses; erff with both ranges evaluated) against the (parts, B) grid with libdevice's erff: element outputs bit-identical (ragged sizes, blocks spanning several images, more images than slots per block),
    per-image sums equal to rounding, and p against the oracle."""
    from neural_image_compression_b200.EntropyModels import gm_likelihood
    from neural_image_compression_b200._lib import Q_NOISE, Q_ROUND
    B, M, Hh, Ww = shape
    g = torch.Generator().manual_seed(11)
    y = 5 * torch.randn(shape, generator=g)
    raw = torch.randn((B, (2 if K == 1 else 3 * K) * M, Hh, Ww), generator=g)
    noise = torch.rand(shape, generator=g) - 0.5
    for qmode, nz in ((Q_ROUND, None), (Q_NOISE, noise)):
        outs = []
        for env in ({"NIC_LIK_FLAT": "0"}, {}):
            monkeypatch.delenv("NIC_LIK_FLAT", raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            r = gm_likelihood(y.to(dev), raw.to(dev), M, K, qmode, noise=None if nz is None else nz.to(dev), full=full)
            outs.append({k: v.cpu() for k, v in r.items() if v is not None})
        for form in outs[1:]:
            for name, ref in outs[0].items():
                if name == "partials":
                    np.testing.assert_allclose(form[name].double().sum(1).numpy(), ref.double().sum(1).numpy(), rtol=1e-5, atol=1e-3)
                else:
                    assert torch.equal(form[name], ref), name
        y_in = torch.round(y) if qmode == Q_ROUND else y + noise
        assert torch.equal(outs[1]["y_in"], y_in)
        p_ref = O.conditional_likelihood(y_in, O.split_parameters(raw, M, K), K)
        bad, worst = H.likelihood_close(outs[1]["p"].numpy(), p_ref.numpy())
        assert bad == 0, f"{bad} likelihoods outside tolerance (max abs err {worst:.3e})"
        s = outs[1]["partials"].double().sum(dim=1)
        np.testing.assert_allclose(s.numpy(), torch.log(outs[1]["p"]).double().sum(dim=(1, 2, 3)).numpy(), rtol=1e-5)


def test_gm_likelihood_properties(dev):
    """p in [1e-9, 1]; mixture weights sum to 1; K = 1 mass is monotone in |y - mu|."""
    from neural_image_compression_b200.EntropyModels import gm_likelihood
    from neural_image_compression_b200._lib import Q_PASSTHRU
    torch.manual_seed(3)
    y = torch.round(20 * torch.randn(2, 32, 4, 8)).to(dev)
    raw = (3 * torch.randn(2, 9 * 32, 4, 8)).to(dev)
    r = gm_likelihood(y, raw, 32, 3, Q_PASSTHRU)
    assert float(r["p"].min()) >= float(np.float32(1e-9)) and float(r["p"].max()) <= 1.0 + 1e-6
    assert torch.allclose(r["weights"].sum(dim=1), torch.ones_like(y), atol=1e-6)
    assert float(r["sigmas"].min()) >= 1e-6
    d = torch.arange(0, 8, dtype=torch.float32, device=dev).reshape(1, 8, 1, 1)
    raw1 = torch.zeros(1, 16, 1, 1, device=dev)                  # mu = 0, sigma = softplus(0) + 1e-6
    p = gm_likelihood(d.expand(1, 8, 1, 1).contiguous(), raw1, 8, 1, Q_PASSTHRU)["p"].flatten()
    assert torch.all(p[1:] <= p[:-1])


def test_gm_likelihood_noise_and_lean(dev):
    from neural_image_compression_b200.EntropyModels import gm_likelihood
    from neural_image_compression_b200._lib import Q_NOISE
    torch.manual_seed(4)
    y, raw = torch.randn(2, 128, 4, 4), torch.randn(2, 9 * 128, 4, 4)
    noise = torch.rand(2, 128, 4, 4) - 0.5
    r = gm_likelihood(y.to(dev), raw.to(dev), 128, 3, Q_NOISE, noise=noise.to(dev), full=False)
    assert torch.equal(r["y_in"].cpu(), y + noise) and "weights" not in r
    p_ref = O.conditional_likelihood(y + noise, O.split_parameters(raw, 128, 3), 3)
    assert H.likelihood_close(r["p"].cpu().numpy(), p_ref.numpy())[0] == 0


def test_gm_likelihood_empty_batch(dev):
    from neural_image_compression_b200.EntropyModels import gm_likelihood
    r = gm_likelihood(torch.zeros(0, 8, 4, 4, device=dev), torch.zeros(0, 72, 4, 4, device=dev), 8, 3, 0)
    assert r["p"].shape == (0, 8, 4, 4)


@pytest.mark.parametrize("shape", [(2, 128, 2, 3), (1, 128, 1, 1), (3, 128, 8, 12)])
def test_factorized_matches_oracle(dev, shape):
    model = H.seeded_model(128, 1, 'plain')
    # move the learned shapes off their init so tanh factors and matrices matter
    torch.manual_seed(5)
    with torch.no_grad():
        for p in model.factorized_entropy_model.parameters():
            p.add_(0.3 * torch.randn_like(p))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    z = 4 * torch.randn(shape)
    fe = model.factorized_entropy_model.to(dev)
    from neural_image_compression_b200._lib import Q_ROUND
    z_in, p, logp, parts = fe.likelihood(z.to(dev), Q_ROUND, want_in=True)
    assert torch.equal(z_in.cpu(), torch.round(z))
    p_ref = O.factorized_likelihood(sd, torch.round(z))
    bad, worst = H.likelihood_close(p.cpu().numpy(), p_ref.numpy())
    assert bad == 0, f"{bad} outside tolerance, max abs err {worst:.3e}"
    assert torch.allclose(fe(torch.round(z).to(dev)).cpu(), p.cpu())           # module call == kernel output
    np.testing.assert_allclose(parts.double().sum(1).cpu().numpy(), logp.double().sum(dim=(1, 2, 3)).cpu().numpy(), rtol=1e-5)


def test_conditional_modules_match_oracle(dev):
    from neural_image_compression_b200.EntropyModels import GaussianConditional, GaussianMixtureConditional
    torch.manual_seed(6)
    x = torch.round(5 * torch.randn(2, 16, 4, 4))
    raw = torch.randn(2, 9 * 16, 4, 4)
    w, mu, s = O.split_parameters(raw, 16, 3)
    p = GaussianMixtureConditional()(x.to(dev), weights=w.to(dev), mus=mu.to(dev), sigmas=s.to(dev))
    assert H.likelihood_close(p.cpu().numpy(), O.conditional_likelihood(x, (w, mu, s), 3).numpy())[0] == 0
    p1 = GaussianConditional()(x.to(dev), mu=mu[:, 0].to(dev), sigma=s[:, 0].to(dev))
    assert H.likelihood_close(p1.cpu().numpy(), O.conditional_likelihood(x, (mu[:, 0], s[:, 0]), 1).numpy())[0] == 0


def test_public_pmf_methods_are_unclamped_like_the_reference(dev):
    """GaussianConditional.discretized_gaussian_pmf / GaussianMixtureConditional.discretized_mixture_pmf (EntropyModels.py:192-230)
    return the raw bin mass: no 1e-9 clamp (it is EntropyModel.forward that clamps, :29-31)."""
    from neural_image_compression_b200.EntropyModels import GaussianConditional, GaussianMixtureConditional
    torch.manual_seed(8)
    x = torch.round(12 * torch.randn(2, 16, 4, 4))                     # far tails: many masses below the clamp
    raw = torch.randn(2, 9 * 16, 4, 4)
    w, mu, s = O.split_parameters(raw, 16, 3)
    ref_k = O.gaussian_pmf(x.unsqueeze(1), mu, s)
    ref = torch.sum(w * ref_k, dim=1)
    assert float(ref.min()) < 1e-9
    gm = GaussianMixtureConditional()
    got = gm.discretized_mixture_pmf(x.to(dev), w.to(dev), mu.to(dev), s.to(dev)).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-4, atol=H.P_ATOL)
    assert float(got.min()) < 1e-9
    got_k = gm.discretized_gaussian_pmf(x.to(dev).unsqueeze(1), mu.to(dev), s.to(dev)).cpu()        # the reference's own inner call
    assert got_k.shape == ref_k.shape
    np.testing.assert_allclose(got_k.numpy(), ref_k.numpy(), rtol=1e-4, atol=H.P_ATOL)
    got1 = GaussianConditional().discretized_gaussian_pmf(x.to(dev), mu[:, 0].to(dev), s[:, 0].to(dev)).cpu()
    np.testing.assert_allclose(got1.numpy(), O.gaussian_pmf(x, mu[:, 0], s[:, 0]).numpy(), rtol=1e-4, atol=H.P_ATOL)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_masked_conv_type_b_keeps_the_centre_tap(dev, precision):
    """MaskedConv2d('B', ...) is accepted by the reference (ContextModels.py:11, 15) though its model only uses 'A'."""
    import torch.nn.functional as F
    from neural_image_compression_b200.ContextModels import MaskedConv2d
    torch.manual_seed(9)
    conv = MaskedConv2d("B", in_channels=128, out_channels=256, kernel_size=5, stride=1, padding=2)
    conv.precision = precision
    assert int(conv.mask[0, 0].sum()) == 13 and float(conv.mask[0, 0, 2, 2]) == 1
    x = torch.round(4 * torch.randn(2, 128, 8, 12))
    ref = F.conv2d(x.double(), (conv.weight.detach() * conv.mask).double(), conv.bias.detach().double(), padding=2)
    got = conv.to(dev)(x.to(dev)).cpu().double()
    assert float((got - ref).abs().max() / ref.abs().max()) < 3e-5
    assert float(conv.weight.detach()[:, :, 2, 3:].abs().max()) == 0 and float(conv.weight.detach()[:, :, 2, 2].abs().max()) > 0


def test_rd_loss_terms_match_oracle(dev):
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    torch.manual_seed(7)
    B = 3
    out = {"x_hat": torch.rand(B, 3, 64, 128), "logp_y": -3 * torch.rand(B, 128, 4, 8), "logp_z": -torch.rand(B, 128, 1, 2)}
    x = torch.rand(B, 3, 64, 128)
    ref = O.rd_loss(out, x, 0.005)
    got = rd_loss({k: v.to(dev) for k, v in out.items()}, x.to(dev), 0.005)
    assert set(got) == set(ref)
    for k in ("bpp_y", "bpp_z", "bpp_total", "mse", "psnr", "bits_y", "bits_z", "bits_total"):
        assert abs(got[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k])), k
    assert abs(float(got["loss"]) - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    np.testing.assert_allclose(got["mse_per_image"].cpu().numpy(), ref["mse_per_image"].numpy(), rtol=1e-5)
    np.testing.assert_allclose(got["psnr_per_image"].cpu().numpy(), ref["psnr_per_image"].numpy(), rtol=1e-5)


# ---------------------------------------------------------------------------------------------------
# transforms (fp32 arm): every layer type on the path against torch CPU convolution
# ---------------------------------------------------------------------------------------------------
def _rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("shape", [(2, 3, 64, 128), (1, 3, 128, 64)])
def test_encoder_fp32_matches_oracle(dev, shape):
    model = H.seeded_model(128, 1, 'gain')
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = H.seeded_input(shape)
    y = model.encoder.to(dev)(x.to(dev)).cpu()
    ref = O.analysis(sd, x)
    assert y.shape == ref.shape and _rel_err(y, ref) < 2e-5


def test_decoder_fp32_matches_oracle(dev):
    model = H.seeded_model(128, 1, 'plain')
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(8)
    y = torch.round(6 * torch.randn(2, 128, 4, 8))
    xh = model.decoder.to(dev)(y.to(dev)).cpu()
    ref = O.synthesis(sd, y)
    assert xh.shape == ref.shape and _rel_err(xh, ref) < 2e-5


def test_hyper_transforms_fp32_match_oracle(dev):
    model = H.seeded_model(128, 1, 'plain')
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(9)
    y = 3 * torch.randn(2, 128, 8, 12)
    z = model.hyper_encoder.to(dev)(y.to(dev)).cpu()
    ref_z = O.hyper_analysis(sd, y)
    assert z.shape == ref_z.shape and _rel_err(z, ref_z) < 2e-5
    zq = torch.round(3 * torch.randn(2, 128, 2, 3))
    psi = model.hyper_decoder.to(dev)(zq.to(dev)).cpu()
    ref_psi = O.hyper_synthesis(sd, zq)
    assert psi.shape == ref_psi.shape and _rel_err(psi, ref_psi) < 2e-5


def test_context_and_entropy_parameters_fp32_match_oracle(dev):
    model = H.seeded_model(128, 3, 'plain')
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(10)
    yq = torch.round(4 * torch.randn(2, 128, 8, 12))
    phi = model.context_model.to(dev)(yq.to(dev)).cpu()
    ref = O.context(sd, yq)
    assert _rel_err(phi, ref) < 2e-5
    # the reference zeroes the masked taps in place on first use (ContextModels.py:19)
    w = model.context_model.masked.weight.detach().cpu()
    assert float(w[:, :, 2, 2:].abs().max()) == 0 and float(w[:, :, 3:].abs().max()) == 0
    comb = torch.randn(2, 512, 8, 12)
    w_, mu_, s_ = model.entropy_parameters.to(dev)(comb.to(dev))
    rw, rmu, rs = O.split_parameters(O.entropy_parameters_raw(sd, comb), 128, 3)
    for a, b in ((w_, rw), (mu_, rmu), (s_, rs)):
        assert a.shape == b.shape and _rel_err(a.cpu(), b) < 5e-5


def test_gdn_standalone_matches_oracle(dev):
    from neural_image_compression_b200.gdn import GDN
    torch.manual_seed(11)
    for inverse in (False, True):
        g = GDN(128, inverse=inverse)
        with torch.no_grad():
            g.gamma.add_(0.02 * torch.rand_like(g.gamma)); g.beta.add_(0.1 * torch.rand_like(g.beta))
        x = torch.randn(2, 128, 5, 7)
        beta, gamma = gdn_effective(g.beta.detach(), g.gamma.detach())
        norm = F.conv2d(x * x, gamma.reshape(128, 128, 1, 1), beta)
        ref = x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
        got = g.to(dev)(x.to(dev)).cpu()
        assert _rel_err(got, ref) < 1e-5


def test_conv_empty_batch_and_bad_shapes(dev):
    from neural_image_compression_b200 import _lib
    model = H.seeded_model(128, 1, 'plain').to(dev)
    assert model.encoder(torch.zeros(0, 3, 64, 64, device=dev)).shape == (0, 128, 4, 4)
    with pytest.raises(ValueError):
        model(torch.zeros(1, 3, 60, 64, device=dev), training=False)
    with pytest.raises(ValueError):
        model(torch.zeros(1, 1, 64, 64, device=dev), training=False)


def test_rd_loss_resums_when_logp_was_modified_in_place(dev):
    """The likelihood kernel's per-image partial sums ride along with logp (engine.attach_partials); a caller that edits logp in place
     must get the loss of the EDITED tensor, not of the stale sums."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = H.seeded_model(128, 3, "plain", precision="fp32").to(dev)
    x = H.seeded_input((1, 3, 64, 64)).to(dev)
    out = model(x, training=False)
    a = rd_loss(out, x, 0.005)["bpp_y"]
    out["logp_y"].mul_(2.0)
    b = rd_loss(out, x, 0.005)["bpp_y"]
    assert abs(b - 2 * a) < 1e-5 * abs(a)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16", "fp32"])
@pytest.mark.parametrize("K", [3, 1])
def test_ctx_ep_one_call_equals_the_separate_calls(dev, precision, K):
    """nic_ctx_ep_fwd (context conv + entropy-parameter stack + likelihoods behind ONE C-ABI call) against the
    same work as four nic_conv_fwd calls + nic_gm_likelihood_fwd: bit-identical raw parameters and likelihood outputs, and the raw
    parameters against the oracle (ContextModels.py:15-20, Models.py:73, ParametersModels.py:29-35)."""
    from neural_image_compression_b200 import _lib, engine
    from neural_image_compression_b200.EntropyModels import gm_likelihood
    M, B, hy, wy = 128, 2, 8, 12
    model = H.seeded_model(M, K, "calib", precision=precision)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    torch.manual_seed(5)
    yq = torch.round(4 * torch.randn(B, hy, wy, M))
    psi = torch.randn(B, hy, wy, 2 * M)
    pair, adt = precision == "bf16x3", engine.act_dtype(precision)
    y_eng = (engine.to_pair(yq) if pair else yq.to(adt)).to(dev)
    y_nchw = yq.permute(0, 3, 1, 2).contiguous().to(dev)

    def fresh_combined():
        c = torch.zeros((B, hy, wy, (2 if pair else 1) * 4 * M), dtype=adt, device=dev)
        if pair:
            pp = engine.to_pair(psi).to(dev)
            c[..., 2 * M:4 * M] = pp[..., :2 * M]; c[..., 6 * M:8 * M] = pp[..., 2 * M:]
        else:
            c[..., 2 * M:] = psi.to(adt).to(dev)
        return c
    model.context_model.masked.apply_mask_()
    ctx_op, ep = model.context_model.masked._op, model.entropy_parameters.ops
    with torch.no_grad():
        c1 = fresh_combined()
        ctx_op.run(y_eng, B, hy, wy, precision, out=c1, out_c_total=4 * M, out_c_offset=0)
        a = ep[0].run(c1, B, hy, wy, precision)
        a = ep[1].run(a, B, hy, wy, precision)
        raw1 = ep[2].run(a, B, hy, wy, precision, out_layout=_lib.LAYOUT_NCHW, out_dtype=torch.float32)
        l1 = gm_likelihood(y_nchw, raw1, M, K, _lib.Q_PASSTHRU, full=True, want_y_in=False)
        c2 = fresh_combined()
        plan = engine.CtxEpPlan(ctx_op, ep, B, hy, wy, precision, M, K, dev, full=True)
        raw2, l2 = plan.run(y_eng, c2, y_in=y_nchw, qmode=_lib.Q_PASSTHRU)
        raw3, none = plan.run(y_eng, fresh_combined())                   # without the likelihood epilogue
    torch.cuda.synchronize()
    assert none is None and torch.equal(raw1, raw2) and torch.equal(raw1, raw3) and torch.equal(c1, c2)
    for k in l1:
        if l1[k] is not None:
            assert torch.equal(l1[k], l2[k]), k
    if precision != "bf16":
        phi = O.context(sd, yq.permute(0, 3, 1, 2))
        ref = O.entropy_parameters_raw(sd, torch.cat([phi, psi.permute(0, 3, 1, 2)], dim=1))
        err = float((raw2.cpu() - ref).abs().max() / ref.abs().max())
        assert err < (2e-5 if precision == "fp32" else 1e-4), err
