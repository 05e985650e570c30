"""GPU: the bf16x3 arm (hi/lo-split operands on the tensor cores, fp32 grade) layer by layer against a float64 CPU
convolution of the SAME fp32 operands.  The split keeps 16 mantissa bits per operand, so every layer must land within
~1e-5 of the exact result relative to the tensor's scale; the bound asserted is 1e-4."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import forward as O
from oracle.gdn import gdn_effective
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _run(conv, x_nchw, epilogue, gdn=None, mask_a=False, out_nchw_f32=False, out_f32=False, out_c_total=0, out_c_offset=0,
         from_image=False):
    from neural_image_compression_b200 import engine
    from neural_image_compression_b200._lib import LAYOUT_NCHW, LAYOUT_NHWC
    dev = torch.device("cuda:0")
    op = engine.ConvOp(conv.to(dev), epilogue, gdn=None if gdn is None else gdn.to(dev), mask_a=mask_a)
    n, c, h, w = x_nchw.shape
    if from_image:
        x, in_layout = x_nchw.contiguous().to(dev), LAYOUT_NCHW
    else:
        x, in_layout = engine.to_pair(x_nchw.permute(0, 2, 3, 1).contiguous().to(dev)), LAYOUT_NHWC
    out = None
    if out_c_total:
        ho, wo = engine.conv_out_hw(conv, h, w)
        out = torch.zeros((n, ho, wo, 2 * out_c_total), dtype=torch.bfloat16, device=dev)
    f32 = out_nchw_f32 or out_f32
    y = op.run(x, n, h, w, "bf16x3", in_layout=in_layout, out_layout=LAYOUT_NCHW if out_nchw_f32 else LAYOUT_NHWC, out=out,
               out_c_total=out_c_total, out_c_offset=out_c_offset, out_dtype=torch.float32 if f32 else torch.bfloat16)
    torch.cuda.synchronize()
    if not f32:
        y = engine.from_pair(y)
    y = y.float().cpu()
    return y if out_nchw_f32 else y.permute(0, 3, 1, 2)


def _ref(conv, x, epilogue, gdn=None, mask=None):
    from neural_image_compression_b200._lib import EPI_GDN, EPI_IGDN, EPI_LRELU
    w = conv.weight.detach().cpu().double()
    if mask is not None:
        w = w * mask.double()
    b = conv.bias.detach().cpu().double()
    if isinstance(conv, nn.ConvTranspose2d):
        v = F.conv_transpose2d(x.double(), w, b, stride=conv.stride, padding=conv.padding, output_padding=conv.output_padding)
    else:
        v = F.conv2d(x.double(), w, b, stride=conv.stride, padding=conv.padding)
    if epilogue == EPI_LRELU:
        v = F.leaky_relu(v, 0.01)
    elif epilogue in (EPI_GDN, EPI_IGDN):
        beta, gamma = gdn_effective(gdn.beta.detach().cpu(), gdn.gamma.detach().cpu())
        C = beta.numel()
        norm = F.conv2d(v * v, gamma.double().reshape(C, C, 1, 1), beta.double())
        v = v * (torch.sqrt(norm) if epilogue == EPI_IGDN else torch.rsqrt(norm))
    return v.float()


def _close(y, ref, tol=TOL):
    err = float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-9))
    assert y.shape == ref.shape and err < tol, err
    return err


def _gdn(inverse=False, c=128):
    from neural_image_compression_b200.gdn import GDN
    g = GDN(c, inverse=inverse)
    with torch.no_grad():
        g.gamma.add_(0.02 * torch.rand_like(g.gamma)); g.beta.add_(0.1 * torch.rand_like(g.beta))
    return g


@pytest.mark.parametrize("hw", [(32, 48), (20, 28), (8, 12)])
def test_x3_conv5x5_s2(hw):
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_LRELU
    torch.manual_seed(40)
    conv = nn.Conv2d(128, 128, 5, 2, 2)
    x = torch.randn(2, 128, *hw)
    print(_close(_run(conv, x, EPI_BIAS, out_f32=True), _ref(conv, x, EPI_BIAS)))
    print(_close(_run(conv, x, EPI_LRELU), _ref(conv, x, EPI_LRELU)))          # pair output


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("hw", [(16, 24), (10, 12)])
def test_x3_conv_then_gdn_on_tensor_cores(inverse, hw):
    from neural_image_compression_b200._lib import EPI_GDN, EPI_IGDN
    torch.manual_seed(41)
    g = _gdn(inverse)
    epi = EPI_IGDN if inverse else EPI_GDN
    conv = nn.ConvTranspose2d(128, 128, 5, 2, 2, output_padding=1) if inverse else nn.Conv2d(128, 128, 5, 2, 2)
    x = torch.randn(2, 128, *hw)
    print(_close(_run(conv, x, epi, gdn=g), _ref(conv, x, epi, gdn=g)))


@pytest.mark.parametrize("inverse,hw", [(False, (32, 48)), (True, (16, 24)), (False, (20, 28)), (True, (5, 9))])
def test_x3_192_channels_conv_then_gdn(inverse, hw):
    """The reference's default capacity (Models.py:17: latent_channels = 192): conv with two N tiles + the GDN kernel that
    streams gamma through a ring (gdn_x3c_kernel)."""
    from neural_image_compression_b200._lib import EPI_GDN, EPI_IGDN
    torch.manual_seed(48)
    g = _gdn(inverse, c=192)
    epi = EPI_IGDN if inverse else EPI_GDN
    conv = nn.ConvTranspose2d(192, 192, 5, 2, 2, output_padding=1) if inverse else nn.Conv2d(192, 192, 5, 2, 2)
    x = torch.randn(2, 192, *hw)
    print(_close(_run(conv, x, epi, gdn=g), _ref(conv, x, epi, gdn=g)))


@pytest.mark.parametrize("shape", [(2, 3, 64, 96), (1, 3, 80, 48)])
def test_x3_first_layer_192_channels(shape):
    from neural_image_compression_b200._lib import EPI_GDN
    torch.manual_seed(49)
    conv, g = nn.Conv2d(3, 192, 5, 2, 2), _gdn(c=192)
    x = torch.rand(*shape)
    print(_close(_run(conv, x, EPI_GDN, gdn=g, from_image=True), _ref(conv, x, EPI_GDN, gdn=g)))


@pytest.mark.parametrize("hw", [(16, 24), (7, 15)])
def test_x3_last_layer_192_channels(hw):
    from neural_image_compression_b200._lib import EPI_BIAS
    torch.manual_seed(50)
    conv = nn.ConvTranspose2d(192, 3, 5, 2, 2, output_padding=1)
    x = torch.randn(2, 192, *hw)
    print(_close(_run(conv, x, EPI_BIAS, out_nchw_f32=True), _ref(conv, x, EPI_BIAS)))


@pytest.mark.parametrize("shape", [(2, 3, 64, 96), (1, 3, 80, 48), (3, 3, 32, 32)])
def test_x3_first_layer_from_nchw_image(shape):
    from neural_image_compression_b200._lib import EPI_GDN
    torch.manual_seed(42)
    conv, g = nn.Conv2d(3, 128, 5, 2, 2), _gdn()
    x = torch.rand(*shape)
    print(_close(_run(conv, x, EPI_GDN, gdn=g, from_image=True), _ref(conv, x, EPI_GDN, gdn=g)))


def test_x3_transposed_to_192_and_3x3_into_pair_channel_window():
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_LRELU
    torch.manual_seed(43)
    conv = nn.ConvTranspose2d(128, 192, 5, 2, 2, output_padding=1)
    x = torch.randn(2, 128, 8, 12)
    print(_close(_run(conv, x, EPI_LRELU), _ref(conv, x, EPI_LRELU)))
    conv3 = nn.Conv2d(192, 256, 3, 1, 1)
    x3 = torch.randn(2, 192, 16, 24)
    y = _run(conv3, x3, EPI_BIAS, out_c_total=512, out_c_offset=256)         # pair tensor [hi(512) | lo(512)], window [256, 512)
    assert float(y[:, :256].abs().max()) == 0
    print(_close(y[:, 256:], _ref(conv3, x3, EPI_BIAS)))


def test_x3_masked_context_conv_and_pointwise_stack():
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_LRELU
    torch.manual_seed(44)
    conv = nn.Conv2d(128, 256, 5, 1, 2)
    x = torch.round(4 * torch.randn(2, 128, 16, 24)) + 0.25 * torch.rand(2, 128, 16, 24)
    mask = O.mask_a(conv.weight.detach())
    y = _run(conv, x, EPI_BIAS, mask_a=True, out_c_total=512, out_c_offset=0)
    assert float(y[:, 256:].abs().max()) == 0
    print(_close(y[:, :256], _ref(conv, x, EPI_BIAS, mask=mask)))
    c1 = nn.Conv2d(512, 640, 1)
    x1 = torch.randn(2, 512, 8, 12)
    print(_close(_run(c1, x1, EPI_LRELU), _ref(c1, x1, EPI_LRELU)))
    c3 = nn.Conv2d(640, 1152, 1)
    x3 = torch.randn(2, 640, 8, 12)
    print(_close(_run(c3, x3, EPI_BIAS, out_nchw_f32=True), _ref(c3, x3, EPI_BIAS)))


@pytest.mark.parametrize("hw", [(16, 24), (12, 20), (6, 14), (7, 15), (1, 1), (37, 53)])
def test_x3_last_layer_to_rgb_nchw(hw):
    from neural_image_compression_b200._lib import EPI_BIAS
    torch.manual_seed(45)
    conv = nn.ConvTranspose2d(128, 3, 5, 2, 2, output_padding=1)
    x = torch.randn(2, 128, *hw)
    print(_close(_run(conv, x, EPI_BIAS, out_nchw_f32=True), _ref(conv, x, EPI_BIAS)))


def test_x3_h_a_shapes():
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_LRELU
    torch.manual_seed(46)
    c3 = nn.Conv2d(128, 128, 3, 1, 1)
    x = torch.randn(2, 128, 8, 12)
    print(_close(_run(c3, x, EPI_LRELU), _ref(c3, x, EPI_LRELU)))
    c5 = nn.Conv2d(128, 128, 5, 2, 2)
    x = torch.randn(1, 128, 4, 4)
    print(_close(_run(c5, x, EPI_BIAS, out_f32=True), _ref(c5, x, EPI_BIAS)))


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (3, 3, 192, 320), (2, 3, 128, 448), (5, 3, 64, 192)])
def test_x3_model_ragged_sizes_match_fp32_arm(shape):
    """Whole model at sizes whose feature maps do not divide into 16 x 16 tiles (edge tiles, one-block tiles, 1x1 z maps, odd
    batch): the bf16x3 arm must agree with the fp32 CUDA-core arm at fp32 grade."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    x = H.seeded_input(shape).cuda()
    ref_model = H.seeded_model(128, 3, "calib", precision="fp32").cuda()
    model = H.seeded_model(128, 3, "calib", precision="bf16x3").cuda()
    ref, out = ref_model(x, training=False), model(x, training=False)
    r0, r1 = rd_loss(ref, x, 0.005), rd_loss(out, x, 0.005)
    yerr = float((out["y"] - ref["y"]).abs().max() / ref["y"].abs().max())
    zerr = float((out["z"] - ref["z"]).abs().max() / ref["z"].abs().max())
    real, ties = H.symbol_mismatches(out["y_in"].cpu().numpy(), ref["y_in"].cpu().numpy(), ref["y"].cpu().numpy(), H.TIE_TAU)
    realz, tiesz = H.symbol_mismatches(out["z_in"].cpu().numpy(), ref["z_in"].cpu().numpy(), ref["z"].cpu().numpy(), H.TIE_TAU)
    print(shape, f"y {yerr:.2e} z {zerr:.2e} ties {ties}+{tiesz} bpp {r1['bpp_total']:.6f}/{r0['bpp_total']:.6f} psnr {r1['psnr']:.6f}/{r0['psnr']:.6f}")
    assert yerr < 1e-4 and zerr < 1e-4 and real == 0 and realz == 0
    assert ties <= H.tie_flip_bound(out["y"].cpu().numpy(), ref["y"].cpu().numpy()) and tiesz <= H.tie_flip_bound(out["z"].cpu().numpy(), ref["z"].cpu().numpy())
    assert abs(r1["bpp_total"] - r0["bpp_total"]) < 1e-3 and abs(r1["psnr"] - r0["psnr"]) < 1e-3
    # per-element comparison off the footprints of the tie flips (tests/helpers.flip_masks) ...
    o = {k: out[k].cpu().numpy() for k in ("y_in", "z_in", "p_y", "p_z", "x_hat")}
    r = {k: ref[k].cpu().numpy() for k in ("y_in", "z_in", "p_y", "p_z", "x_hat")}
    ok_y, ok_z, ok_x = H.flip_masks(o["y_in"], r["y_in"], o["z_in"], r["z_in"], r["x_hat"].shape)
    rep = {"shape": list(shape), "y_rel_err": yerr, "z_rel_err": zerr, "ties": [ties, tiesz]}
    for name, ok in (("p_y", ok_y), ("p_z", ok_z)):
        bad, worst, n, frac = H.masked_likelihood_close(o[name], r[name], ok)
        rep[name] = {"outliers": bad, "max_abs_err": worst, "compared": n, "fraction": frac}
        assert bad <= 1e-5 * n + 1, (name, bad, n, worst)
    okx = np.broadcast_to(ok_x, r["x_hat"].shape)
    if okx.any():
        xe = float(np.abs(o["x_hat"] - r["x_hat"])[okx].max() / np.abs(r["x_hat"]).max())
        rep["x_hat"] = {"rel_err": xe, "fraction": float(okx.mean())}
        assert xe < 3e-4, xe
    # ... and, if anything flipped, on all elements against the oracle evaluated on this run's symbols
    if ties + tiesz:
        p_y, p_z, x_hat = H.oracle_given_symbols(model.state_dict(), o["y_in"], o["z_in"], 128, 3)
        for name, rr in (("p_y", p_y), ("p_z", p_z)):
            bad, worst = H.likelihood_close(o[name], rr)
            rep[name + "_given_symbols"] = {"outliers": bad, "max_abs_err": worst}
            assert bad <= 1e-5 * rr.size + 1, (name, bad, worst)
        xe2 = float(np.abs(o["x_hat"] - x_hat).max() / np.abs(x_hat).max())
        rep["x_hat_given_symbols_rel_err"] = xe2
        assert xe2 < 3e-4, xe2
    H.record_report(f"ragged/{'x'.join(map(str, shape))}", rep)


@pytest.mark.parametrize("kind", ["context", "synthesis", "hyper_synthesis"])
def test_x3_two_pass_on_integer_symbols(kind):
    """The first consumers of the quantised symbols (context conv, g_s layer 1 incl. its IGDN, h_s layer 1) skip the lo . W_hi pass
    when the hand-off kernel flags the lo half of its pair output as all zero - exact, since that term is 0.  Symbols of 256 and
    more do not split exactly: the flag must then read 1 and all three passes run."""
    from neural_image_compression_b200 import engine
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_IGDN, EPI_LRELU, Q_ROUND
    dev = torch.device("cuda:0")
    torch.manual_seed(47)
    if kind == "context":
        conv, epi, g, mask_a = nn.Conv2d(128, 256, 5, 1, 2), EPI_BIAS, None, 1
    elif kind == "synthesis":
        conv, epi, g, mask_a = nn.ConvTranspose2d(128, 128, 5, 2, 2, output_padding=1), EPI_IGDN, _gdn(inverse=True), 0
    else:
        conv, epi, g, mask_a = nn.ConvTranspose2d(128, 128, 5, 2, 2, output_padding=1), EPI_LRELU, None, 0
    mask = O.mask_a(conv.weight.detach().cpu()) if mask_a else None
    op = engine.ConvOp(conv.to(dev), epi, gdn=None if g is None else g.to(dev), mask_a=mask_a)
    for scale, want_flag in ((6.0, 0), (200.0, 1)):
        v = (scale * torch.randn(2, 16, 24, 128)).to(dev)               # NHWC, pre-rounding
        _, v_in, v_in_nhwc, _ = engine.latent_handoff(v, Q_ROUND, None, "bf16x2")
        flag = v_in_nhwc._nic_lo_flag
        torch.cuda.synchronize()
        assert int(flag) == want_flag, (scale, int(flag), float(v_in.abs().max()))
        y = op.run(v_in_nhwc, 2, 16, 24, "bf16x3", in_lo_flag=flag)
        y_full = op.run(v_in_nhwc, 2, 16, 24, "bf16x3")                   # no hint: three passes
        torch.cuda.synchronize()
        got, full = engine.from_pair(y).float().cpu().permute(0, 3, 1, 2), engine.from_pair(y_full).float().cpu().permute(0, 3, 1, 2)
        ref = _ref(conv, v_in.cpu(), epi, gdn=g, mask=mask)
        print(kind, scale, _close(got, ref), float((got - full).abs().max() / full.abs().max()))
        assert float((got - full).abs().max() / full.abs().max()) < 2e-6


@pytest.mark.parametrize("kind,precision", [("conv", "bf16x3"), ("conv_big", "bf16x3"), ("conv_192", "bf16x3"), ("convT", "bf16x3"), ("convT_big", "bf16x3"),
                                            ("conv", "bf16")])
def test_cta_pair_form_is_bit_identical_to_the_one_cta_form(kind, precision, monkeypatch):
    """conv_tc_kernel<true> (clusters of two CTAs, cta_group::2 MMAs of M = 256, half a weight slab per CTA, barriers across the pair)
    accumulates exactly what two M = 128 MMAs do: the outputs must match bit for bit, at a shape large enough for the pair form
    (>= 2 waves of two-block tiles) including ragged right / bottom edges."""
    from neural_image_compression_b200 import engine, _lib
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_LRELU
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    if kind == "conv":
        conv, h, w, b = nn.Conv2d(128, 128, 5, 2, 2).to(dev), 250, 382, 5
    elif kind == "conv_big":    # 1536 tiles = 10 full waves + 56: the last wave runs as half tiles (tail_block) in both forms
        conv, h, w, b = nn.Conv2d(128, 128, 5, 2, 2).to(dev), 256, 384, 16
    elif kind == "conv_192":    # two N tiles (192 output channels): the two CTAs of a cluster share the N tile
        conv, h, w, b = nn.Conv2d(128, 192, 5, 2, 2).to(dev), 250, 382, 5
    elif kind == "convT":
        conv, h, w, b = nn.ConvTranspose2d(128, 128, 5, 2, 2, output_padding=1).to(dev), 63, 96, 6
    else:       # enough tiles (>= 16 waves) for the phase-interleaved tile order of the pair form
        conv, h, w, b = nn.ConvTranspose2d(128, 128, 5, 2, 2, output_padding=1).to(dev), 128, 192, 16
    op = engine.ConvOp(conv, EPI_BIAS if kind.startswith("convT") else EPI_LRELU)
    x = torch.randn(b, h, w, 128, device=dev)
    x = engine.to_pair(x) if precision == "bf16x3" else x.to(torch.bfloat16)
    monkeypatch.setenv("NIC_TC_PAIR", "0")
    ref = op.run(x, b, h, w, precision).clone()
    monkeypatch.setenv("NIC_TC_PAIR", "1")
    out = op.run(x, b, h, w, precision).clone()
    torch.cuda.synchronize()
    assert _lib.load().nic_pipeline_status() == 0
    assert torch.equal(ref.view(torch.int16), out.view(torch.int16))
    if precision == "bf16x3":       # and both are the conv: torch's fp32 conv (TF32 off) of the same hi + lo input
        xin = (x[..., :128].float() + x[..., 128:].float()).permute(0, 3, 1, 2).contiguous()
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.no_grad():
                want = conv(xin)
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        want = F.leaky_relu(want, 0.01) if not kind.startswith("convT") else want
        co = conv.out_channels
        got = (out[..., :co].float() + out[..., co:].float()).permute(0, 3, 1, 2)
        assert (got - want).abs().max().item() <= 2e-4 * want.abs().max().item()


def test_phase_interleaved_tile_order_with_two_n_tiles(monkeypatch):
    """Transposed conv to 192 channels (two N tiles) on an input larger than L2 - g_s of the M = 192 models at 2048 x 1536: the
    phase-interleaved tile order (decode_tile) must give exactly what the phase-major order gives."""
    from neural_image_compression_b200 import engine, _lib
    from neural_image_compression_b200._lib import EPI_BIAS
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    conv = nn.ConvTranspose2d(192, 192, 5, 2, 2, output_padding=1).to(dev)
    op = engine.ConvOp(conv, EPI_BIAS)
    b, h, w = 1, 512, 768
    x = engine.to_pair(torch.randn(b, h, w, 192, device=dev))
    monkeypatch.setenv("NIC_TC_PHASE_ORDER", "0")
    ref = op.run(x, b, h, w, "bf16x3").clone()
    monkeypatch.setenv("NIC_TC_PHASE_ORDER", "1")
    out = op.run(x, b, h, w, "bf16x3").clone()
    torch.cuda.synchronize()
    assert _lib.load().nic_pipeline_status() == 0
    assert torch.equal(ref.view(torch.int16), out.view(torch.int16))
