"""GPU: the tcgen05 / TMEM / TMA arm (precision "bf16") layer by layer against a float64 CPU convolution of the SAME
bf16-rounded operands (so only accumulation order and the output rounding differ), then the whole model in bf16 mode
against the fp32 oracle with the looser tolerances that bf16 operands imply (reported, not hidden)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import forward as O
from oracle.gdn import gdn_effective
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float64)


def _run_layer(conv, x_nchw, epilogue, gdn=None, mask_a=False, out_layout_nchw=False, out_f32=False, out_c_total=0, out_c_offset=0):
    from neural_image_compression_b200 import engine
    from neural_image_compression_b200._lib import LAYOUT_NCHW, LAYOUT_NHWC
    dev = torch.device("cuda:0")
    conv = conv.to(dev)
    if gdn is not None:
        gdn = gdn.to(dev)
    op = engine.ConvOp(conv, epilogue, gdn=gdn, mask_a=mask_a)
    n, c, h, w = x_nchw.shape
    x = x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
    out = None
    if out_c_total:
        ho, wo = engine.conv_out_hw(conv, h, w)
        out = torch.zeros((n, ho, wo, out_c_total), dtype=torch.bfloat16, device=dev)
    y = op.run(x, n, h, w, "bf16", out_layout=LAYOUT_NCHW if out_layout_nchw else LAYOUT_NHWC, out=out,
               out_c_total=out_c_total, out_c_offset=out_c_offset, out_dtype=torch.float32 if out_f32 else torch.bfloat16)
    torch.cuda.synchronize()
    y = y.float().cpu()
    if not out_layout_nchw:
        y = y.permute(0, 3, 1, 2)
    return y


def _ref_layer(conv, x, epilogue, gdn=None, mask=None):
    from neural_image_compression_b200._lib import EPI_GDN, EPI_IGDN, EPI_LRELU
    w = _bf(conv.weight.detach().cpu())
    if mask is not None:
        w = w * mask.double()
    b = conv.bias.detach().cpu().double()
    xin = _bf(x)
    if isinstance(conv, nn.ConvTranspose2d):
        v = F.conv_transpose2d(xin, w, b, stride=conv.stride, padding=conv.padding, output_padding=conv.output_padding)
    else:
        v = F.conv2d(xin, w, b, stride=conv.stride, padding=conv.padding)
    if epilogue == EPI_LRELU:
        v = F.leaky_relu(v, 0.01)
    elif epilogue in (EPI_GDN, EPI_IGDN):
        beta, gamma = gdn_effective(gdn.beta.detach().cpu(), gdn.gamma.detach().cpu())
        C = beta.numel()
        norm = F.conv2d(_bf(v * v), _bf(gamma).reshape(C, C, 1, 1), beta.double())
        v = v * (torch.sqrt(norm) if epilogue == EPI_IGDN else torch.rsqrt(norm))
    return v.float()


def _close(y, ref, tol):
    err = float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-9))
    assert y.shape == ref.shape and err < tol, err
    return err


@pytest.mark.parametrize("hw", [(32, 48), (20, 28), (8, 12)])
def test_tc_conv5x5_s2(hw):
    from neural_image_compression_b200._lib import EPI_BIAS
    torch.manual_seed(20)
    conv = nn.Conv2d(128, 128, 5, 2, 2)
    x = torch.randn(2, 128, *hw)
    y = _run_layer(conv, x, EPI_BIAS, out_f32=True)
    _close(y, _ref_layer(conv, x, EPI_BIAS), 2e-3)


@pytest.mark.parametrize("inverse", [False, True])
def test_tc_conv_with_fused_gdn(inverse):
    from neural_image_compression_b200._lib import EPI_GDN, EPI_IGDN
    from neural_image_compression_b200.gdn import GDN
    torch.manual_seed(21)
    g = GDN(128, inverse=inverse)
    with torch.no_grad():
        g.gamma.add_(0.02 * torch.rand_like(g.gamma)); g.beta.add_(0.1 * torch.rand_like(g.beta))
    epi = EPI_IGDN if inverse else EPI_GDN
    conv = nn.ConvTranspose2d(128, 128, 5, 2, 2, output_padding=1) if inverse else nn.Conv2d(128, 128, 5, 2, 2)
    x = torch.randn(2, 128, 16, 24)
    y = _run_layer(conv, x, epi, gdn=g)
    _close(y, _ref_layer(conv, x, epi, gdn=g), 1.5e-2)


def test_tc_transposed_to_192_lrelu_and_3x3_into_channel_window():
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_LRELU
    torch.manual_seed(22)
    conv = nn.ConvTranspose2d(128, 192, 5, 2, 2, output_padding=1)
    x = torch.randn(2, 128, 8, 12)
    _close(_run_layer(conv, x, EPI_LRELU), _ref_layer(conv, x, EPI_LRELU), 1e-2)
    conv3 = nn.Conv2d(192, 256, 3, 1, 1)
    x3 = torch.randn(2, 192, 16, 24)
    y = _run_layer(conv3, x3, EPI_BIAS, out_c_total=512, out_c_offset=256)
    assert float(y[:, :256].abs().max()) == 0                       # the other half of the concat buffer is untouched
    _close(y[:, 256:], _ref_layer(conv3, x3, EPI_BIAS), 1e-2)


def test_tc_masked_context_conv():
    from neural_image_compression_b200._lib import EPI_BIAS
    torch.manual_seed(23)
    conv = nn.Conv2d(128, 256, 5, 1, 2)
    x = torch.round(4 * torch.randn(2, 128, 16, 24))
    mask = O.mask_a(conv.weight.detach())
    _close(_run_layer(conv, x, EPI_BIAS, mask_a=True), _ref_layer(conv, x, EPI_BIAS, mask=mask), 1e-2)


def test_tc_pointwise_stack_shapes():
    from neural_image_compression_b200._lib import EPI_BIAS, EPI_LRELU
    torch.manual_seed(24)
    c1 = nn.Conv2d(512, 640, 1)
    x = torch.randn(2, 512, 8, 12)
    _close(_run_layer(c1, x, EPI_LRELU), _ref_layer(c1, x, EPI_LRELU), 1e-2)
    c3 = nn.Conv2d(640, 1152, 1)
    x3 = torch.randn(2, 640, 8, 12)
    _close(_run_layer(c3, x3, EPI_BIAS, out_layout_nchw=True, out_f32=True), _ref_layer(c3, x3, EPI_BIAS), 2e-3)


def test_tc_last_layer_to_rgb_nchw():
    from neural_image_compression_b200._lib import EPI_BIAS
    torch.manual_seed(25)
    conv = nn.ConvTranspose2d(128, 3, 5, 2, 2, output_padding=1)
    x = torch.randn(2, 128, 16, 24)
    _close(_run_layer(conv, x, EPI_BIAS, out_layout_nchw=True, out_f32=True), _ref_layer(conv, x, EPI_BIAS), 2e-3)


def test_tc_first_layer_from_nchw_image():
    from neural_image_compression_b200 import engine
    from neural_image_compression_b200._lib import EPI_GDN, LAYOUT_NCHW
    from neural_image_compression_b200.gdn import GDN
    torch.manual_seed(26)
    dev = torch.device("cuda:0")
    conv, g = nn.Conv2d(3, 128, 5, 2, 2), GDN(128)
    x = torch.rand(2, 3, 64, 96)
    op = engine.ConvOp(conv.to(dev), EPI_GDN, gdn=g.to(dev))
    y = op.run(x.to(dev), 2, 64, 96, "bf16", in_layout=LAYOUT_NCHW).float().cpu().permute(0, 3, 1, 2)
    v = F.conv2d(x.double(), conv.weight.detach().cpu().double(), conv.bias.detach().cpu().double(), stride=2, padding=2)
    beta, gamma = gdn_effective(g.beta.detach().cpu(), g.gamma.detach().cpu())
    ref = (v * torch.rsqrt(F.conv2d(v * v, gamma.double().reshape(128, 128, 1, 1), beta.double()))).float()
    _close(y, ref, 1.5e-2)


@pytest.mark.parametrize("init", ["calib"])
def test_model_bf16_against_oracle(init):
    """bf16 operands move y by ~1e-2 relative: symbols within half a step of a rounding boundary flip (reported), and the
    rate / distortion follow.  Bounds here are what bf16 arithmetic can meet, NOT the fp32-grade parity criterion
    (that is tests/test_gpu_model.py on precision="fp32")."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = H.seeded_model(128, 3, init, precision="bf16")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = H.seeded_input((2, 3, 256, 384))
    ref = O.forward(sd, x, 128, 3)
    ref_rd = O.rd_loss(ref, x, 0.005)
    model = model.cuda()
    out = model(x.cuda(), training=False)
    rd = rd_loss(out, x.cuda(), 0.005)
    flips = float((out["y_in"].cpu() != ref["y_in"]).float().mean())
    yerr = float((out["y"].cpu() - ref["y"]).abs().max() / ref["y"].abs().max())
    print(f"bf16 {init}: y rel err {yerr:.3e}, symbol flips {flips:.4f}, bpp {rd['bpp_total']:.5f} vs {ref_rd['bpp_total']:.5f}, "
          f"psnr {rd['psnr']:.5f} vs {ref_rd['psnr']:.5f}")
    assert yerr < 3e-2 and flips < 0.08
    assert abs(rd["bpp_total"] - ref_rd["bpp_total"]) < 0.05 and abs(rd["psnr"] - ref_rd["psnr"]) < 0.05
    for k in ("p_y", "p_z"):
        assert float(out[k].min()) >= float(np.float32(1e-9)) and float(out[k].max()) <= 1 + 1e-6


def test_model_bf16_full_size_batch_is_batch_independent():
    model = H.seeded_model(128, 3, "calib", precision="bf16").cuda()
    x = H.seeded_input((16, 3, 512, 768)).cuda()
    out = model(x, training=False)
    one = model(x[3:4], training=False)
    for k in ("y_in", "z_in", "p_y", "x_hat"):
        assert torch.equal(one[k][0], out[k][3]), k


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (3, 3, 192, 320), (2, 3, 128, 448)])
def test_model_bf16_ragged_sizes_match_fp32_arm(shape):
    """Sizes whose feature maps do not divide into 16 x 16 tiles (edge tiles, one-block tiles, 1x1 z maps): the
    tensor-core arm must agree with the fp32 CUDA-core arm (itself checked against the oracle) at bf16 tolerances."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    x = H.seeded_input(shape).cuda()
    ref_model = H.seeded_model(128, 3, "calib", precision="fp32").cuda()
    model = H.seeded_model(128, 3, "calib", precision="bf16").cuda()
    ref, out = ref_model(x, training=False), model(x, training=False)
    r0, r1 = rd_loss(ref, x, 0.005), rd_loss(out, x, 0.005)
    yerr = float((out["y"] - ref["y"]).abs().max() / ref["y"].abs().max())
    # x_hat follows the (few) flipped symbols locally, so the max error is a flip's footprint; the RMS is the bf16 noise
    xerr = float((out["x_hat"] - ref["x_hat"]).pow(2).mean().sqrt() / ref["x_hat"].pow(2).mean().sqrt())
    xmax = float((out["x_hat"] - ref["x_hat"]).abs().max() / ref["x_hat"].abs().max())
    flips = float((out["y_in"] != ref["y_in"]).float().mean())
    print(shape, f"y {yerr:.2e} x_hat {xerr:.2e} flips {flips:.4f} bpp {r1['bpp_total']:.5f}/{r0['bpp_total']:.5f}")
    assert yerr < 3e-2 and xerr < 5e-2 and xmax < 0.3 and flips < 0.08
    assert abs(r1["bpp_total"] - r0["bpp_total"]) < 0.05 and abs(r1["psnr"] - r0["psnr"]) < 0.05


def test_model_bf16_training_forward_and_k1():
    """training=True (injected noise) and the K = 1 mean-scale variant through the tensor-core arm."""
    torch.manual_seed(30)
    x = H.seeded_input((2, 3, 128, 192))
    for K in (1, 3):
        model = H.seeded_model(128, K, "calib", precision="bf16")
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        nz, ny = torch.rand(2, 128, 2, 3) - 0.5, torch.rand(2, 128, 8, 12) - 0.5
        ref = O.forward(sd, x, 128, K, training=True, noise_z=nz, noise_y=ny)
        with torch.no_grad():                      # forward-only path of this arm (with autograd on, training=True is the fp32 train step)
            out = model.cuda()(x.cuda(), training=True, noise=(nz.cuda(), ny.cuda()))
        assert out["training"] is True and set(out) >= ({"mu", "sigma"} if K == 1 else {"weights", "mus", "sigmas"})
        assert float((out["y_in"].cpu() - ref["y_in"]).abs().max() / ref["y_in"].abs().max()) < 3e-2
        bits = -out["logp_y"].double().sum().item() / np.log(2)
        ref_bits = -ref["logp_y"].double().sum().item() / np.log(2)
        assert abs(bits - ref_bits) / ref_bits < 2e-2, (K, bits, ref_bits)


def test_model_mixed_precision_symbols_match_fp32_arm_exactly():
    """precision="mixed": g_a / h_a in fp32 (same kernels as the parity arm) -> y, z and every symbol are bit-identical to
    the fp32 arm; the bf16 entropy path and g_s then only move likelihoods / x_hat at bf16 tolerances."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    x = H.seeded_input((2, 3, 256, 384)).cuda()
    ref_model = H.seeded_model(128, 3, "calib", precision="fp32").cuda()
    model = H.seeded_model(128, 3, "calib", precision="mixed").cuda()
    ref, out = ref_model(x, training=False), model(x, training=False)
    for k in ("y", "z", "y_in", "z_in"):
        assert torch.equal(out[k], ref[k]), k
    r0, r1 = rd_loss(ref, x, 0.005), rd_loss(out, x, 0.005)
    rel = ((out["p_y"] - ref["p_y"]).abs() / ref["p_y"]).max().item()
    print(f"mixed: bpp {r1['bpp_total']:.6f} vs {r0['bpp_total']:.6f}, psnr {r1['psnr']:.6f} vs {r0['psnr']:.6f}, max rel p_y err {rel:.2e}")
    assert abs(r1["bpp_total"] - r0["bpp_total"]) < 1e-3 and abs(r1["psnr"] - r0["psnr"]) < 1e-3


def test_model_bf16x3_symbols_match_fp32_arm_up_to_ties():
    """precision="bf16x3": g_a / h_a on the tensor cores with hi/lo-split operands.  y must agree with the fp32 arm to
    ~1e-5 relative and the symbols must be equal except where the fp32 arm's y sits within 2e-3 of a rounding tie."""
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    for init in ("calib", "gain"):
        x = H.seeded_input((2, 3, 256, 384)).cuda()
        ref_model = H.seeded_model(128, 3, init, precision="fp32").cuda()
        model = H.seeded_model(128, 3, init, precision="bf16x3").cuda()
        ref, out = ref_model(x, training=False), model(x, training=False)
        yerr = float((out["y"] - ref["y"]).abs().max() / ref["y"].abs().max())
        zerr = float((out["z"] - ref["z"]).abs().max() / ref["z"].abs().max())
        real, ties = H.symbol_mismatches(out["y_in"].cpu().numpy(), ref["y_in"].cpu().numpy(), ref["y"].cpu().numpy(), 2e-3)
        realz, tiesz = H.symbol_mismatches(out["z_in"].cpu().numpy(), ref["z_in"].cpu().numpy(), ref["z"].cpu().numpy(), 2e-3)
        r0, r1 = rd_loss(ref, x, 0.005), rd_loss(out, x, 0.005)
        print(f"bf16x3 {init}: y rel err {yerr:.2e}, z rel err {zerr:.2e}, symbol flips at ties {ties}+{tiesz} (elsewhere {real}+{realz}), "
              f"bpp {r1['bpp_total']:.6f} vs {r0['bpp_total']:.6f}, psnr {r1['psnr']:.6f} vs {r0['psnr']:.6f}")
        assert yerr < 1e-4 and zerr < 1e-4 and real == 0 and realz == 0
        if init == "calib":
            assert abs(r1["bpp_total"] - r0["bpp_total"]) < 1e-3 and abs(r1["psnr"] - r0["psnr"]) < 1e-3
