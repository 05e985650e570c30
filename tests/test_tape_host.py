"""CPU (no GPU): the host logic of the residual family's training step - ``training.Tape`` and the ``tape=`` plumbing of
Layers.py / Components._Chain - with torch stand-ins for the C-ABI calls (conv forward / weight gradient / adjoint conv, GDN forward /
backward, LeakyReLU backward, add).  What is checked is the BOOKKEEPING: every op of the 3x3 transforms is recorded once, gradients of
tensors with two consumers (the block input feeding the main path and the skip) are summed, the gradient of the transform's input
comes back, and every parameter gets its gradient - against torch autograd over the oracle's restatement of the same transforms
(oracle/forward.py: analysis_3x3 / synthesis_3x3 / hyper_*_3x3, pinned to the reference's own classes by tests/test_oracle.py).
The kernels themselves are checked on the GPU (tests/test_gpu_residual.py)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import forward as O
from oracle.gdn import gdn_effective

NCHW, NHWC, EPI_LRELU = 0, 1, 1


def _nchw(t, layout):
    return t if layout == NCHW else t.permute(0, 3, 1, 2)


def _fconv(conv, x, w, b):
    if isinstance(conv, nn.ConvTranspose2d):
        return F.conv_transpose2d(x, w, b, stride=conv.stride, padding=conv.padding, output_padding=conv.output_padding)
    return F.conv2d(x, w, b, stride=conv.stride, padding=conv.padding)


def _gdn_fn(gdn, x, beta, gamma):
    be, ga = gdn_effective(beta, gamma, gdn.beta_min)
    c = be.numel()
    norm = F.conv2d(x * x, ga.reshape(c, c, 1, 1), be)
    return x * (torch.sqrt(norm) if gdn.inverse else torch.rsqrt(norm))


@pytest.fixture
def torch_backend(monkeypatch):
    from neural_image_compression_b200 import training as T

    def conv_forward(arm, conv, epi, x, n, h, w, in_layout=NHWC, out_layout=NHWC, out=None, out_c_total=0, out_c_offset=0, mask_a=False,
                     pair_out=False):
        with torch.no_grad():
            y = _fconv(conv, _nchw(x, in_layout).double(), conv.weight.double(), conv.bias.double())
            if epi == EPI_LRELU:
                y = F.leaky_relu(y, 0.01)
            y = y if out_layout == NCHW else y.permute(0, 2, 3, 1)
            if out is not None:
                out[..., out_c_offset:out_c_offset + y.shape[-1]] = y
                return out
            return y.contiguous()

    def conv_wgrad(conv, x, g, n, h, w, in_layout, out_layout, arm="fp32", pairs=None):
        wt, b = conv.weight.detach().double().requires_grad_(), conv.bias.detach().double().requires_grad_()
        with torch.enable_grad():
            y = _fconv(conv, _nchw(x, in_layout).double(), wt, b)
            go = _nchw(g, out_layout).double()
            if go.shape[1] != y.shape[1]:                      # a channel window of a wider buffer (psi inside `combined`)
                go = go[:, -y.shape[1]:]
            return torch.autograd.grad(y, (wt, b), go)

    def conv_dgrad(conv, g, n, h_in, w_in, g_layout=NHWC, weight=None, c_in=None, arm="fp32"):
        x = torch.zeros(n, conv.in_channels, h_in, w_in, dtype=torch.float64, requires_grad=True)
        with torch.enable_grad():
            y = _fconv(conv, x, conv.weight.detach().double(), conv.bias.detach().double())
            return torch.autograd.grad(y, x, _nchw(g, g_layout).double())[0].permute(0, 2, 3, 1).contiguous()

    def gdn_forward(arm, gdn, u, n, h, w):
        with torch.no_grad():
            return _gdn_fn(gdn, u.permute(0, 3, 1, 2), gdn.beta.double(), gdn.gamma.double()).permute(0, 2, 3, 1).contiguous(), None

    def gdn_bwd(gdn, u, g, n, h, w, norm=None, side=None, keep=None):
        x = u.permute(0, 3, 1, 2).detach().requires_grad_()
        beta, gamma = gdn.beta.detach().double().requires_grad_(), gdn.gamma.detach().double().requires_grad_()
        with torch.enable_grad():
            y = _gdn_fn(gdn, x, beta, gamma)
            dx, db, dg = torch.autograd.grad(y, (x, beta, gamma), g.permute(0, 3, 1, 2))
        return dx.permute(0, 2, 3, 1).contiguous(), db, dg

    def lrelu_bwd_(g, out):
        return g.mul_(torch.where(out > 0, torch.ones_like(g), torch.full_like(g, 0.01)))

    def add_(dst, src):
        return dst.add_(src)
    for name, fn in dict(conv_forward=conv_forward, conv_wgrad=conv_wgrad, conv_dgrad=conv_dgrad, gdn_forward=gdn_forward, gdn_bwd=gdn_bwd,
                         lrelu_bwd_=lrelu_bwd_, add_=add_).items():
        monkeypatch.setattr(T, name, fn)
    return T


def _perturb(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in mod.modules():
            if type(m).__name__ == "GDN":
                m.gamma.add_(0.02 * torch.rand(m.gamma.shape, generator=g))
    return mod


@pytest.mark.parametrize("which", ["encoder", "decoder", "hyper_encoder", "hyper_decoder"])
def test_tape_routes_gradients_through_the_3x3_transforms(torch_backend, which):
    from neural_image_compression_b200 import Components as Cm
    T = torch_backend
    M = 8
    torch.manual_seed(3)
    mod = _perturb({"encoder": Cm.Encoder3x3, "decoder": Cm.Decoder3x3, "hyper_encoder": Cm.HyperEncoder3x3,
                    "hyper_decoder": Cm.HyperDecoder3x3}[which](latent_channels=M), 4)
    n, h, w = 2, (32 if which == "encoder" else 4), (48 if which == "encoder" else 6)
    x_nchw = (torch.rand(n, 3, h, w) if which == "encoder" else torch.randn(n, M, h, w)).double()
    # oracle: autograd over the restated transform
    sd = {f"{which}.{k}": v.detach().double().clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(("pedestal", "bound")))
          for k, v in mod.state_dict().items()}
    xr = x_nchw.clone().requires_grad_()
    prev, O.DIFFERENTIABLE = O.DIFFERENTIABLE, True
    try:
        fn = {"encoder": O.analysis_3x3, "decoder": O.synthesis_3x3, "hyper_encoder": O.hyper_analysis_3x3,
              "hyper_decoder": O.hyper_synthesis_3x3}[which]
        y_ref = fn(sd, xr, torch.float64)
        g_out = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
        y_ref.backward(g_out)
    finally:
        O.DIFFERENTIABLE = prev
    # the package's chain with a tape
    tape = T.Tape("fp32", n)
    if which == "encoder":
        x_in, in_layout = x_nchw, NCHW                                     # the image arrives NCHW and needs no gradient
    else:
        x_in, in_layout = x_nchw.permute(0, 2, 3, 1).contiguous(), NHWC
    y, ho, wo = mod.run_nhwc(x_in, n, h, w, "fp32", in_layout=in_layout, tape=tape)
    rgb = which == "decoder"                                               # the RGB layer stays NCHW (x_hat)
    y_nchw = y if rgb else y.permute(0, 3, 1, 2)
    assert torch.allclose(y_nchw, y_ref.detach(), rtol=1e-10, atol=1e-12)
    n_convs = sum(isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) for m in mod.modules())
    assert sum(r[0] == "conv" for r in tape.recs) == n_convs and tape.output is y
    grads = {}

    def put(p, g):
        grads[id(p)] = g if id(p) not in grads else grads[id(p)] + g
    g_in = tape.backward(g_out if rgb else g_out.permute(0, 2, 3, 1).contiguous(), NCHW if rgb else NHWC, put, T.conv_wgrad,
                         root=None if which == "encoder" else x_in)
    for k, p in mod.named_parameters():
        ref = sd[f"{which}.{k}"].grad
        assert id(p) in grads, k
        assert torch.allclose(grads[id(p)], ref, rtol=1e-8, atol=1e-10 * float(ref.abs().max() + 1)), k
    if which == "encoder":
        assert g_in is None
    else:
        assert torch.allclose(g_in.permute(0, 3, 1, 2), xr.grad, rtol=1e-8, atol=1e-12)


def test_chain_writes_its_last_conv_into_a_channel_window(torch_backend):
    """h_s of the training forward writes psi straight into its window of the entropy-parameter stack's concat buffer
    (`final_kw`): same values as the plain call, the other window untouched, and the backward starts from the window's gradient."""
    from neural_image_compression_b200 import Components as Cm
    T = torch_backend
    M, n, h, w = 8, 1, 2, 3
    torch.manual_seed(7)
    mod = Cm.HyperDecoder3x3(latent_channels=M)
    x = torch.randn(n, h, w, M).double()
    plain, ho, wo = mod.run_nhwc(x, n, h, w, "fp32", tape=T.Tape("fp32", n))
    combined = torch.full((n, ho, wo, 4 * M), 7.0, dtype=torch.float64)
    tape = T.Tape("fp32", n)
    mod.run_nhwc(x, n, h, w, "fp32", tape=tape, final_kw=dict(out=combined, out_c_total=4 * M, out_c_offset=2 * M))
    assert tape.output is combined and torch.equal(combined[..., 2 * M:], plain) and bool((combined[..., :2 * M] == 7.0).all())
    grads = {}
    g = torch.randn(n, ho, wo, 2 * M, dtype=torch.float64)
    gx = tape.backward(g, NHWC, lambda p, v: grads.__setitem__(id(p), v), T.conv_wgrad, root=x)
    assert gx.shape == x.shape and len(grads) == len(list(mod.parameters()))
