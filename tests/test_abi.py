"""CPU: the C-ABI shared object loads and exports every symbol include/nic.h declares.
No compute entry is called (there is no GPU here); only the host-side shape logic."""
import ctypes as C

import pytest
import torch

from neural_image_compression_b200 import _lib
from neural_image_compression_b200._lib import ConvDesc


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    declared = _lib.header_functions()
    assert declared, "no functions parsed from include/nic.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in nic.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes signature table out of sync with nic.h"
    assert lib.nic_version() == 1
    assert lib.nic_partials_per_image() == 64


def _desc(**kw):
    d = ConvDesc()
    base = dict(n=1, c_in=128, h_in=32, w_in=48, c_out=128, h_out=16, w_out=24, kh=5, kw=5, stride=2, pad=2)
    base.update(kw)
    for k, v in base.items():
        setattr(d, k, v)
    return d


def test_tap_counts_via_packed_sizes():
    lib = _lib.load()
    assert lib.nic_packed_weight_elems(C.byref(_desc())) == 25 * 128 * 128
    # transposed 5x5 s2: 9 + 6 + 6 + 4 taps over the four output phases
    assert lib.nic_packed_weight_elems(C.byref(_desc(transposed=1, output_padding=1, h_out=64, w_out=96))) == 25 * 128 * 128
    # mask 'A' keeps 12 of 25 taps (ContextModels.py:15-16)
    assert lib.nic_packed_weight_elems(C.byref(_desc(stride=1, h_out=32, w_out=48, mask_a=1, c_out=256))) == 12 * 128 * 256
    assert lib.nic_packed_weight_elems(C.byref(_desc(kh=1, kw=1, pad=0, stride=1, h_out=32, w_out=48, c_in=512, c_out=640))) == 512 * 640


def test_bad_descriptor_is_an_error_not_a_crash():
    lib = _lib.load()
    assert lib.nic_packed_weight_elems(C.byref(_desc(h_out=17))) == 0
    assert b"does not match" in lib.nic_last_error()
    assert lib.nic_packed_weight_elems(C.byref(_desc(stride=3))) == 0


def test_product_path_fails_loudly_without_a_gpu():
    from neural_image_compression_b200.Models import JointAutoregressiveHierarchical
    from neural_image_compression_b200.RateDistortionLoss import rd_loss
    model = JointAutoregressiveHierarchical(128, K=1)
    with pytest.raises(_lib.NicError):
        model(torch.rand(1, 3, 64, 64), training=False)
    with pytest.raises(_lib.NicError):
        model.encoder(torch.rand(1, 3, 64, 64))
    with pytest.raises(_lib.NicError):
        rd_loss({"x_hat": torch.rand(1, 3, 64, 64), "logp_y": torch.zeros(1, 1), "logp_z": torch.zeros(1, 1)},
                torch.rand(1, 3, 64, 64), 0.005)


def test_training_path_fails_loudly_without_a_gpu():
    """model(x) with autograd on (the reference trainer's call) and the optimizer have no CPU path either."""
    import pytest
    import torch
    from neural_image_compression_b200 import _lib
    from neural_image_compression_b200.Models import JointAutoregressiveHierarchical
    from neural_image_compression_b200.training import Adam, step_gradients
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    model = JointAutoregressiveHierarchical(128, K=1)
    x = torch.rand(1, 3, 64, 64)
    with pytest.raises(_lib.NicError):
        model(x)                                   # training=True, grad enabled -> the differentiable path
    with pytest.raises(_lib.NicError):
        step_gradients(model, x, 0.005)
    opt = Adam(model.parameters(), lr=1e-4)
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(_lib.NicError):
        opt.step()
