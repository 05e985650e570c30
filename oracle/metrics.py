"""CPU restatement of the evaluator's distortion metrics (TEST INFRASTRUCTURE, see oracle/__init__.py).

``compute_metrics`` follows /root/reference/Evaluator.py:26-53.  ``ms_ssim`` restates ``pytorch_msssim.ms_ssim`` (PyPI package
``pytorch_msssim``, named without a version in /root/reference/requirements.txt; imported at Evaluator.py:7; absent from
/root/reference and from this image - PARITY UNPINNED at this boundary): the algorithm its 0.2.x releases publish - 11-tap
Gaussian window (sigma 1.5) applied separably with VALID convolution per channel, K = (0.01, 0.03), five scales with
``F.avg_pool2d(kernel_size=2, padding=size % 2)`` between them, per-channel product of relu(cs)^w over the first four scales and
relu(ssim)^w of the last, weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333), mean over channels and batch.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _gauss(size=11, sigma=1.5, dtype=torch.float32):
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x, win):
    c = x.shape[1]
    k = win.to(x.dtype)
    x = F.conv2d(x, k.view(1, 1, -1, 1).repeat(c, 1, 1, 1), groups=c)        # along H
    return F.conv2d(x, k.view(1, 1, 1, -1).repeat(c, 1, 1, 1), groups=c)     # along W


def _ssim(x, y, win, data_range, K=(0.01, 0.03)):
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter(x, win), _filter(y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1, s2, s12 = _filter(x * x, win) - mu1_sq, _filter(y * y, win) - mu2_sq, _filter(x * y, win) - mu12
    cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim_map = ((2 * mu12 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def ms_ssim(x, y, data_range=1.0, size_average=True):
    assert min(x.shape[-2:]) > (11 - 1) * 2 ** 4
    win = _gauss(dtype=x.dtype)
    mcs = []
    for i in range(5):
        ssim_pc, cs = _ssim(x, y, win, data_range)
        if i < 4:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in x.shape[2:]]
            x, y = F.avg_pool2d(x, kernel_size=2, padding=pad), F.avg_pool2d(y, kernel_size=2, padding=pad)
    stack = torch.stack(mcs + [torch.relu(ssim_pc)], dim=0)
    val = torch.prod(stack ** torch.tensor(WEIGHTS, dtype=x.dtype).view(-1, 1, 1), dim=0)
    return val.mean() if size_average else val.mean(1)


def rgb_to_luma(x):
    return 0.299 * x[:, 0] + 0.587 * x[:, 1] + 0.114 * x[:, 2]


def compute_metrics(orig, recon):
    """Evaluator.py:32-53 on already-clamped `recon`."""
    mse_rgb = torch.mean((orig - recon) ** 2).item()
    y_o, y_r = rgb_to_luma(orig).unsqueeze(1), rgb_to_luma(recon).unsqueeze(1)
    mse_y = torch.mean((y_o - y_r) ** 2).item()
    return {"MSE(255)": mse_rgb * 255 ** 2, "PSNR(RGB)": 10 * np.log10(1.0 / mse_rgb) if mse_rgb > 0 else float("inf"),
            "MS-SSIM(RGB)": ms_ssim(recon, orig).item(), "PSNR(Y)": 10 * np.log10(1.0 / mse_y) if mse_y > 0 else float("inf"),
            "MS-SSIM(Y)": ms_ssim(y_r, y_o).item()}
