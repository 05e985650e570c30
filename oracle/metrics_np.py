"""A SECOND, independent restatement of MS-SSIM (TEST INFRASTRUCTURE, see oracle/__init__.py) - float64 numpy / scipy only.

``pytorch_msssim`` (Evaluator.py:7, 38, 45; named without a version in /root/reference/requirements.txt) is absent from
/root/reference and from this image and cannot be installed offline, so neither the kernel nor oracle/metrics.py can be pinned to
a run of the package: PARITY UNPINNED at that boundary.  What can be done is to state the published algorithm twice, by different
routes, and require the two statements (and the GPU kernel) to agree:

  oracle/metrics.py   torch, float32, separable depth-wise F.conv2d, F.avg_pool2d - shaped like the package's own code;
  this file           the multi-scale structural similarity of Wang, Simoncelli & Bovik (2003) written from the formulas:
                      one dense 11 x 11 Gaussian window (outer product), `scipy.signal.correlate2d(mode="valid")` per plane,
                      explicit 2 x 2 block means with zero padding on odd sizes for the dyadic pyramid, float64 throughout.

Conventions taken from the package's documentation (the parts a formula-only reading leaves open): sigma = 1.5, K = (0.01, 0.03),
weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333), the contrast-structure terms and the last scale's SSIM clipped at 0 before the
weighted product, `avg_pool2d(kernel 2, padding size % 2)` = zero-padded block mean that divides by 4 (count_include_pad).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import correlate2d

WEIGHTS = np.array([0.0448, 0.2856, 0.3001, 0.2363, 0.1333])


def window(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    r = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-0.5 * (r / sigma) ** 2)
    g /= g.sum()
    return np.outer(g, g)


def _local(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    return correlate2d(a, w, mode="valid")


def ssim_and_cs(x: np.ndarray, y: np.ndarray, data_range: float = 1.0):
    """Mean SSIM and mean contrast-structure term of one plane pair."""
    w = window()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mx, my = _local(x, w), _local(y, w)
    vx, vy, cxy = _local(x * x, w) - mx * mx, _local(y * y, w) - my * my, _local(x * y, w) - mx * my
    cs = (2.0 * cxy + c2) / (vx + vy + c2)
    lum = (2.0 * mx * my + c1) / (mx * mx + my * my + c1)
    return float((lum * cs).mean()), float(cs.mean())


def halve(a: np.ndarray) -> np.ndarray:
    """2 x 2 block mean; an odd side is zero-padded by one row / column on BOTH ends first (avg_pool2d(2, padding=1) semantics:
    windows start at -1, padded zeros count in the divisor)."""
    ph, pw = a.shape[0] % 2, a.shape[1] % 2
    if ph or pw:
        a = np.pad(a, ((ph, ph), (pw, pw)))
    h2, w2 = a.shape[0] // 2, a.shape[1] // 2
    return a[:2 * h2, :2 * w2].reshape(h2, 2, w2, 2).mean(axis=(1, 3))


def ms_ssim_plane(x: np.ndarray, y: np.ndarray, data_range: float = 1.0) -> float:
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    if min(x.shape) <= (11 - 1) * 2 ** 4:
        raise ValueError("image side must exceed 160 pixels for five scales of an 11-tap window")
    terms = []
    for level in range(5):
        s, cs = ssim_and_cs(x, y, data_range)
        terms.append(max(cs, 0.0) if level < 4 else max(s, 0.0))
        if level < 4:
            x, y = halve(x), halve(y)
    return float(np.prod(np.array(terms) ** WEIGHTS))


def ms_ssim(x: np.ndarray, y: np.ndarray, data_range: float = 1.0) -> float:
    """[B, C, H, W] arrays -> mean over planes (the package's size_average=True)."""
    x, y = np.asarray(x), np.asarray(y)
    return float(np.mean([ms_ssim_plane(x[b, c], y[b, c], data_range) for b in range(x.shape[0]) for c in range(x.shape[1])]))
