"""``gaussian_cdf`` of /root/reference/utils.py:6-8 (the plotting helpers of that file are out of scope)."""
import math

import torch


def gaussian_cdf(x: torch.Tensor):
    """Standard normal CDF via erf, the reference's (non tail-stable) form: 0.5 * (1 + erf(x / sqrt(2)))."""
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
