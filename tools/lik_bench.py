"""Timing of gm_likelihood_kernel<3, full> alone (CUDA events, L2 flushed between launches): the (parts, B) grid against the flat
balanced form, at batch 16 and 256.  python tools/lik_bench.py"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_image_compression_b200.EntropyModels import gm_likelihood  # noqa: E402
from neural_image_compression_b200._lib import Q_ROUND  # noqa: E402

dev = torch.device("cuda:0")
M, K, H, W = 128, 3, 32, 48
PEAK = 6553.6
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run(lb, env, clean=False):
    for k in ("NIC_LIK_FLAT", "NIC_LIK_GRID", "NIC_LIK_PARTS", "NIC_LIK_STAGED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    y = 5 * torch.randn((lb, M, H, W), device=dev)
    raw = torch.randn((lb, 3 * K * M, H, W), device=dev)
    d = []
    for i in range(12):
        flush.zero_()
        if clean:
            flush.view(torch.int32).max()       # read pass: the L2 now holds CLEAN lines of the flush buffer (nothing to write back)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gm_likelihood(y, raw, M, K, Q_ROUND, full=True); b.record()
        torch.cuda.synchronize()
        if i >= 4:
            d.append(a.elapsed_time(b))
    ms = statistics.median(d)
    return ms, y.numel() * 88 / (ms / 1e3) / 1e9 / PEAK


for env in ({"NIC_LIK_FLAT": "0"}, {"NIC_LIK_STAGED": "0"}, {}, {"NIC_LIK_GRID": "296"}, {"NIC_LIK_GRID": "888"}):
    r16, r256, c16 = run(16, env), run(256, env), run(16, env, clean=True)
    print(env, f"batch16 {r16[0] * 1e3:.1f} us {r16[1]:.3f} | batch256 {r256[0] * 1e3:.1f} us {r256[1]:.3f} | batch16 after a clean flush "
          f"{c16[0] * 1e3:.1f} us {c16[1]:.3f}", flush=True)
