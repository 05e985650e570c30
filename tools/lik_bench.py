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
    for k in ("NIC_LIK_FLAT", "NIC_LIK_GRID", "NIC_LIK_PARTS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    y = 5 * torch.randn((lb, M, H, W), device=dev)
    raw = torch.randn((lb, 3 * K * M, H, W), device=dev)
    d = []
    out = gm_likelihood(y, raw, M, K, Q_ROUND, full=True)
    for i in range(24):
        flush.zero_()
        if clean:
            flush.view(torch.int32).max()       # read pass: the L2 now holds CLEAN lines of the flush buffer (nothing to write back)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gm_likelihood(y, raw, M, K, Q_ROUND, full=True, out=out); b.record()
        torch.cuda.synchronize()
        if i >= 4:
            d.append(a.elapsed_time(b))
    ms = statistics.mean(d)
    return ms, y.numel() * 88 / (ms / 1e3) / 1e9 / PEAK


envs = ({"NIC_LIK_FLAT": "0"}, {}, {"NIC_LIK_GRID": "296"}, {"NIC_LIK_GRID": "888"})
acc = {i: ([], []) for i in range(len(envs))}
for rnd in range(4):                                    # interleaved rounds: box drift hits every form alike
    for i, env in enumerate(envs):
        acc[i][0].append(run(16, env)[0])
        acc[i][1].append(run(256, env)[0])
for i, env in enumerate(envs):
    m16, m256 = statistics.mean(acc[i][0]), statistics.mean(acc[i][1])
    f = lambda ms, lb: lb * M * H * W * 88 / (ms / 1e3) / 1e9 / PEAK
    print(env, f"batch16 {m16 * 1e3:.1f} us {f(m16, 16):.3f} (rounds {[round(v * 1e3, 1) for v in acc[i][0]]}) | "
          f"batch256 {m256 * 1e3:.1f} us {f(m256, 256):.3f} (rounds {[round(v * 1e3) for v in acc[i][1]]})", flush=True)
