"""Pure GPU time of one conv layer of the M = 128 model (GPU box): 10 launches captured in a CUDA graph, replayed; environment
overrides (NIC_TC_MT, NIC_TC_NSA, NIC_TC_NSB, NIC_TC_SWAP ...) are read per launch, so set them for the process.
    python tools/layer_bench.py [bf16x3|bf16] layer [layer ...]      (layers: see CFG; 'all' = every layer)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from neural_image_compression_b200 import engine
from neural_image_compression_b200._lib import EPI_BIAS, EPI_GDN, EPI_IGDN, EPI_LRELU, LAYOUT_NCHW
from neural_image_compression_b200.gdn import GDN

T = nn.ConvTranspose2d
CFG = {
    "k1": (nn.Conv2d(3, 128, 5, 2, 2), EPI_GDN, (512, 768), dict(in_layout=LAYOUT_NCHW, image=True)),
    "k2": (nn.Conv2d(128, 128, 5, 2, 2), EPI_GDN, (256, 384), {}),
    "k3": (nn.Conv2d(128, 128, 5, 2, 2), EPI_GDN, (128, 192), {}),
    "k4": (nn.Conv2d(128, 128, 5, 2, 2), EPI_BIAS, (64, 96), dict(out_dtype=torch.float32)),
    "ha1": (nn.Conv2d(128, 128, 3, 1, 1), EPI_LRELU, (32, 48), {}),
    "ha2": (nn.Conv2d(128, 128, 5, 2, 2), EPI_LRELU, (32, 48), {}),
    "ha3": (nn.Conv2d(128, 128, 5, 2, 2), EPI_BIAS, (16, 24), dict(out_dtype=torch.float32)),
    "hs1": (T(128, 128, 5, 2, 2, output_padding=1), EPI_LRELU, (8, 12), {}),
    "hs2": (T(128, 192, 5, 2, 2, output_padding=1), EPI_LRELU, (16, 24), {}),
    "hs3": (nn.Conv2d(192, 256, 3, 1, 1), EPI_BIAS, (32, 48), {}),
    "ctx": (nn.Conv2d(128, 256, 5, 1, 2), EPI_BIAS, (32, 48), dict(mask_a=1)),
    "ep1": (nn.Conv2d(512, 640, 1), EPI_LRELU, (32, 48), {}),
    "ep2": (nn.Conv2d(640, 640, 1), EPI_LRELU, (32, 48), {}),
    "ep3": (nn.Conv2d(640, 1152, 1), EPI_BIAS, (32, 48), dict(out_layout=LAYOUT_NCHW, out_dtype=torch.float32)),
    "d1": (T(128, 128, 5, 2, 2, output_padding=1), EPI_IGDN, (32, 48), {}),
    "d2": (T(128, 128, 5, 2, 2, output_padding=1), EPI_IGDN, (64, 96), {}),
    "d3": (T(128, 128, 5, 2, 2, output_padding=1), EPI_IGDN, (128, 192), {}),
    "d4": (T(128, 3, 5, 2, 2, output_padding=1), EPI_BIAS, (256, 384), dict(out_layout=LAYOUT_NCHW, out_dtype=torch.float32)),
}
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
layers = sys.argv[2:] or ["all"]
if layers == ["all"]:
    layers = list(CFG)
B = int(os.environ.get("LB_BATCH", "16"))
dev = torch.device("cuda:0")
tot = 0.0
for which in layers:
    conv, epi, (h, w), kw = CFG[which]
    kw = dict(kw)
    mask_a = kw.pop("mask_a", 0)
    image = kw.pop("image", False)
    conv = conv.to(dev)
    g = GDN(128, inverse=(epi == EPI_IGDN)).to(dev) if epi in (EPI_GDN, EPI_IGDN) else None
    op = engine.ConvOp(conv, epi, gdn=g, mask_a=mask_a)
    x = torch.round(3 * torch.randn(B, h, w, conv.in_channels, device=dev)) if os.environ.get("LB_INT") else torch.randn(B, h, w, conv.in_channels, device=dev)
    x = engine.to_pair(x) if prec == "bf16x3" else x.to(torch.bfloat16)
    if image:
        x = torch.rand(B, 3, h, w, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):
            op.run(x, B, h, w, prec, **kw)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(10):
                y = op.run(x, B, h, w, prec, **kw)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            gr.replay()
        e1.record(s)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 50
    tot += us
    taps = 12 if mask_a else conv.kernel_size[0] * conv.kernel_size[1]
    px = B * h * w if isinstance(conv, T) else B * (h // conv.stride[0]) * (w // conv.stride[0])
    gf = 2.0 * px * conv.in_channels * conv.out_channels * taps / 1e9
    print(f"{which:5s} {us:8.1f} us  {gf / us * 1e3:7.1f} TFLOP/s algorithmic", flush=True)
print(f"total {tot:.1f} us")
