"""CTA-pair form of conv_tc_kernel against the one-CTA form, bit for bit, on the layer shapes of the M = 128 model (GPU box):
    python tools/pair_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from neural_image_compression_b200 import engine, _lib
from neural_image_compression_b200._lib import EPI_BIAS, EPI_GDN, EPI_IGDN, EPI_LRELU
from neural_image_compression_b200.gdn import GDN
dev = torch.device("cuda:0")
T = nn.ConvTranspose2d
torch.manual_seed(0)
def run(conv, epi, h, w, B, gdn=None):
    op = engine.ConvOp(conv.to(dev), epi, gdn=gdn.to(dev) if gdn is not None else None)
    x = engine.to_pair(torch.randn(B, h, w, conv.in_channels, device=dev))
    os.environ.pop("NIC_TC_PAIR", None)
    a = op.run(x, B, h, w, "bf16x3").clone()
    os.environ["NIC_TC_PAIR"] = "1"
    b = op.run(x, B, h, w, "bf16x3").clone()
    torch.cuda.synchronize()
    os.environ.pop("NIC_TC_PAIR", None)
    same = torch.equal(a.view(torch.int16), b.view(torch.int16)) if a.dtype == torch.bfloat16 else torch.equal(a, b)
    d = (a.float() - b.float()).abs().max().item()
    print(type(conv).__name__, conv.stride, (h, w), B, "identical" if same else f"DIFF max {d:.3e}", "status", _lib.load().nic_pipeline_status(), flush=True)
run(nn.Conv2d(128, 128, 5, 2, 2), EPI_GDN, 256, 384, 16, GDN(128))
run(nn.Conv2d(128, 128, 5, 2, 2), EPI_GDN, 128, 192, 16, GDN(128))
run(nn.Conv2d(128, 128, 5, 2, 2), EPI_BIAS, 250, 382, 5)
run(T(128, 128, 5, 2, 2, output_padding=1), EPI_IGDN, 64, 96, 16, GDN(128, inverse=True))
run(T(128, 128, 5, 2, 2, output_padding=1), EPI_IGDN, 128, 192, 16, GDN(128, inverse=True))
run(nn.Conv2d(128, 128, 3, 1, 1), EPI_LRELU, 128, 192, 8)
