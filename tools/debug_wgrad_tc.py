"""Debug aid (GPU box): nic_conv_wgrad_tc against the fp32 weight-gradient kernel on a few layer shapes, with per-tap and
per-64-channel-slab error breakdowns when they disagree.   python tools/debug_wgrad_tc.py"""
import sys, torch, torch.nn as nn
sys.path.insert(0, ".")
from neural_image_compression_b200 import _lib
from neural_image_compression_b200.training import conv_wgrad
torch.manual_seed(4)
def run(conv, n, h, w, mode="rand"):
    conv = conv.cuda()
    ho, wo = (h, w) if conv.stride[0] == 1 else ((h // 2, w // 2) if not isinstance(conv, nn.ConvTranspose2d) else (h * 2, w * 2))
    x = torch.randn(n, h, w, conv.in_channels, device="cuda")
    g = torch.randn(n, ho, wo, conv.out_channels, device="cuda")
    a, _ = conv_wgrad(conv, x, g, n, h, w, _lib.LAYOUT_NHWC, _lib.LAYOUT_NHWC, arm="fp32")
    b, _ = conv_wgrad(conv, x, g, n, h, w, _lib.LAYOUT_NHWC, _lib.LAYOUT_NHWC, arm="bf16x3")
    torch.cuda.synchronize()
    err = float((a - b).norm() / a.norm())
    print(f"{str(conv)[:70]:70s} n={n} {h}x{w}: rel err {err:.2e}  status {_lib.load().nic_pipeline_status()}")
    if err > 1e-3:
        k = conv.kernel_size[0]
        for kh in range(k):
            print("   tap row", kh, ["%.1e" % float((a[:, :, kh, kw] - b[:, :, kh, kw]).norm() / a[:, :, kh, kw].norm()) for kw in range(k)])
        d = (a - b)[:, :, 0, 0]
        print("   quadrants (i-slab, j-slab) of tap 0:", [["%.1e" % float(d[i*64:(i+1)*64, j*64:(j+1)*64].norm() / a[:, :, 0, 0][i*64:(i+1)*64, j*64:(j+1)*64].norm()) for j in range(2)] for i in range(2)])
        print("   ratio b/a sample:", (b[:2, :4, 0, 0] / a[:2, :4, 0, 0]).tolist())
run(nn.Conv2d(128, 128, 1), 1, 8, 8)
run(nn.Conv2d(128, 128, 1), 1, 8, 16)
run(nn.Conv2d(128, 128, 1), 1, 16, 8)
run(nn.Conv2d(128, 128, 1), 2, 8, 8)
run(nn.Conv2d(128, 128, 1), 2, 16, 16)
run(nn.Conv2d(128, 128, 3, 1, 1), 1, 8, 8)
run(nn.Conv2d(128, 128, 5, 2, 2), 1, 16, 16)
run(nn.ConvTranspose2d(128, 128, 5, 2, 2, 1), 1, 8, 8)
run(nn.Conv2d(640, 1152, 1), 2, 8, 16)
run(nn.Conv2d(128, 128, 1), 2, 8, 16)
