"""Per-call CUDA-event timing of one eval forward + rd terms (run on the GPU box):  python tools/layer_times.py [precision] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neural_image_compression_b200 import engine, _lib
from neural_image_compression_b200.RateDistortionLoss import rd_terms
from tests import helpers as H

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
model = H.seeded_model(128, 3, "calib", precision=prec).cuda()
x = H.seeded_input((B, 3, 512, 768)).cuda()
records = []
orig_run = engine.ConvOp.run
def timed_run(self, xin, n, h, w, precision, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig_run(self, xin, n, h, w, precision, *a, **k); e1.record()
    cv = self.conv
    macs = n * out.numel() // out.shape[0] // 1  # placeholder
    ho, wo = engine.conv_out_hw(cv, h, w)
    taps = 12 if self.mask_a else cv.kernel_size[0] * cv.kernel_size[1]
    if isinstance(cv, torch.nn.ConvTranspose2d):
        gmac = n * h * w * cv.in_channels * cv.out_channels * taps / 1e9
    else:
        gmac = n * ho * wo * cv.in_channels * cv.out_channels * taps / 1e9
    if self.gdn is not None:
        gmac += n * ho * wo * cv.out_channels * cv.out_channels / 1e9
    records.append((f"{type(cv).__name__}({cv.in_channels}->{cv.out_channels},k{cv.kernel_size[0]},s{cv.stride[0]}) {h}x{w} epi{self.epilogue}", e0, e1, gmac))
    return out
engine.ConvOp.run = timed_run
for it in range(3):
    records.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out = model(x, training=False)
    per, sc = rd_terms(out, x, 0.005)
    t1.record()
    torch.cuda.synchronize()
tot = 0.0
for name, e0, e1, gmac in records:
    ms = e0.elapsed_time(e1); tot += ms
    print(f"{ms*1000:9.1f} us  {2*gmac/ms:8.1f} TFLOP/s  {name}")
print(f"conv calls total {tot:.3f} ms; whole step {t0.elapsed_time(t1):.3f} ms (B={B}, {prec})")
