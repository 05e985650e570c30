"""The "library kernels to beat" (SURVEY.md section 8d): the reference's own PyTorch ops run EAGERLY ON THE GPU (cuDNN / cuBLAS / ATen),
i.e. what a user of the reference gets from `model.cuda()` - fp32 with TF32 off, and with torch's default (TF32 convs allowed).
The ops are those of oracle/forward.py (the restated reference forward, tensors moved to the device); 16 x 3 x 512 x 768, CUDA events.
    python tools/library_baseline.py        (GPU box; not part of bench.py - the oracle is test infrastructure)
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import forward as O  # noqa: E402
from tests import helpers as H  # noqa: E402

O.DEVICE = "cuda"
model = H.seeded_model(128, 3, "calib")
sd = {k: v.cuda() for k, v in model.state_dict().items()}
x = H.seeded_input((16, 3, 512, 768)).cuda()
for name, tf32 in (("fp32 (TF32 off)", False), ("torch default (TF32 convs)", True)):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        for _ in range(3):
            rd = O.rd_loss(O.forward(sd, x, 128, 3), x, 0.005)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            rd = O.rd_loss(O.forward(sd, x, 128, 3), x, 0.005)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"reference ops, torch eager on the GPU, {name}: {ms:.1f} ms per 16 images = {16 / ms * 1e3:.0f} images/s; bpp {rd['bpp_total']:.6f} psnr {rd['psnr']:.6f}")
