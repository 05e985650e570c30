"""One evaluation forward (+ rd_loss) of HierarchicalMixtureResidual(128, K=3) at 4 x 3 x 512 x 768, eager, for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_residual.py [train]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_image_compression_b200 import parallel  # noqa: E402
from neural_image_compression_b200.Models import HierarchicalMixtureResidual  # noqa: E402
from neural_image_compression_b200.RateDistortionLoss import rd_loss  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda:0")
model = HierarchicalMixtureResidual(128, K=3, precision="bf16x3").to(dev)
if len(sys.argv) > 1 and sys.argv[1] == "train":
    x = torch.rand(8, 3, 256, 256, device=dev)
    tr = parallel.ShardedTrainer(model, 0.005, lr=1e-4, graph=False)
    for _ in range(2):
        tr.step(x)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("timed")
    tr.step(x)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
else:
    x = torch.rand(4, 3, 512, 768, device=dev)
    with torch.no_grad():
        for _ in range(2):
            rd_loss(model(x, training=False, lean=True), x, 0.005)
        torch.cuda.synchronize()
        marker = torch.zeros(7, device=dev)          # a recognisable launch (FillFunctor) in front of the measured pass
        rd_loss(model(x, training=False, lean=True), x, 0.005)
        torch.cuda.synchronize()
