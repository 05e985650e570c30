"""Print the relative error of every gradient tensor of one training step against the oracle's autograd (GPU box)."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import backward as OB
from tests import helpers as H
from neural_image_compression_b200.RateDistortionLoss import rd_loss

K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
import os
M = int(os.environ.get("M", "128")); ARM = os.environ.get("ARM", "fp32")
shape = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (2, 3, 128, 128)
model = H.seeded_model(M, K, "calib", precision="fp32")
model.train_precision = ARM
sd = {k: v.clone() for k, v in model.state_dict().items()}
x = H.seeded_input(shape)
B, _, Hh, W = shape
torch.manual_seed(int(sys.argv[3]) if len(sys.argv) > 3 else 11)
_, nz, ny = OB.noise_with_margin(sd, x, M, K, int(sys.argv[3]) if len(sys.argv) > 3 else 11)
ref_rd, ref_g, _ = OB.loss_and_grads(sd, x, M, K, nz, ny, 0.005)
_, g64, _ = OB.loss_and_grads(sd, x, M, K, nz, ny, 0.005, dtype=torch.float64)
model = model.cuda()
out = model(x.cuda(), training=True, noise=(nz.cuda(), ny.cuda()))
rd = rd_loss(out, x.cuda(), 0.005)
rd["loss"].backward()
torch.cuda.synchronize()
print("loss", float(rd["loss"]), ref_rd["loss"])
for k, p in model.named_parameters():
    r64 = g64[k]
    e = float((p.grad.double().cpu() - r64).norm() / r64.norm())
    eref = float((ref_g[k].double() - r64).norm() / r64.norm())
    if e > 2e-5 or "-v" in sys.argv: print(f"{k:45s} ours-vs-fp64 {e:.2e}   oracle-fp32-vs-fp64 {eref:.2e}   norm {float(r64.norm()):.3e}")
