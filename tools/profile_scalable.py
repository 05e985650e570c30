"""Kernel-time totals of one ScalableImageCoding(192, 128) eval forward at 2048 x 1536 (GPU box): python tools/profile_scalable.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from tests import helpers as H
from neural_image_compression_b200.RateDistortionLoss import vision_rd_loss

model = H.seeded_scalable_model(192, 128, 1, "calib", precision="bf16x3").cuda()
x = H.seeded_input((1, 3, 1536, 2048)).cuda()
with torch.no_grad():
    for _ in range(2):
        vision_rd_loss(model(x, training=False), x, 0.005)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        vision_rd_loss(model(x, training=False), x, 0.005)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
