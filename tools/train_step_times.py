"""Time the training step (BASELINE configs[3]: 256x256 crops, 8 images per GPU, K=3 M=128) with CUDA events (GPU box).
    python tools/train_step_times.py [batch] [steps]          (torchrun for N > 1: adds the gradient all-reduce)
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from neural_image_compression_b200 import _lib, parallel  # noqa: E402
from neural_image_compression_b200.RateDistortionLoss import rd_loss  # noqa: E402
from tests import helpers as H  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
model = H.seeded_model(128, 3, "calib", precision="fp32").to(dev)
trainer = parallel.ShardedTrainer(model, 0.005, lr=1e-4)
torch.manual_seed(100 + rank)
xs = [torch.rand(B, 3, 256, 256, device=dev) for _ in range(4)]
lib = _lib.load()


def ev():
    return torch.cuda.Event(enable_timing=True)


for i in range(3):
    rd = trainer.step(xs[i % 4])
torch.cuda.synchronize()
l0 = lib.nic_launch_count()
t = [ev() for _ in range(5)]
acc = [0.0] * 4
e0, e1 = ev(), ev()
e0.record()
for i in range(steps):
    x = xs[i % 4]
    trainer.optimizer.zero_grad()
    t[0].record()
    out = model(x, training=True, lean=True)
    rd = rd_loss(out, x, 0.005)
    t[1].record()
    rd["loss"].backward()
    t[2].record()
    parallel.allreduce_gradients(trainer.buckets, None, trainer._flats)
    t[3].record()
    trainer.optimizer.step()
    t[4].record()
    torch.cuda.synchronize()
    for j in range(4):
        acc[j] += t[j].elapsed_time(t[j + 1])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
launches = (lib.nic_launch_count() - l0) / steps
if rank == 0:
    print(f"batch {B}/GPU x {world} GPU: step {ms:.2f} ms = {B * world / ms * 1e3:.0f} images/s; forward+loss {acc[0] / steps:.2f}  backward {acc[1] / steps:.2f}  "
          f"all-reduce {acc[2] / steps:.2f}  adam {acc[3] / steps:.2f} ms; {launches:.0f} kernels of this library per step; loss {float(rd['loss']):.4f}")
# back-to-back steps without the per-phase synchronisation
torch.cuda.synchronize()
e0.record()
for i in range(steps):
    rd = trainer.step(xs[i % 4])
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"ShardedTrainer.step back to back: {e0.elapsed_time(e1) / steps:.2f} ms / step = {B * world * steps / e0.elapsed_time(e1) * 1e3:.0f} images/s")
# the same step replayed from a CUDA graph (forward + loss + backward [+ Adam on one rank] captured once)
gtrainer = parallel.ShardedTrainer(model, 0.005, lr=1e-4, graph=True)
for i in range(3):
    rd = gtrainer.step(xs[i % 4])
torch.cuda.synchronize()
e0.record()
for i in range(steps):
    rd = gtrainer.step(xs[i % 4])
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"ShardedTrainer(graph=True).step: {e0.elapsed_time(e1) / steps:.2f} ms / step = {B * world * steps / e0.elapsed_time(e1) * 1e3:.0f} images/s; "
          f"loss {float(rd['loss']):.4f}, device step counter {int(gtrainer.optimizer._t_dev) if gtrainer.optimizer._t_dev is not None else -1}")
assert lib.nic_pipeline_status() == 0
if world > 1:
    dist.destroy_process_group()
