"""torchrun, N GPUs: the data-parallel training step (shards + bucketed NCCL gradient all-reduce) gives the gradient of the
full-batch step computed on one GPU.   torchrun --nproc-per-node N tools/train_dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from neural_image_compression_b200 import parallel  # noqa: E402
from neural_image_compression_b200.RateDistortionLoss import rd_loss  # noqa: E402
from tests import helpers as H  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
per = 2
torch.manual_seed(3)
x = torch.rand(per * world, 3, 128, 128)
nz, ny = torch.rand(per * world, 128, 2, 2) - 0.5, torch.rand(per * world, 128, 8, 8) - 0.5
model = H.seeded_model(128, 3, "calib", precision="fp32").to(dev)
sl = slice(rank * per, (rank + 1) * per)
out = model(x[sl].to(dev), training=True, noise=(nz[sl].to(dev), ny[sl].to(dev)))
rd = rd_loss(out, x[sl].to(dev), 0.005)
rd["loss"].backward()
buckets = parallel.grad_buckets(model.parameters(), 8 << 20)
n = parallel.allreduce_gradients(buckets)
torch.cuda.synchronize()
if rank == 0:
    ref = H.seeded_model(128, 3, "calib", precision="fp32").to(dev)
    o2 = ref(x.to(dev), training=True, noise=(nz.to(dev), ny.to(dev)))
    rd_loss(o2, x.to(dev), 0.005)["loss"].backward()
    worst = max(float((p.grad - q.grad).norm() / q.grad.norm()) for p, q in zip(model.parameters(), ref.parameters()))
    print(f"{world} ranks, {n} buckets: worst relative difference between the all-reduced shard gradients and the full-batch gradient: {worst:.2e}")
    assert worst < 1e-3, worst
dist.destroy_process_group()
