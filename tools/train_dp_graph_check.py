"""torchrun, N GPUs: ShardedTrainer(graph=True) with the gradient all-reduce captured INSIDE the step's CUDA graph (early bucket
on a side stream beside g_a's backward, late bucket, Adam with 1 / world folded in).  Checks that after K steps on rank-specific
data every rank holds bit-identical parameters (they all applied the same summed gradient), that the loss goes down, and times
the step against the variant with the all-reduce outside the graph (NIC_GRAPH_ALLREDUCE=0).
    torchrun --nproc-per-node N tools/train_dp_graph_check.py"""
import hashlib
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from neural_image_compression_b200 import parallel  # noqa: E402
from tests import helpers as H  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)


def digest(model):
    h = hashlib.sha256()
    for p in model.parameters():
        h.update(p.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def run(in_graph: bool, steps: int = 24):
    os.environ["NIC_GRAPH_ALLREDUCE"] = "1" if in_graph else "0"
    model = H.seeded_model(128, 3, "calib", precision="bf16x3").to(dev)
    tr = parallel.ShardedTrainer(model, 0.005, lr=1e-4, graph=True)
    g = torch.Generator(device="cpu"); g.manual_seed(100 + rank)
    xs = [torch.rand((8, 3, 256, 256), generator=g).to(dev) for _ in range(4)]
    losses = []
    for i in range(4):
        losses.append(float(tr.step(xs[i % 4])["loss"]))
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        rd = tr.step(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    losses.append(float(rd["loss"]))
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    d = digest(model)
    all_d = [None] * world
    dist.all_gather_object(all_d, d)
    return float(ms), losses, all_d, tr.optimizer.t


for mode in (True, False):
    ms, losses, digests, t = run(mode)
    if rank == 0:
        same = len(set(digests)) == 1
        print(f"{world} ranks, all-reduce {'inside' if mode else 'outside'} the graph: {ms:.3f} ms / step (max over ranks), loss {losses[0]:.3f} -> {losses[-1]:.3f}, "
              f"optimizer steps {t}, parameters identical across ranks: {same}", flush=True)
        assert same and losses[-1] < losses[0], (digests, losses)
dist.destroy_process_group()
