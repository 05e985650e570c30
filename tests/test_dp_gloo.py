"""CPU, world_size 2 over gloo: the sharded rate-distortion reduction equals the single-process terms
(RateDistortionLoss.py:19-34 on the whole batch) - PSNR is taken after the SSE reduction, never averaged per rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_image_compression_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, per_image_all, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = parallel.shard_batch(per_image_all.t().contiguous(), rank, world).t().contiguous()   # [3, B/world]
        full = parallel.gather_per_image(local)
        terms = parallel.rd_terms_from_per_image(full, 512 * 768, 0.005)
        q.put((rank, full.numpy(), {k: v.numpy() for k, v in terms.items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_reduction_matches_single_process():
    torch.manual_seed(0)
    B = 8
    per_image = torch.stack([4e5 + 1e4 * torch.rand(B), 6e3 + 1e2 * torch.rand(B), 0.01 + 0.2 * torch.rand(B)]).float()
    ref = parallel.rd_terms_from_per_image(per_image, 512 * 768, 0.005)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, per_image, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, full, terms in got:
        assert np.array_equal(full, per_image.numpy()), "gathered per-image vectors differ from the unsharded batch"
        for k, v in ref.items():
            assert np.array_equal(terms[k], v.numpy()), (rank, k)
    # averaging per-rank PSNRs would be wrong: check the formula really is log-of-mean
    naive = np.mean([-10 * np.log10(per_image[2][i * 4:(i + 1) * 4].mean().item() + 1e-8) for i in range(2)])
    assert abs(naive - float(ref["psnr"])) > 1e-6


def test_shard_batch_rejects_ragged():
    import pytest
    with pytest.raises(ValueError):
        parallel.shard_batch(torch.zeros(5, 3, 64, 64), 0, 2)
    assert parallel.shard_batch(torch.zeros(4, 3, 64, 64), 1, 2).shape[0] == 2


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                   # identical weights on every rank
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
        net[3].weight.requires_grad_(False)                    # a frozen tensor is skipped by the buckets
        torch.manual_seed(1)
        x = torch.randn(8, 6)
        xl = parallel.shard_batch(x, rank, world)
        net(xl).pow(2).mean().backward()                       # mean over the local shard, like rd_loss
        net[2].bias.grad = None                                # a parameter without a gradient on this rank counts as zero
        buckets = parallel.grad_buckets(net.parameters(), bucket_bytes=64)     # tiny buckets: several all-reduces
        flats = [None] * len(buckets)
        n = parallel.allreduce_gradients(buckets, flats=flats)
        q.put((rank, n, len(buckets), {k: (p.grad.clone().numpy() if p.grad is not None else None) for k, p in net.named_parameters()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_gives_the_full_batch_gradient():
    """Config 4's exchange: bucketed all-reduce (average) of the gradients = gradient of the batch-mean loss on the whole batch."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    torch.manual_seed(1)
    x = torch.randn(8, 6)
    net(x).pow(2).mean().backward()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, n, nb, grads in got:
        assert n == nb and nb >= 2, (n, nb)
        assert grads["3.weight"] is None                       # frozen: untouched
        for k, p in net.named_parameters():
            if k in ("3.weight", "2.bias"):
                continue
            np.testing.assert_allclose(grads[k], p.grad.numpy(), rtol=1e-5, atol=1e-7, err_msg=f"rank {rank} {k}")
        assert np.array_equal(grads["2.bias"], np.zeros(3, np.float32))
    assert all(np.array_equal(got[0][3][k], got[1][3][k]) for k in got[0][3] if got[0][3][k] is not None)


def test_grad_buckets_cover_every_trainable_parameter_once_in_reverse_order():
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 4))
    b = parallel.grad_buckets(net.parameters(), bucket_bytes=48)
    flat = [p for bucket in b for p in bucket]
    assert [id(p) for p in flat] == [id(p) for p in reversed(list(net.parameters()))] and len(b) >= 2
