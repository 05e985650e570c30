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
