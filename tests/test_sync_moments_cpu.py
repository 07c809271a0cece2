"""CPU, world_size 2, gloo: the cross-rank statistics of the SyncBatchNorm path (nextou_b200.ops.sync_moments) against the
moments of the concatenated batch — what torch.nn.SyncBatchNorm / upstream nnU-Net's DDP conversion computes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nextou_b200.ops import sync_moments
        rows, C = 37, 12
        xs = [torch.randn(rows, C, generator=torch.Generator().manual_seed(10 + r)) * (1 + r) + r for r in range(world)]
        x = xs[rank]
        sums = torch.stack([x.sum(0), (x * x).sum(0)])
        mean, invstd, unbiased, n = sync_moments(sums, rows, 1e-5)
        full = torch.cat(xs)
        ok = int(n) == rows * world
        ok = ok and torch.allclose(mean, full.mean(0), atol=1e-5)
        ok = ok and torch.allclose(invstd, torch.rsqrt(full.var(0, unbiased=False) + 1e-5), rtol=1e-4)
        ok = ok and torch.allclose(unbiased, full.var(0, unbiased=True), rtol=1e-4)
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sync_moments_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]
