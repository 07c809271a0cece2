"""CPU, world_size 2, gloo: the data-parallel gradient exchange (nextou_b200.parallel) used by bench.py at N > 1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, overlap, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nextou_b200.parallel import GradientAllReducer
        torch.manual_seed(0)                                   # identical weights on every rank
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LeakyReLU(), torch.nn.Linear(16, 4),
                                    torch.nn.Linear(4, 4))
        for p in model[3].parameters():                        # a parameter that never receives a gradient
            p.requires_grad_(True)
        params = [p for p in model.parameters()]
        red = GradientAllReducer(params, world, bucket_mb=0.0002, overlap=overlap)   # several tiny buckets
        assert len(red.buckets) > 2
        g = torch.Generator().manual_seed(100 + rank)          # different patch per rank
        x = torch.randn(5, 8, generator=g)
        for it in range(2):
            red.zero_grad()
            model[2](model[1](model[0](x))).pow(2).sum().backward()   # model[3] unused -> no grad produced
            red.all_reduce()
        grads = [p.grad.clone() for p in params]
        # reference: mean over ranks of the single-process gradients
        ref_model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LeakyReLU(), torch.nn.Linear(16, 4))
        ref_model.load_state_dict({k: v for k, v in model.state_dict().items() if not k.startswith("3.")})
        acc = [torch.zeros_like(p) for p in ref_model.parameters()]
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            xr = torch.randn(5, 8, generator=gr)
            ref_model.zero_grad()
            ref_model(xr).pow(2).sum().backward()
            for a, p in zip(acc, ref_model.parameters()):
                a += p.grad / world
        ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(grads[:4], acc))
        ok = ok and all(float(gz.abs().max()) == 0.0 for gz in grads[4:])
        ok = ok and all(p.grad.data_ptr() >= red.buckets[red._bucket_of[p]].data_ptr() for p in params)
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False], ids=["overlap", "blocking"])
def test_gradient_all_reduce_world2_gloo(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, True), (1, True)]
