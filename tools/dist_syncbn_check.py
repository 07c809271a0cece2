"""2-GPU check of the SyncBatchNorm path (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_syncbn_check.py
Each rank normalises its own rows with nn.SyncBatchNorm through nextou_b200.dense; outputs, input gradients, parameter
gradients and running statistics are compared with an fp64 BatchNorm over the concatenated rows of all ranks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from nextou_b200 import dense


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rows, C = 4099, 33
    xs = [torch.randn(rows, C, generator=torch.Generator().manual_seed(7 + r)) * (1 + 0.5 * r) + 0.3 * r for r in range(world)]
    gs = [torch.randn(rows, C, generator=torch.Generator().manual_seed(70 + r)) for r in range(world)]
    bn = torch.nn.SyncBatchNorm(C).to(dev).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=torch.Generator(device=dev).manual_seed(1))
        bn.bias.uniform_(-0.5, 0.5, generator=torch.Generator(device=dev).manual_seed(2))
    ok = True
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        x = xs[rank].to(dev, dtype).requires_grad_(True)
        bn.zero_grad()
        y = dense.batch_norm_tokens(x, bn, 0.01)
        y.backward(gs[rank].to(dev, dtype))
        # fp64 reference over all rows
        ref = torch.nn.BatchNorm1d(C).double()
        ref.load_state_dict({k: v.detach().cpu().double() if v.dtype.is_floating_point else v.cpu() for k, v in bn.state_dict().items()})
        ref.running_mean.zero_(); ref.running_var.fill_(1.0)
        xa = torch.cat([t.to(dtype).double() for t in xs]).requires_grad_(True)
        ya = torch.nn.functional.leaky_relu(ref(xa), 0.01)
        ya.backward(torch.cat([t.to(dtype).double() for t in gs]))
        sl = slice(rank * rows, (rank + 1) * rows)
        def rel(a, b):
            return float((a.double().cpu() - b).norm() / b.norm().clamp_min(1e-30))
        errs = dict(y=rel(y.detach(), ya.detach()[sl]), dx=rel(x.grad, xa.grad[sl]))
        # parameter gradients: local sums; their sum over ranks equals the reference
        gw, gb = bn.weight.grad.clone(), bn.bias.grad.clone()
        dist.all_reduce(gw); dist.all_reduce(gb)
        errs["dgamma"], errs["dbeta"] = rel(gw, ref.weight.grad), rel(gb, ref.bias.grad)
        good = all(v <= tol for v in errs.values())
        ok = ok and good
        if rank == 0:
            print(f"syncbn {dtype}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()) + (" OK" if good else " FAIL"), flush=True)
        bn.running_mean.zero_(); bn.running_var.fill_(1.0)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if float(t) == 1.0 else 1)


if __name__ == "__main__":
    main()
