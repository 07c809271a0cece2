"""Kernel-time breakdown of one training step (torch.profiler, CUDA activities). Diagnostic only.
    python tools/profile_step.py [steps]
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import profile, ProfilerActivity
from tests import helpers as H
from nextou_b200.losses import DC_and_CE_and_BTI_Loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss
import numpy as np

def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda", 0)
    model = H.build_product(bench.CFG, seed=0).to(dev).train()
    exclusion = bench.make_tensors(bench.SYNAPSE_EXCLUSION)
    inner = DC_and_CE_and_BTI_Loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {},
                                   {"dim": 3, "connectivity": 26, "inclusion": [], "exclusion": exclusion, "min_thick": 1},
                                   weight_ce=1, weight_dice=1, weight_ti=1e-6, ignore_label=None, dice_class=MemoryEfficientSoftDiceLoss)
    w = np.array([1 / (2 ** i) for i in range(5)]); w[-1] = 0
    loss_fn = DeepSupervisionWrapper(inner, (w / w.sum()).tolist())
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-2, momentum=0.99, nesterov=True, weight_decay=3e-5, fused=True)
    x, t = bench.synthetic_batch(0)
    x = x.to(dev); t = [a.to(dev) for a in t]
    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(x); loss = loss_fn(outs, t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 12); opt.step()
    global _step_fn
    _step_fn = step
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps): step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    rows = [(e.key, e.device_time_total / 1e3 / steps, e.count / steps) for e in ka if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"total CUDA kernel time per step: {tot:.1f} ms")
    for k, ms, n in rows[:45]:
        print(f"{ms:9.2f} ms {100*ms/tot:5.1f}% n={n:6.1f}  {k[:130]}")


def copy_breakdown(steps=1):
    """Which aten::copy_ / contiguous / fill calls cost GPU time, by shape and Python call site."""
    import torch
    from torch.profiler import profile, ProfilerActivity
    global _step_fn
    dev = torch.device("cuda", 0)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
        for _ in range(steps):
            _step_fn()
        torch.cuda.synchronize()
    ka = prof.key_averages(group_by_input_shape=True, group_by_stack_n=6)
    rows = [e for e in ka if e.key in ("aten::copy_", "aten::fill_", "aten::add", "aten::sum", "aten::add_", "aten::zero_", "aten::mul")]
    rows.sort(key=lambda e: -e.device_time_total)
    for e in rows[:40]:
        stack = [s for s in e.stack if "nextou_b200" in s or "bench" in s or "tools" in s][:3]
        print(f"{e.device_time_total/1e3/steps:8.2f} ms n={e.count/steps:5.0f} {e.key:12s} {str(e.input_shapes)[:70]:70s} {' <- '.join(x.split('/')[-1][:60] for x in stack)}")


main()
if len(sys.argv) > 2 and sys.argv[2] == 'copies':
    copy_breakdown()
