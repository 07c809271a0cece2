"""Race check of the deferred weight-gradient join (native._complete_wgrad): the same backward pass with the join deferred to
the end of the pass and with a join per layer must give the same gradients (up to the rounding of the split-K atomics).
    python tools/defer_check.py [repetitions]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from nextou_b200 import ops
from nextou_b200.factory import build_nextou

dev = torch.device("cuda", 0)
model = build_nextou(bench.CFG, seed=0).to(dev).train()
x, t = bench.synthetic_batch(0)
x = x.to(dev)
params = [p for p in model.parameters() if p.requires_grad]
names = [n for n, p in model.named_parameters() if p.requires_grad]


from nextou_b200 import native

_returned = []
_orig = native._complete_wgrad


def _recording(side, param, finish, keep):
    dw = _orig(side, param, finish, keep)
    _returned.append((param, dw.untyped_storage().data_ptr()))
    return dw


native._complete_wgrad = _recording


def grads(defer):
    ops.DEFER_WGRAD_JOIN = defer
    _returned.clear()
    for p in params:
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    sum(o.float().square().mean() for o in outs).backward()
    torch.cuda.synchronize()
    return [None if p.grad is None else p.grad.detach().clone() for p in params]


reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
base = grads(False)
again = grads(False)
noise = max(float((a - b).norm() / b.norm().clamp_min(1e-20)) for a, b in zip(again, base) if a is not None)
print(f"joined vs joined (atomics noise): worst relative L2 {noise:.2e}")
bad = 0
for r in range(reps):
    g = grads(True)
    worst = ("", 0.0)
    for n, a, b in zip(names, g, base):
        if a is None:
            continue
        e = float((a - b).norm() / b.norm().clamp_min(1e-20))
        if e > worst[1]:
            worst = (n, e)
    errs = sorted(((float((a - b).norm() / b.norm().clamp_min(1e-20)), n) for n, a, b in zip(names, g, base) if a is not None), reverse=True)
    if r == 0:
        for e, n in errs[:12]:
            print(f"      {n} {e:.2e}")
    # every deferred gradient must have become param.grad WITHOUT a copy (a copy would be a main-stream kernel that reads the
    # tensor before the side stream has written it)
    copied = [1 for p, sp in _returned if p.grad is None or p.grad.untyped_storage().data_ptr() != sp]
    ok = worst[1] <= max(20 * noise, 1e-4) and not copied
    bad += not ok
    print(f"deferred run {r}: worst {worst[0]} {worst[1]:.2e}, {len(_returned) - len(copied)}/{len(_returned)} gradients taken without a copy "
          + ("OK" if ok else "MISMATCH"))
print("DEFER CHECK " + ("PASSED" if bad == 0 else "FAILED"))
