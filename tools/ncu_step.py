"""One training step inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`), after 3 warm-up steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from tests import helpers as H
from nextou_b200.losses import DC_and_CE_and_BTI_Loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss

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
        loss = loss_fn(model(x), t)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 12)
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
