"""One training step inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`), after 3 warm-up steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from nextou_b200.factory import build_nextou
from nextou_b200.optim import FusedSGD
from nextou_b200.losses import DC_and_CE_and_BTI_Loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss

dev = torch.device("cuda", 0)
model = build_nextou(bench.CFG, seed=0).to(dev).train()
exclusion = bench.make_tensors(bench.SYNAPSE_EXCLUSION)
inner = DC_and_CE_and_BTI_Loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {},
                               {"dim": 3, "connectivity": 26, "inclusion": [], "exclusion": exclusion, "min_thick": 1},
                               weight_ce=1, weight_dice=1, weight_ti=1e-6, ignore_label=None, dice_class=MemoryEfficientSoftDiceLoss)
w = np.array([1 / (2 ** i) for i in range(5)]); w[-1] = 0
loss_fn = DeepSupervisionWrapper(inner, (w / w.sum()).tolist())
params = [p for p in model.parameters() if p.requires_grad]
TORCH_SGD = os.environ.get("NEXTOU_TORCH_SGD") == "1"
opt = torch.optim.SGD(params, lr=1e-2, momentum=0.99, nesterov=True, weight_decay=3e-5, fused=True) if TORCH_SGD else \
    FusedSGD(params, lr=1e-2, momentum=0.99, nesterov=True, weight_decay=3e-5, max_grad_norm=12)
x, t = bench.synthetic_batch(0)
x = x.to(dev); t = [a.to(dev) for a in t]


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = loss_fn(model(x), t)
    loss.backward()
    if TORCH_SGD:
        torch.nn.utils.clip_grad_norm_(params, 12)
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
