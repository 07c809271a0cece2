"""Whole-step CUDA graph (nextou_b200/graphed.py): replay must reproduce the eager step."""
import copy

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(seed=0):
    from nextou_b200.losses import DC_and_CE_and_BTI_Loss, DeepSupervisionWrapper
    from oracle.ref_shims import SYNAPSE_EXCLUSION, make_tensors
    cfg = dict(H.MINI3D)
    model = H.build_product(cfg, seed=seed).to(DEV).train()
    exc = make_tensors(SYNAPSE_EXCLUSION)
    exc = [[e.to(DEV) for e in p] if isinstance(p, list) else p.to(DEV) for p in exc]
    inner = DC_and_CE_and_BTI_Loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {},
                                   {"dim": 3, "connectivity": 26, "inclusion": [], "exclusion": exc, "min_thick": 1},
                                   weight_ce=1, weight_dice=1, weight_ti=1e-6)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x.to(DEV))
    n_cls = outs[0].shape[1]
    targets = [torch.randint(0, n_cls, (1, 1, *o.shape[2:]), generator=g).float() for o in outs]
    w = np.array([1 / 2 ** i for i in range(len(outs))])
    w[-1] = 0
    loss_fn = DeepSupervisionWrapper(inner, (w / w.sum()).tolist())
    return model, loss_fn, x, targets


def test_graph_replay_matches_eager_steps():
    from nextou_b200.graphed import GraphedTrainStep, graph_safe
    model, loss_fn, x, targets = _setup()
    assert graph_safe(model) is None
    state0 = copy.deepcopy(model.state_dict())
    xd, td = x.to(DEV), [t.to(DEV) for t in targets]

    def make_opt(m):
        return torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-2, momentum=0.99, nesterov=True,
                               weight_decay=3e-5)

    # eager: 2 steps from state0 with a fresh optimizer
    opt = make_opt(model)
    eager = []
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(model(xd), td)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(opt.param_groups[0]["params"], 12)
        opt.step()
        eager.append(loss.item())
    w_eager = {k: v.clone() for k, v in model.state_dict().items()}
    del loss    # a live autograd graph pins AccumulateGrad nodes to the eager stream, which would invalidate the capture

    # graphed: build (its warm-up steps move the weights), rewind to state0 and a clean momentum, replay 2 steps
    opt2 = make_opt(model)
    step = GraphedTrainStep(model, loss_fn, opt2, xd, td, clip_grad_norm=12, warmup=1)
    assert step.launches_per_step > 100
    with torch.no_grad():
        model.load_state_dict(state0)
        for st in opt2.state.values():
            st["momentum_buffer"].zero_()
    got = [step(x.pin_memory(), [t.pin_memory() for t in targets]).item() for _ in range(2)]
    # first step: identical kernels on identical data (atomics in gather / pooling backward reorder fp32 sums, hence
    # not bit-exact); momentum buffer zero == "first step" semantics of SGD since buf = grad either way up to dampening 0
    assert abs(got[0] - eager[0]) <= 1e-6 * abs(eager[0])
    assert abs(got[1] - eager[1]) <= 5e-2 * abs(eager[1])      # second step sees chaotic kNN flips of the updated net
    num = sum((model.state_dict()[k].float() - w_eager[k].float()).pow(2).sum().item() for k in w_eager
              if w_eager[k].dtype.is_floating_point)
    den = sum(w_eager[k].float().pow(2).sum().item() for k in w_eager if w_eager[k].dtype.is_floating_point)
    assert (num / den) ** 0.5 < 1e-2


def test_graph_refuses_host_random_dilation():
    from nextou_b200._lib import NextouError
    from nextou_b200.blocks import DyGraphConv
    from nextou_b200.graphed import GraphedTrainStep, graph_safe
    model, loss_fn, x, targets = _setup()
    first = next(m for m in model.modules() if isinstance(m, DyGraphConv))
    first.d = 2
    assert "dilation" in graph_safe(model)
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    with pytest.raises(NextouError):
        GraphedTrainStep(model, loss_fn, opt, x.to(DEV), [t.to(DEV) for t in targets])


def test_graph_replay_follows_learning_rate_schedule():
    """nnU-Net's PolyLRScheduler assigns param_groups[i]['lr'] every epoch.  A captured step must follow it: with a fused
    optimizer the learning rate lives in a device tensor that is re-filled when the group holds a new number (lr = 0 must
    freeze the weights); an optimizer that bakes a host-scalar lr into the graph must raise instead of training on."""
    from nextou_b200._lib import NextouError
    from nextou_b200.graphed import GraphedTrainStep
    model, loss_fn, x, targets = _setup()
    xd, td = x.to(DEV), [t.to(DEV) for t in targets]
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-2, momentum=0.0, weight_decay=3e-5, fused=True)
    step = GraphedTrainStep(model, loss_fn, opt, xd, td, clip_grad_norm=12, warmup=1)
    assert isinstance(opt.param_groups[0]["lr"], torch.Tensor) and opt.param_groups[0]["lr"].is_cuda
    step(xd, td)
    w0 = [p.detach().clone() for p in params]
    opt.param_groups[0]["lr"] = 0.0                      # what a scheduler does
    step(xd, td)
    assert all(torch.equal(a, b.detach()) for a, b in zip(w0, params)), "replay ignored the new learning rate"
    opt.param_groups[0]["lr"] = 1e-2
    step(xd, td)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(w0, params))
    opt.param_groups[0]["weight_decay"] = 0.0            # cannot be patched into a captured graph
    with pytest.raises(NextouError):
        step(xd, td)
    opt.param_groups[0]["weight_decay"] = 3e-5
    del step
    # host-scalar lr baked into the capture: changing it must be detected
    opt2 = torch.optim.SGD(params, lr=1e-2, momentum=0.0)
    step2 = GraphedTrainStep(model, loss_fn, opt2, xd, td, clip_grad_norm=12, warmup=1)
    step2(xd, td)
    opt2.param_groups[0]["lr"] = 5e-3
    with pytest.raises(NextouError):
        step2(xd, td)


def test_prefetched_steps_match_plain_replays():
    """GraphedTrainStep.prefetch / step_prefetched (next batch staged on a copy stream during the current replay) must train
    like step(x_host, t_host): the same losses for a sequence of DIFFERENT batches fed from pinned host memory."""
    from nextou_b200.graphed import GraphedTrainStep
    from nextou_b200._lib import NextouError
    model, loss_fn, x, targets = _setup()
    state0 = copy.deepcopy(model.state_dict())
    g = torch.Generator().manual_seed(9)
    batches = []
    for _ in range(4):
        xb = torch.randn(x.shape, generator=g).pin_memory()
        tb = [torch.randint(0, 3, t.shape, generator=g).float().pin_memory() for t in targets]
        batches.append((xb, tb))

    def run(prefetched):
        model.load_state_dict(state0)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.SGD(params, lr=1e-2, momentum=0.9, fused=True)
        step = GraphedTrainStep(model, loss_fn, opt, x.to(DEV), [t.to(DEV) for t in targets], clip_grad_norm=12, warmup=1)
        model.load_state_dict(state0)           # the warm-up steps moved the weights
        losses = []
        if prefetched:
            with pytest.raises(NextouError):
                step.step_prefetched()
            step.prefetch(*batches[0])
            for i in range(len(batches)):
                loss = step.step_prefetched()
                if i + 1 < len(batches):
                    step.prefetch(*batches[i + 1])
                losses.append(loss.item())
        else:
            for xb, tb in batches:
                losses.append(step(xb, tb).item())
        del step
        return losses

    plain, pre = run(False), run(True)
    assert max(plain) - min(plain) > 1e-2        # different batches, visibly different losses
    # split-K weight gradients are reduced with floating-point atomics: two runs agree to rounding, not bit for bit
    assert pre == pytest.approx(plain, rel=2e-3)
