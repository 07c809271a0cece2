"""Optimizer-side kernels (csrc/optim.cu, nextou_b200/optim.py): multi-tensor clip + SGD-Nesterov against torch.optim.SGD on
fp32 master weights, and the persistent operand packs the fused step keeps up to date."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tensors(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(33, 1, 1, 3, 3), (33,), (66, 33, 3, 3, 3), (132, 66, 3, 3, 3), (7,), (324, 324, 2, 2, 2), (1296, 324, 1, 1, 1), (1,),
              (14, 33, 1, 1, 1), (100003,)]
    return [torch.randn(s, generator=g) * 0.1 for s in shapes]


@pytest.mark.parametrize("momentum,nesterov,wd,max_norm", [(0.99, True, 3e-5, 12.0), (0.9, False, 0.0, 0.5), (0.0, False, 1e-2, None),
                                                           (0.99, True, 3e-5, 1e9)])
def test_fused_sgd_matches_torch_sgd(momentum, nesterov, wd, max_norm):
    """3 steps of clip_grad_norm_ + torch.optim.SGD vs FusedSGD on the same fp32 tensors and gradients: parameters and
    momentum buffers to 1e-6 (relative to the tensor's largest entry), the reported gradient norm to 1e-6."""
    from nextou_b200.optim import FusedSGD
    ref = [torch.nn.Parameter(t.clone().to(DEV)) for t in _tensors(1)]
    own = [torch.nn.Parameter(t.clone().to(DEV)) for t in _tensors(1)]
    o_ref = torch.optim.SGD(ref, lr=0.05, momentum=momentum, nesterov=nesterov, weight_decay=wd)
    o_own = FusedSGD(own, lr=0.05, momentum=momentum, nesterov=nesterov, weight_decay=wd, max_grad_norm=max_norm)
    for step in range(3):
        grads = [t.to(DEV) * (3.0 if step == 1 else 0.3) for t in _tensors(10 + step)]
        for p, q, gr in zip(ref, own, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        if step == 2:                                   # a scheduler assigns a new learning rate
            o_own.param_groups[0]["lr"] = 0.01
            o_ref.param_groups[0]["lr"] = 0.01
        norm_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm) if max_norm is not None else None
        o_ref.step()
        o_own.step()
        if norm_ref is not None:
            assert abs(o_own.grad_norm.item() - norm_ref.item()) <= 1e-6 * norm_ref.item()
        for p, q in zip(ref, own):
            scale = p.detach().abs().max().item() + 1e-12
            assert (p.detach() - q.detach()).abs().max().item() <= 1e-6 * scale
            if momentum:
                a, b = o_ref.state[p]["momentum_buffer"], o_own.state[q]["momentum_buffer"]
                assert (a - b).abs().max().item() <= 1e-6 * (a.abs().max().item() + 1e-12)


def test_fused_sgd_follows_lr_and_skips_params_without_grad():
    from nextou_b200.optim import FusedSGD
    own = [torch.nn.Parameter(t.clone().to(DEV)) for t in _tensors(2)]
    before = [p.detach().clone() for p in own]
    opt = FusedSGD(own, lr=0.0, momentum=0.9, nesterov=True, weight_decay=1e-3, max_grad_norm=12)
    for p in own[:-1]:
        p.grad = torch.randn_like(p)
    opt.step()                                           # lr = 0: nothing moves; the last tensor has no gradient at all
    assert all(torch.equal(a, p.detach()) for a, p in zip(before, own))
    opt.param_groups[0]["lr"] = 0.1
    opt.step()
    assert all(not torch.equal(a, p.detach()) for a, p in zip(before[:-1], own[:-1]))
    assert torch.equal(before[-1], own[-1].detach())


def test_pack_weight_gap_layout():
    """Packs with a zero input-channel gap == packs of the weight with explicit zero columns (decoder concat layout)."""
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(3)
    w = torch.randn(33, 66, 1, 3, 3, generator=g).to(DEV)
    ca, pa = 33, 40
    wz = torch.cat([w[:, :ca], w.new_zeros(33, pa - ca, 1, 3, 3), w[:, ca:]], 1)
    a0, b0 = ops.pack_weight_pair(wz, conv=True, flip_b=True)
    a1, b1 = ops.pack_weight_pair(w, conv=True, flip_b=True, gap=(ca, pa))
    assert a0.shape == a1.shape and b0.shape == b1.shape
    assert torch.equal(a0, a1) and torch.equal(b0, b1)


def test_fused_step_keeps_operand_packs_current():
    """A conv + 1x1 layer trained with FusedSGD: after the first step the forward launches NO pack kernel (the optimizer
    rewrote the persistent packs), outputs equal those of a twin trained with torch.optim.SGD (which re-packs every step),
    and an in-place edit of the weight by torch invalidates the packs."""
    from nextou_b200 import _lib, dense, ops
    from nextou_b200.optim import FusedSGD
    torch.manual_seed(0)

    def make():
        torch.manual_seed(5)
        conv = torch.nn.Conv3d(16, 24, 3, padding=1).to(DEV)
        lin = torch.nn.Conv3d(24, 40, 1).to(DEV)
        return conv, lin

    def fwd(conv, lin, x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            tok = ops.as_tokens(x)
            h, sp = dense.conv_tokens(tok, 1, (6, 8, 8), conv)
            return dense.linear_tokens(h, lin)

    x = torch.randn(1, 16, 6, 8, 8, device=DEV)
    (c1, l1), (c2, l2) = make(), make()
    p1, p2 = list(c1.parameters()) + list(l1.parameters()), list(c2.parameters()) + list(l2.parameters())
    o1 = torch.optim.SGD(p1, lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-4)
    o2 = FusedSGD(p2, lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-4, max_grad_norm=None)
    names = []
    for step in range(3):
        o1.zero_grad(set_to_none=True)
        o2.zero_grad(set_to_none=True)
        y1 = fwd(c1, l1, x)
        n0 = _lib.launch_count()
        y2 = fwd(c2, l2, x)
        launched = _lib.launch_count() - n0
        names.append(launched)
        # (the twins' weights agree to 1e-6, so a bf16 rounding of a weight may differ here and there)
        assert ((y1.float() - y2.float()).norm() / y1.float().norm()).item() <= 2e-3, step
        y1.float().square().mean().backward()
        y2.float().square().mean().backward()
        o1.step()
        o2.step()
    assert names[0] == names[1] + 2 == names[2] + 2          # two pack launches in the first forward, none afterwards
    for a, b in zip(p1, p2):
        assert (a.detach() - b.detach()).abs().max().item() <= 1e-6 * a.detach().abs().max().item()
    with torch.no_grad():
        c2.weight.mul_(2.0)                                    # torch edits the weight: version bump -> packs are rebuilt
    n0 = _lib.launch_count()
    y = fwd(c2, l2, x)
    assert _lib.launch_count() - n0 == names[1] + 1
    want = F.conv3d(F.conv3d(x.bfloat16().float(), c2.weight.bfloat16().float(), c2.bias, padding=1).bfloat16().float(),
                    l2.weight.bfloat16().float(), l2.bias)
    got = ops.from_tokens(y, 1, (6, 8, 8)).float()
    assert ((got - want).norm() / want.norm()).item() < 1e-2


def test_trainer_hooks_select_the_tensor_core_path():
    """nnUNetTrainer_NexToU.initialize() switches the process-wide CUDA autocast default to bf16, so that upstream's unchanged
    train_step (`with autocast("cuda")`) reaches the tcgen05 kernels; configure_optimizers() returns FusedSGD with upstream's
    hyper-parameters and a poly learning-rate schedule that the fused step follows."""
    import types
    from nextou_b200 import dense, trainers
    from nextou_b200.optim import FusedSGD
    old = torch.get_autocast_dtype("cuda")
    try:
        tr = trainers.nnUNetTrainer_NexToU(device="cuda")
        tr.initialize()
        assert torch.get_autocast_dtype("cuda") == torch.bfloat16
        tr.network = torch.nn.Conv3d(8, 16, 1).to(DEV)
        tr.initial_lr, tr.weight_decay, tr.num_epochs = 1e-2, 3e-5, 10
        opt, sched = tr.configure_optimizers()
        assert isinstance(opt, FusedSGD) and opt.param_groups[0]["momentum"] == 0.99 and opt.param_groups[0]["nesterov"]
        dense.stats.clear()
        x = torch.randn(1, 8, 4, 8, 8, device=DEV)
        with torch.autocast("cuda"):                       # what upstream's train_step opens
            y = dense.linear_tokens(__import__("nextou_b200").ops.as_tokens(x), tr.network)
        assert y.dtype == torch.bfloat16 and dense.stats["tcgen05.linear"] == 1
        y.float().mean().backward()
        opt.step()
        sched.step()
        assert abs(float(opt.param_groups[0]["lr"]) - 1e-2 * 0.9 ** 0.9) < 1e-9
        opt.step()
    finally:
        torch.set_autocast_dtype("cuda", old)
