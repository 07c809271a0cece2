"""GPU parity of the whole network (product modules on CUDA) against the CPU oracle and the reference goldens.

A randomly initialised NexToU is chaotic in its neighbour lists, so whole-network comparisons are teacher-forced:
the product's own kNN results are recorded per site, VERIFIED against the oracle's fp64 distances (every chosen
neighbour within 1e-4 of the true k-th distance) and replayed into the oracle forward.
"""
import os
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import torch_oracle as TO
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _record_graphs(model):
    from nextou_b200.blocks import PoolGrapher, SwinGrapher
    rec = []

    def hook(mod, inp, out):
        rec.append(mod.graph_conv.last_nn_idx.detach().long().cpu())

    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, (PoolGrapher, SwinGrapher))]
    return rec, hs


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("fname,cfg", [("model_mini3d_reference.npz", H.MINI3D), ("model_mini2d_reference.npz", H.MINI2D)],
                         ids=["mini3d", "mini2d"])
def test_model_fp32_forward_backward_vs_oracle(fname, cfg):
    npz = H.golden_model(fname)
    model = H.build_product(cfg)
    H.load_golden_into(model, npz)
    sd = H.full_state_dict_for_oracle(model)
    model = model.to(DEV).train()
    rec, hooks = _record_graphs(model)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    outs = model(x.to(DEV))
    loss = sum(o.float().mean() for o in outs)
    loss.backward()
    for h in hooks:
        h.remove()
    n_sites = 2 * (2 * 4 - 1)  # (pool + swin) x (4 encoder + 3 decoder GNN stages)
    assert len(rec) == n_sites

    for k in sd:
        if sd[k].dtype.is_floating_point:
            sd[k].requires_grad_(not k.endswith(("running_mean", "running_var", "relative_pos")))
    replay = TO.ReplayKnn(rec, tol=1e-4)
    want = TO.nextou_forward(sd, x, cfg["patch"], cfg["strides"], training=True, knn=replay)
    assert replay.pos == n_sites
    # Per-module parity is 1e-5 (test_every_module_fp32_vs_oracle).  End to end a few voxels differ more: besides the
    # neighbour lists the network has other discrete choices (arg-max of the 2x2x2 max-pool decides WHERE max-unpool
    # writes, ED:524-549), which flip on 1e-7 differences.  So: relative L2 error and a bound on the outlier fraction.
    for i, (a, b) in enumerate(zip(outs, want)):
        assert a.shape == b.shape
        diff = (a.detach().float().cpu() - b.detach()).abs()
        rel = (diff.norm() / b.detach().norm()).item()
        frac = (diff > 1e-3 * max(1.0, b.abs().max().item())).float().mean().item()
        assert rel <= 5e-4 and frac <= 1e-3, (i, rel, frac, diff.max().item())
    wl = sum(o.float().mean() for o in want)
    assert abs(loss.item() - wl.item()) < 1e-4
    wl.backward()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and not n.startswith("decoder.encoder.")]
    gmax = max(sd[n].grad.norm().item() for n, _ in named)
    worst = ("", 0.0)
    for name, p in named:
        if name.endswith(("conv.bias", "fc1.0.bias", "fc2.0.bias", "nn.0.bias")) and "seg_layers" not in name:
            continue  # bias in front of a batch / instance norm: the gradient is analytically zero (round-off only)
        gref = sd[name].grad
        assert p.grad is not None and gref is not None, name
        # relative L2 per parameter tensor; gradients that are analytically zero (a conv bias in front of a BatchNorm)
        # are pure round-off on both sides, so the denominator is floored at 1e-3 of the largest gradient norm
        rel = ((p.grad.float().cpu() - gref).norm() / max(gref.norm().item(), 1e-3 * gmax)).item()
        if rel > worst[1]:
            worst = (name, rel)
    assert worst[1] < 5e-2, worst  # end-to-end bound only; the strict gate is the per-module test below
    # the golden outputs of the real reference, where its graphs coincide with ours
    same_graphs = all(torch.equal(a, b) for a, b in zip(rec, H.golden_knn_list(npz)))
    if same_graphs:
        for i, o in enumerate(outs):
            ref = torch.from_numpy(npz[f"out/{i}"])
            got = o.detach().float().cpu()
            got = got if got.numel() <= 70000 else got.reshape(-1)[::97]
            assert torch.allclose(got, ref, rtol=1e-3, atol=2e-3)


@pytest.fixture
def no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False          # the fp32 route of the product calls cuDNN / cuBLAS: compare real fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _load_eval_statistics(model):
    """Running statistics of the reference's eval-mode fixture (momentum-1 forward of the golden network, generated by
    oracle/make_golden.py::gen_model_eval): with the default (0, 1) a random network amplifies its input ~3000x in eval."""
    ev = H.golden_model("model_mini3d_reference_eval.npz")
    model.load_state_dict({k[3:]: torch.from_numpy(ev[k]) for k in ev.files if k.startswith("sd/")}, strict=False)
    return ev


@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("fname,cfg", [("model_mini3d_reference.npz", H.MINI3D), ("model_mini2d_reference.npz", H.MINI2D)],
                         ids=["mini3d", "mini2d"])
def test_every_module_vs_oracle(fname, cfg, precision, mode, no_tf32):
    """Each conv stack / PoolGrapher / SwinGrapher / FFN of the network, fed the SAME input as the oracle.
    fp32: |err| <= 1e-5 * max(1, max|ref|) (north-star fp32 tolerance), kNN lists exactly optimal (excess <= 1e-6).
    bf16 autocast: bf16 has an 8-bit mantissa (eps 3.9e-3; the north star's 1e-3 is the fp16 figure) and every module
    ends in a batch / instance norm over as few as 24 tokens here, so: relative L2 <= 2e-2 against the fp32 oracle on
    the same (bf16-valued) input, with the product's own neighbour lists replayed.
    mode == "eval" is the inference path: BatchNorm with running statistics — under bf16 folded into
    the epilogue of the producing conv / GEMM — against the oracle's eval-mode blocks, which are pinned to the reference's
    own eval forward (tests/test_oracle_golden.py::test_torch_oracle_eval_forward_matches_reference_golden)."""
    from nextou_b200 import dense
    from nextou_b200.blocks import FFN, PoolGrapher, SwinGrapher
    from nextou_b200.conv_blocks import StackedConvBlocks
    training = mode == "train"
    if not training and "2d" in fname:
        pytest.skip("the eval-mode fixture of the reference exists for the 3-D configuration")
    npz = H.golden_model(fname)
    model = H.build_product(cfg)
    H.load_golden_into(model, npz)
    if not training:
        _load_eval_statistics(model)
    sd = H.full_state_dict_for_oracle(model)
    model = model.to(DEV).train(training)
    dense.stats.clear()
    dim = len(cfg["patch"])
    plan = TO.derive_plan(cfg["patch"], cfg["strides"])
    caps = []
    for name, mod in model.named_modules():
        if not name.startswith("decoder.encoder") and isinstance(mod, (PoolGrapher, SwinGrapher, FFN, StackedConvBlocks)):
            mod.register_forward_hook(lambda m, i, o, name=name: caps.append((name, m, i[0].detach(), o.detach(),
                                      getattr(getattr(m, "graph_conv", None), "last_nn_idx", None))))
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
        model(x.to(DEV))
    assert len(caps) >= 35
    for name, mod, inp, out, idx in caps:
        xin = inp.float().cpu()
        s = int(name.split(".")[2])
        level = s if name.startswith("encoder") else len(cfg["feats"]) - 2 - s
        st = plan["stages"][level]
        with torch.no_grad():
            if isinstance(mod, (PoolGrapher, SwinGrapher)):
                rk = TO.ReplayKnn([idx.long().cpu()], tol=1e-6, verify=precision == "fp32")
                fn = TO.pool_grapher if isinstance(mod, PoolGrapher) else TO.swin_grapher
                want = fn(xin, sd, name, dim, st, training, rk)
            elif isinstance(mod, FFN):
                want = TO.ffn(xin, sd, name, dim, training)
            else:
                want = xin
                gap = getattr(mod.convs[0].conv, "in_gap", None)
                if gap is not None and gap[1] > gap[0] and want.shape[1] == mod.convs[0].conv.in_channels + gap[1] - gap[0]:
                    # decoder input in the product's concat layout [up | zero gap | skip] (nextou_b200.dense.up_cat):
                    # the reference's torch.cat((up, skip), 1) is the same tensor without the gap channels
                    assert float(want[:, gap[0]:gap[1]].abs().max()) == 0.0
                    want = torch.cat([want[:, :gap[0]], want[:, gap[1]:]], 1)
                for i in range(len(mod.convs)):
                    want = TO._conv_block(want, sd, f"{name}.convs.{i}", dim, tuple(mod.convs[i].conv.stride), training)
        got = out.float().cpu()
        if precision == "fp32":
            err = (got - want).abs().max().item()
            assert err <= 1e-5 * max(1.0, want.abs().max().item()), (name, err)
        else:
            assert out.dtype == torch.bfloat16
            rel = ((got - want).norm() / want.norm()).item()
            # measured: 3e-3 .. 7e-3 for every module, except PoolGraphers that max-pool (ED:524): bf16 rounding creates
            # ties inside the 2x2x2 blocks, the arg-max (hence the max-unpool position, ED:549) may legitimately differ
            # from the fp32 oracle's, which moves a few features to a neighbouring voxel (measured 6e-2 .. 7e-2)
            pooled = isinstance(mod, PoolGrapher) and any(p > 1 for p in mod.pool_size)
            assert rel <= (0.12 if pooled else 2e-2), (name, rel)
    if not training and precision == "bf16":          # the whole inference forward ran without a statistics / normalisation pass
        assert dense.stats["tcgen05.conv_folded_norm"] >= 10 and dense.stats["tcgen05.linear_folded_norm"] >= 30
        assert dense.stats["native.batch_norm"] == 0


def test_whole_network_inference_forward_vs_oracle(no_tf32):
    """The inference forward nnU-Net's predictor calls (eval mode, deep supervision off; NexToU_Encoder_Decoder.py:324-337) of the
    WHOLE network against the oracle's eval forward with the product's neighbour lists replayed (each verified optimal under
    the oracle's own fp64 distances).  fp32: the eval network has no batch statistics to re-centre it, so the few max-unpool
    position flips (ED:524-549) spread further than in training — relative L2 <= 5e-3, <= 1 % of the voxels off by more than
    1e-3 of the logit range, and the predicted segmentation (arg-max over classes) agrees on >= 99.9 % of the voxels.
    (Under bf16 a randomly initialised eval network is chaotic end to end — the neighbour lists themselves change — so bf16
    inference parity is asserted module by module above, on identical inputs.)"""
    from nextou_b200.blocks import PoolGrapher, SwinGrapher
    cfg = H.MINI3D
    model = H.build_product(cfg)
    H.load_golden_into(model, H.golden_model("model_mini3d_reference.npz"))
    ev = _load_eval_statistics(model)
    sd = H.full_state_dict_for_oracle(model)
    model = model.to(DEV).eval()
    model.decoder.deep_supervision = False
    rec, hooks = _record_graphs(model)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    with torch.no_grad():
        y = model(x.to(DEV))
    assert tuple(y.shape) == tuple(int(v) for v in ev["out_shape"])
    replay = TO.ReplayKnn(rec, tol=1e-4)
    with torch.no_grad():
        want = TO.nextou_forward(sd, x, cfg["patch"], cfg["strides"], deep_supervision=False, training=False, knn=replay)
    assert replay.pos == 14
    got = y.float().cpu()
    diff = (got - want).abs()
    rel = (diff.norm() / want.norm()).item()
    frac = (diff > 1e-3 * max(1.0, want.abs().max().item())).float().mean().item()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    assert rel <= 5e-3 and frac <= 1e-2 and agree >= 0.999, (rel, frac, agree)
    if all(torch.equal(a, b) for a, b in zip(rec, H.golden_knn_list(ev))):      # the reference's own logits, where graphs coincide
        assert torch.allclose(got.reshape(-1)[::97], torch.from_numpy(ev["out/0"]), rtol=1e-3, atol=2e-3)


def test_model_running_stats_and_eval_mode():
    """BatchNorm running statistics are updated like nn.BatchNorm (momentum 0.1) and eval uses them."""
    cfg = H.MINI3D
    npz = H.golden_model("model_mini3d_reference.npz")
    model = H.build_product(cfg)
    H.load_golden_into(model, npz)
    sd0 = H.full_state_dict_for_oracle(model)
    model = model.to(DEV).train()
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    with torch.no_grad():
        model(x.to(DEV))
    bn = model.encoder.stages[0][0].convs[0].norm
    assert int(bn.num_batches_tracked) == 1
    with torch.no_grad():
        import torch.nn.functional as F
        c = F.conv3d(x, sd0["encoder.stages.0.0.convs.0.conv.weight"], sd0["encoder.stages.0.0.convs.0.conv.bias"], padding=(0, 1, 1))
        mean = c.mean((0, 2, 3, 4))
        var = c.var((0, 2, 3, 4), unbiased=True)
    assert torch.allclose(bn.running_mean.cpu(), 0.1 * mean, atol=1e-5)
    assert torch.allclose(bn.running_var.cpu(), 0.9 + 0.1 * var, atol=1e-5)
    model.eval()
    model.decoder.deep_supervision = False
    with torch.no_grad():
        y = model(x.to(DEV))
    assert tuple(y.shape) == (1, cfg["num_classes"], *cfg["patch"])


def test_full_size_3d_forward_backward_smoke():
    """3d_fullres_nextou 1 x 1 x 64 x 224 x 192 under bf16 autocast: shapes, finiteness, every trainable
    parameter receives a gradient, and all 14 kNN sites return valid (in-range, duplicate-free) lists."""
    cfg = H.FULL3D
    model = H.build_product(cfg).to(DEV).train()
    rec, hooks = _record_graphs(model)
    x = torch.randn(1, 1, *cfg["patch"], device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    want_shapes = [(1, 14, 64, 224, 192), (1, 14, 64, 112, 96), (1, 14, 32, 56, 48), (1, 14, 16, 28, 24), (1, 14, 8, 14, 12)]
    assert [tuple(o.shape) for o in outs] == want_shapes
    sum(o.float().mean() for o in outs).backward()
    assert all(torch.isfinite(o).all() for o in outs)
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing[:5]
    assert len(rec) == 14
    sizes = {(10752, 14): 168, (168, 7): 168, (10752, 28): 1344, (168, 14): 168, (1344, 32): 1344, (168, 32): 168, (168, 28): 168}
    for idx in rec:
        m = sizes[(idx.shape[1], idx.shape[2])]
        assert int(idx.min()) >= 0 and int(idx.max()) < m
        srt = idx.sort(-1).values
        assert bool((srt[..., 1:] != srt[..., :-1]).all())


def test_inference_forward_folds_batch_norm():
    """Eval-mode forward under bf16 autocast (what nnU-Net's sliding-window predictor calls, deep supervision off): every
    BatchNorm is folded into the producing conv / GEMM epilogue (no statistics / normalisation pass).  A randomly
    initialised NexToU is chaotic end to end in bf16 (neighbour lists and max-unpool positions flip on rounding, see
    DESIGN.md 5), so numerical parity is asserted where it is well defined: each conv stack against the oracle's eval-mode
    blocks on the SAME input (per-layer folding parity: test_gpu_gemm.py::test_folded_eval_norm_matches_fp32_reference)."""
    from nextou_b200 import dense
    from nextou_b200.conv_blocks import StackedConvBlocks
    cfg = H.MINI3D
    npz = H.golden_model("model_mini3d_reference.npz")
    model = H.build_product(cfg)
    H.load_golden_into(model, npz)
    model = model.to(DEV).train()
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    # give the running statistics the batch statistics of this input (momentum 1), as a trained network would have: with
    # the defaults (0, 1) a random network amplifies its input ~3000x in eval mode
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = 1.0
    with torch.no_grad():
        model(x.to(DEV))
    sd = H.full_state_dict_for_oracle(model)
    model.eval()
    model.decoder.deep_supervision = False
    caps = []
    for name, mod in model.named_modules():
        if name.startswith("encoder.stages") and isinstance(mod, StackedConvBlocks):
            mod.register_forward_hook(lambda m, i, o, name=name: caps.append((name, m, i[0].detach(), o.detach())))
    dense.stats.clear()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y = model(x.to(DEV))
    assert tuple(y.shape) == (1, cfg["num_classes"], *cfg["patch"]) and bool(torch.isfinite(y).all())
    assert dense.stats["tcgen05.conv_folded_norm"] >= 10 and dense.stats["tcgen05.linear_folded_norm"] >= 30
    assert dense.stats["native.batch_norm"] == 0                     # no statistics / normalisation pass for BatchNorm
    dim = len(cfg["patch"])
    assert len(caps) >= 4
    for name, mod, inp, out in caps:
        want = inp.float().cpu()
        with torch.no_grad():
            for i in range(len(mod.convs)):
                want = TO._conv_block(want, sd, f"{name}.convs.{i}", dim, tuple(mod.convs[i].conv.stride), False)
        rel = ((out.float().cpu() - want).norm() / want.norm()).item()
        assert rel <= 1e-2, (name, rel)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_plain_grapher_api_smoke(precision):
    """The reference's plain `Grapher` (ED:553-632) is never built by NexToU but is part of the module surface: it must run
    forward + backward (fc1 -> kNN graph conv -> fc2 + residual) and return finite values of the input's shape."""
    from nextou_b200.blocks import Grapher
    torch.manual_seed(0)
    m = Grapher(12, kernel_size=5, dilation=1, conv="mr", act="leakyrelu", norm="batch", n=64).to(DEV).train()
    x = torch.randn(2, 12, 4, 4, 4, device=DEV, requires_grad=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
        y = m(x)
    assert y.shape == x.shape
    y.float().square().mean().backward()
    assert bool(torch.isfinite(y).all()) and x.grad is not None and bool(torch.isfinite(x.grad).all())
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters() if p.requires_grad)


# ------------------------------------------------------------------------------------------------------
# BASELINE.json config 3: Swin-GNN block in isolation — 132 channels (nearest constructible width to 128), 28 x 28 x 28 tokens,
# 7 x 7 x 7 windows shifted by 3 (64 windows x 343 tokens), k = 9, dilation 2 (top-18, every second), MRConv, relative
# position table (1, 343, 343) — and the global 21 952-token kNN graph that takes the reference's 10 000-row chunked path
# (torch_edge.py:70-82).  SURVEY.md 8d.
# ------------------------------------------------------------------------------------------------------
def _config3_block(seed=0):
    from nextou_b200.blocks import SwinGrapher
    torch.manual_seed(seed)
    return SwinGrapher(132, (28, 28, 28), kernel_size=9, dilation=2, conv="mr", act="leakyrelu", norm="instance", bias=True,
                       stochastic=False, epsilon=0.2, r=1, n=343, relative_pos=True, conv_op=torch.nn.Conv3d,
                       norm_op=torch.nn.BatchNorm3d, norm_op_kwargs={"eps": 1e-5, "affine": True},
                       window_size=(7, 7, 7), shift_size=[3, 3, 3])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_config3_swin_gnn_block_vs_oracle(precision):
    """SwinGrapher at the config-3 shape vs the torch oracle's swin_grapher on the same input and parameters: fp32 1e-5 with the
    neighbour lists verified optimal under the oracle's fp64 distances; bf16 relative L2 2e-2 with the product's lists replayed."""
    m = _config3_block().to(DEV).train()
    assert tuple(m.relative_pos.shape) == (1, 343, 343)
    sd = {"blk." + k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
    st = dict(window=(7, 7, 7), shift=[3, 3, 3], swin_k=9, dilation=2)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 132, 28, 28, 28, generator=g)
    if precision == "bf16":
        x = x.bfloat16().float()
    xg = x.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
        y = m(xg)
    idx = m.graph_conv.last_nn_idx.long().cpu()                       # (64 windows, 343, 9): the dilated lists
    assert tuple(idx.shape) == (64, 343, 9)
    y.float().square().mean().backward()
    assert xg.grad is not None and bool(torch.isfinite(xg.grad).all())

    class Replay:                                                      # replays the dilated lists, checks them at k * d = 18
        def __call__(self, h4, y4, relpos, k, dilation=1):
            assert (k, dilation) == (9, 2) and tuple(h4.shape) == (64, 132, 343, 1)
            if precision == "fp32":
                d = TO.knn_distances(h4, None, relpos)
                full = d.topk(18, dim=-1, largest=False).values
                chosen = torch.gather(d, 2, idx)
                # neighbour j of the dilated list must be the (2j + 1)-th nearest under the oracle's distances (up to ties)
                assert (chosen - full[..., ::2]).abs().max().item() <= 1e-4
            return idx
    with torch.no_grad():
        want = TO.swin_grapher(x, sd, "blk", 3, st, True, knn=Replay())
    got = y.detach().float().cpu()
    if precision == "fp32":
        assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    else:
        assert ((got - want).norm() / want.norm()).item() <= 2e-2


def test_config3_global_knn_21952_tokens_matches_oracle():
    """DenseDilatedKnnGraph(9, 2) on all 21 952 tokens at once (the reference splits this into 10 000-row chunks, TE:70-82; the
    kernel needs no chunking).  Bit-exact against the C oracle on 768 query rows spread over the three reference chunks, and
    chunk-invariant: the list of a query row does not depend on which rows are queried together with it."""
    from nextou_b200 import ops
    from nextou_b200.graph import DenseDilatedKnnGraph
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 132, 21952, 1, generator=g)
    mod = DenseDilatedKnnGraph(9, 2, stochastic=False).to(DEV)
    edge = mod(x.to(DEV))
    assert tuple(edge.shape) == (2, 1, 21952, 9) and edge.dtype == torch.int64
    assert torch.equal(edge[1, 0, :, 0].cpu(), torch.arange(21952))
    tok = x[0, :, :, 0].t().contiguous()                                             # [21952, 132]
    rows = torch.cat([torch.arange(0, 256), torch.arange(9900, 10156), torch.arange(21696, 21952)])
    want = c_oracle.knn_graph(tok[rows][None].numpy(), tok[None].numpy(), None, 18, 1)[0][:, ::2]     # xy graph of the sampled queries
    assert np.array_equal(edge[0, 0][rows].cpu().numpy(), want)
    full18, _ = ops.knn_graph(tok.to(DEV), 1, 21952, k=18)
    part18, _ = ops.knn_graph(tok[9000:12000].to(DEV), 1, 3000, tok.to(DEV), 21952, k=18)
    assert torch.equal(full18[0, 9000:12000], part18[0])


def test_deferred_wgrad_join_is_race_free_on_the_full_size_network():
    """native._complete_wgrad leaves the weight-gradient kernels of a backward pass on the side stream until the pass ends.
    tools/defer_check.py runs the 3d_fullres network (kernels long enough for a missing dependency to show): gradients with
    the deferred join == gradients with a join per layer (up to the split-K atomics), and every deferred gradient becomes
    param.grad without a main-stream copy."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "defer_check.py"), "3"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DEFER CHECK PASSED" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])
