"""GPU parity tests of the CUDA kernels (through the C-ABI) against the CPU oracle and the reference goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import torch_oracle as TO
from oracle.make_golden import BTI_CASES, bti_case, bti_interactions
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ------------------------------------------------------------------------------------------------
# kNN: bit-exact vs the plain-C oracle
# ------------------------------------------------------------------------------------------------
def _gpu_knn(x, y, relpos, k, d, dtype=torch.float32, normalize=True):
    from nextou_b200 import ops
    B, N, C = x.shape
    xt = x.to(DEV, dtype).reshape(B * N, C)
    rp = None if relpos is None else relpos.to(DEV)
    if y is None:
        idx, idx32 = ops.knn_graph(xt, B, N, relpos=rp, k=k, dilation=d, normalize=normalize)
    else:
        M = y.shape[1]
        idx, idx32 = ops.knn_graph(xt, B, N, y.to(DEV, dtype).reshape(B * M, C), M, relpos=rp, k=k, dilation=d,
                                   normalize=normalize)
    assert idx.dtype == torch.int64 and idx32.dtype == torch.int32
    assert torch.equal(idx, idx32.long())
    return idx.cpu().numpy()


@pytest.mark.parametrize("case", H.KNN_CASES, ids=[c[0] for c in H.KNN_CASES])
def test_knn_bit_exact_vs_c_oracle(case):
    name, B, N, M, C, k, d, rp = case
    x, y, relpos = H.knn_inputs(case)
    want = c_oracle.knn_graph(x.numpy(), None if y is None else y.numpy(), None if relpos is None else relpos[0].numpy(), k, d)
    got = _gpu_knn(x, y, relpos, k, d)
    assert np.array_equal(got, want)


def test_knn_reference_golden_tie_aware():
    """Against the unmodified reference's own index tensors (tests/golden/knn_reference.npz)."""
    gold = np.load(os.path.join(H.GOLDEN, "knn_reference.npz"))
    for case in H.KNN_CASES:
        name, B, N, M, C, k, d, rp = case
        x, y, relpos = H.knn_inputs(case)
        got = _gpu_knn(x, y, relpos, k, d)
        ref = gold[name].astype(np.int64)
        x4 = x.permute(0, 2, 1).unsqueeze(-1)
        y4 = None if y is None else y.permute(0, 2, 1).unsqueeze(-1)
        dist = TO.knn_distances(x4, y4, relpos)
        for b, i in zip(*np.nonzero((got != ref).any(-1))):
            dm = dist[b, i, torch.from_numpy(got[b, i])]
            dr = dist[b, i, torch.from_numpy(ref[b, i])]
            assert torch.allclose(dm, dr, rtol=0, atol=1e-5), (name, b, i)


@pytest.mark.parametrize("shape", [(1, 10752, 1344, 264, 28), (1, 10752, 168, 132, 14), (64, 168, 0, 264, 14),
                                   (128, 168, 0, 132, 7), (2, 1344, 1344, 324, 32)],
                         ids=["pool_s3_full", "pool_s2_full", "swin_s3_full", "swin_s2_128win", "pool_s4_x2"])
def test_knn_full_size_sites_bit_exact(shape):
    B, N, M, C, k = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, N, C, generator=g)
    y = torch.randn(B, M, C, generator=g) if M else None
    relpos = 0.1 * torch.randn(1, N, M or N, generator=g)
    want = c_oracle.knn_graph(x.numpy(), None if y is None else y.numpy(), relpos[0].numpy(), k, 1)
    assert np.array_equal(_gpu_knn(x, y, relpos, k, 1), want)


@pytest.mark.parametrize("distinct", [1, 3, 40], ids=["all_equal", "3_distinct", "40_distinct"])
@pytest.mark.parametrize("shape", [(4, 168, 0, 7, 1), (2, 168, 0, 9, 2), (1, 512, 1344, 28, 1)], ids=["swin_k7", "swin_k9_d2", "pool_k28"])
def test_knn_massive_distance_ties_bit_exact(shape, distinct):
    """Duplicated tokens (background of a CT volume): far more than KNN_CAP candidates tie with the K-th best distance, so the
    kernel takes its warp-level slow path (and the running-list merge across candidate chunks); ties go to the lowest index."""
    B, N, M, k, d = shape
    g = torch.Generator().manual_seed(11)
    C = 48
    protos = torch.randn(distinct, C, generator=g)
    x = protos[torch.randint(0, distinct, (B, N), generator=g)]
    y = protos[torch.randint(0, distinct, (B, M), generator=g)] if M else None
    want = c_oracle.knn_graph(x.numpy(), None if y is None else y.numpy(), None, k, d)
    assert np.array_equal(_gpu_knn(x, y, None, k, d), want)


def test_knn_bf16_input_unnormalized_and_row_map():
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, N, C, k = 4, 96, 48, 6
    x = torch.randn(B, N, C, generator=g)
    xb = x.bfloat16()
    want = c_oracle.knn_graph(xb.float().numpy(), None, None, k, 1)
    assert np.array_equal(_gpu_knn(xb, None, None, k, 1, dtype=torch.bfloat16), want)
    want_raw = c_oracle.knn_graph(x.numpy(), None, None, k, 1, normalize=False)
    assert np.array_equal(_gpu_knn(x, None, None, k, 1, normalize=False), want_raw)
    # gather rows through a permutation map
    perm = torch.randperm(B * N, generator=g)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(B * N)
    shuffled = x.reshape(B * N, C)[inv]             # shuffled[perm[r]] = x_flat[r]
    idx, _ = ops.knn_graph(shuffled.to(DEV), B, N, k=k, x_row_map=perm.to(DEV, torch.int32))
    assert np.array_equal(idx.cpu().numpy(), c_oracle.knn_graph(x.numpy(), None, None, k, 1))


def test_knn_rejects_bad_arguments():
    from nextou_b200 import ops
    from nextou_b200._lib import NextouError
    x = torch.randn(64, 12, device=DEV)
    with pytest.raises(NextouError):
        ops.knn_graph(x, 1, 64, k=33)            # k*dilation > 32
    with pytest.raises(NextouError):
        ops.knn_graph(x, 4, 16, k=9, dilation=2)  # k*dilation > M
    with pytest.raises(NextouError):
        ops.knn_graph(x, 1, 64, k=4, relpos=torch.zeros(1, 64, 32, device=DEV))


def test_dense_dilated_knn_graph_module_api():
    """Drop-in module: (B, C, N, 1) in, (2, B, N, k) int64 edge_index out (TE:139-163)."""
    from nextou_b200.graph import DenseDilatedKnnGraph, dense_knn_matrix, xy_dense_knn_matrix
    case = H.KNN_CASES[1]
    name, B, N, M, C, k, d, rp = case
    x, y, relpos = H.knn_inputs(case)
    x4 = x.permute(0, 2, 1).unsqueeze(-1).contiguous().to(DEV)
    y4 = y.permute(0, 2, 1).unsqueeze(-1).contiguous().to(DEV)
    mod = DenseDilatedKnnGraph(k, d, stochastic=False, epsilon=0.0).to(DEV)
    e = mod(x4, y4, relpos.to(DEV))
    assert e.shape == (2, B, N, k) and e.dtype == torch.int64
    want = c_oracle.knn_graph(x.numpy(), y.numpy(), relpos[0].numpy(), k, d)
    assert np.array_equal(e[0].cpu().numpy(), want)
    assert torch.equal(e[1].cpu(), torch.arange(N).view(1, N, 1).expand(B, N, k))
    # un-normalised helpers (TE:58-110)
    e2 = xy_dense_knn_matrix(x4, y4, k, relpos.to(DEV))
    assert np.array_equal(e2[0].cpu().numpy(), c_oracle.knn_graph(x.numpy(), y.numpy(), relpos[0].numpy(), k, 1, normalize=False))
    e3 = dense_knn_matrix(x4, k)
    assert np.array_equal(e3[0].cpu().numpy(), c_oracle.knn_graph(x.numpy(), None, None, k, 1, normalize=False))
    # stochastic branch consumes the host RNG exactly like the reference (TE:128-130)
    torch.manual_seed(11)
    smod = DenseDilatedKnnGraph(k, d, stochastic=True, epsilon=0.2).to(DEV).train()
    smod(x4, y4, relpos.to(DEV))
    after = torch.rand(1)
    torch.manual_seed(11)
    if torch.rand(1) < 0.2:
        torch.randperm(k * d)
    assert torch.equal(after, torch.rand(1))


# ------------------------------------------------------------------------------------------------
# message passing
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 0.0), (torch.bfloat16, 0.0)], ids=["f32", "bf16"])
@pytest.mark.parametrize("self_graph", [True, False], ids=["self", "xy"])
def test_mrconv_gather_forward_backward(dtype, tol, self_graph):
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, N, M, C, k = 3, 168, (168 if self_graph else 40), 132, 7
    x = torch.randn(B, N, C, generator=g).to(dtype)
    y = None if self_graph else torch.randn(B, M, C, generator=g).to(dtype)
    idx = torch.randint(0, M, (B, N, k), generator=g)
    xg = x.to(DEV).reshape(B * N, C).requires_grad_(True)
    yg = None if y is None else y.to(DEV).reshape(B * M, C).requires_grad_(True)
    out = ops.mrconv_gather(xg, idx.to(DEV, torch.int32), N, M, y_tok=yg)
    # oracle in the reference layout (B, C, N, 1)
    xo = x.float().permute(0, 2, 1).unsqueeze(-1).clone().requires_grad_(True)
    yo = None if y is None else y.float().permute(0, 2, 1).unsqueeze(-1).clone().requires_grad_(True)
    want = TO.max_relative(xo, idx, yo)                                     # (B, 2C, N, 1)
    want_tok = want.squeeze(-1).permute(0, 2, 1).reshape(B * N, 2 * C)
    assert out.dtype == dtype
    assert torch.equal(out.float().cpu(), want_tok.to(dtype).float())       # exact (max of rounded == rounded max)
    w = torch.randn(B * N, 2 * C, generator=g)
    (out.float() * w.to(DEV)).sum().backward()
    (want_tok * w).sum().backward()
    gx = xo.grad.squeeze(-1).permute(0, 2, 1).reshape(B * N, C)
    atol = 1e-5 if dtype == torch.float32 else 0.15
    if dtype == torch.float32:
        assert torch.allclose(xg.grad.cpu(), gx, rtol=1e-5, atol=atol)
        if yo is not None:
            gy = yo.grad.squeeze(-1).permute(0, 2, 1).reshape(B * M, C)
            assert torch.allclose(yg.grad.cpu(), gy, rtol=1e-5, atol=atol)
    else:  # bf16: ties in the max are frequent and may route the gradient to another (equally maximal) neighbour;
        # the total is conserved: sum(dx) + sum(dy) == sum of the upstream gradient over the x-channels
        tot = xg.grad.float().sum() + (0 if yg is None else yg.grad.float().sum())
        assert abs(tot.item() - w[:, 0::2].sum().item()) < 8.0


def test_mrconv_gather_row_maps_equal_window_partition():
    """Shifted windows through row maps == roll + partition + gather + reverse + roll of the reference."""
    from nextou_b200 import ops
    from nextou_b200.blocks import shifted_window_row_map
    g = torch.Generator().manual_seed(2)
    B, C, S, ws, sh, k = 2, 12, (4, 6, 8), (2, 3, 4), (1, 1, 2), 3
    x = torch.randn(B, C, *S, generator=g)
    n = int(np.prod(ws))
    nW = B * int(np.prod(S)) // n
    idx = torch.randint(0, n, (nW, n, k), generator=g)
    rolled = torch.roll(x, shifts=tuple(-s for s in sh), dims=(2, 3, 4))
    win = TO._windows(rolled, ws).reshape(nW, C, n, 1)
    want_w = TO.max_relative(win, idx).reshape(nW, 2 * C, *ws)
    want = torch.roll(TO._unwindows(want_w, ws, B, S), shifts=sh, dims=(2, 3, 4))
    tok = ops.as_tokens(x.to(DEV))
    rm = shifted_window_row_map(B, S, ws, sh, DEV)
    out = ops.mrconv_gather(tok, idx.to(DEV, torch.int32), n, n, q_row_map=rm, y_row_map=rm)
    got = ops.from_tokens(out, B, S)
    assert torch.equal(got.cpu(), want)


# ------------------------------------------------------------------------------------------------
# pooling
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("spatial,pool", [((8, 12, 10), (2, 2, 2)), ((6, 9, 8), (1, 3, 2)), ((12, 16), (2, 2)), ((8, 8, 8), (4, 4, 4))])
def test_pool_unpool_forward_backward(dtype, spatial, pool):
    import torch.nn.functional as F
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, C = 2, 12
    dim = len(spatial)
    x = torch.randn(B, C, *spatial, generator=g).to(dtype)
    mp, ap, up = (F.max_pool3d, F.avg_pool3d, F.max_unpool3d) if dim == 3 else (F.max_pool2d, F.avg_pool2d, F.max_unpool2d)
    xo = x.float().clone().requires_grad_(True)
    xg = x.clone().to(DEV).requires_grad_(True)
    tok = ops.as_tokens(xg)
    pooled_spatial = tuple(s // p for s, p in zip(spatial, pool))
    # max pool
    q, arg = ops.maxpool_tokens(tok, B, spatial, pool)
    qo, ind = mp(xo, pool, pool, return_indices=True)
    assert torch.equal(ops.from_tokens(q, B, pooled_spatial).float().cpu(), qo.to(dtype).float())
    # avg pool
    a = ops.avgpool_tokens(tok, B, spatial, pool)
    ao = ap(xo, pool, pool)
    assert torch.allclose(ops.from_tokens(a, B, pooled_spatial).float().cpu(), ao.to(dtype).float(), rtol=1e-6, atol=1e-6)
    # unpool of a 2C tensor with the duplicated indices (ED:536-549)
    f = torch.randn(B, 2 * C, *pooled_spatial, generator=g).to(dtype)
    fo = f.float().clone().requires_grad_(True)
    fg = f.clone().to(DEV).requires_grad_(True)
    u = ops.maxunpool_tokens(ops.as_tokens(fg), arg, B, spatial, pool)
    uo = up(fo, torch.cat((ind, ind), 1), pool, pool)
    assert torch.equal(ops.from_tokens(u, B, spatial).float().cpu(), uo.to(dtype).float())
    # backward of everything at once
    w1 = torch.randn(qo.shape, generator=g)
    w2 = torch.randn(ao.shape, generator=g)
    w3 = torch.randn(uo.shape, generator=g)
    ((qo * w1).sum() + (ao * w2).sum() + (uo * w3).sum()).backward()
    loss = (ops.from_tokens(q, B, pooled_spatial).float() * w1.to(DEV)).sum() \
        + (ops.from_tokens(a, B, pooled_spatial).float() * w2.to(DEV)).sum() \
        + (ops.from_tokens(u, B, spatial).float() * w3.to(DEV)).sum()
    loss.backward()
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    assert torch.allclose(xg.grad.float().cpu(), xo.grad, **tol)
    assert torch.allclose(fg.grad.float().cpu(), fo.grad, **tol)


# ------------------------------------------------------------------------------------------------
# batch / instance norm + LeakyReLU
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("C,rows,instances,slope", [(33, 1000, 1, 0.01), (66, 777, 1, None), (132, 4099, 1, 0.01),
                                                   (324, 168, 1, 0.01), (1296, 50, 1, 0.01), (264, 96, 2, 0.01),
                                                   (40, 3, 1, 0.01), (33, 200003, 1, 0.01)])
def test_norm_act_forward_backward(dtype, C, rows, instances, slope):
    import torch.nn.functional as F
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(C + rows)
    x = (torch.randn(instances * rows, C, generator=g) * 2 + 0.5).to(dtype)
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g)
    w = torch.randn(instances * rows, C, generator=g)
    rm, rv = torch.zeros(C), torch.ones(C)
    # oracle: torch CPU fp32 on the (possibly bf16-valued) input, reference layout (B, C, N)
    xo = x.float().reshape(instances, rows, C).permute(0, 2, 1).clone().requires_grad_(True)
    go, bo = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    if instances == 1:
        rmo, rvo = rm.clone(), rv.clone()
        yo = F.batch_norm(xo, rmo, rvo, go, bo, training=True, momentum=0.1, eps=1e-5)
    else:
        yo = F.instance_norm(xo, weight=go, bias=bo, eps=1e-5)
    if slope is not None:
        yo = F.leaky_relu(yo, slope)
    wo = w.reshape(instances, rows, C).permute(0, 2, 1)
    (yo * wo).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    gg, bg = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    rmg, rvg = (rm.to(DEV), rv.to(DEV)) if instances == 1 else (None, None)
    y = ops.norm_act_tokens(xg, gg, bg, rmg, rvg, 0.1, 1e-5, 1.0 if slope is None else slope, instances)
    assert y.dtype == dtype and y.shape == x.shape
    (y.float() * w.to(DEV)).sum().backward()
    want = yo.detach().permute(0, 2, 1).reshape(instances * rows, C)
    gx = xo.grad.permute(0, 2, 1).reshape(instances * rows, C)
    if dtype == torch.float32:
        assert torch.allclose(y.cpu(), want, rtol=1e-5, atol=2e-5)
        assert torch.allclose(xg.grad.cpu(), gx, rtol=1e-4, atol=1e-4 * max(1.0, gx.abs().max().item()))
        assert torch.allclose(gg.grad.cpu(), go.grad, rtol=1e-4, atol=1e-3)
        assert torch.allclose(bg.grad.cpu(), bo.grad, rtol=1e-4, atol=1e-3)
        if instances == 1:
            assert torch.allclose(rmg.cpu(), rmo, rtol=1e-5, atol=1e-6)
            assert torch.allclose(rvg.cpu(), rvo, rtol=1e-5, atol=1e-6)
    else:
        assert torch.allclose(y.float().cpu(), want, rtol=1e-2, atol=2e-2)
        assert ((xg.grad.float().cpu() - gx).norm() / gx.norm()).item() < 2e-2
        assert ((gg.grad.cpu() - go.grad).norm() / go.grad.norm()).item() < 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("C,rows", [(36, 1536), (33, 5000), (40, 1536), (324, 168)])
def test_eval_mode_affine_act_on_channel_padded_rows(dtype, C, rows):
    """Eval-mode BatchNorm (+ LeakyReLU) as a per-channel affine map on rows whose pitch is padded to 8 channels (the scale /
    shift vectors are padded to the pitch: regression test for a freed-temporary bug that aliased the two vectors)."""
    import torch.nn.functional as F
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(C + rows)
    x = torch.randn(rows, ops.pad8(C), generator=g).to(dtype).to(DEV)[:, :C]
    scale, shift = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
    y = ops.affine_act_tokens(x, scale, shift, 0.01)
    want = F.leaky_relu(x.float() * scale + shift, 0.01)
    tol = 1e-6 if dtype == torch.float32 else 2 ** -8
    assert (y.float() - want).abs().max().item() <= tol * want.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("rows,C", [(5000, 33), (777, 66), (64, 132), (3, 8)])
def test_rows_copy_add_between_different_pitches(dtype, rows, C):
    """Skip half of the concat buffer / fan-out gradient sum: views with a column offset and different pitches."""
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(rows + C)
    P = ops.pad8(C)
    big = torch.randn(rows, 2 * P, generator=g).to(dtype).to(DEV)         # [up | gap | skip] rows
    small = torch.randn(rows, P, generator=g).to(dtype).to(DEV)
    a, b = big[:, P:P + C], small[:, :C]
    out = ops.padded_like(rows, C, dtype, DEV)
    assert ops.rows_copy_add(a, b, out)
    assert torch.equal(out, a + b)
    dst = torch.zeros(rows, 2 * P, device=DEV, dtype=dtype)
    assert ops.rows_copy_add(b, None, dst[:, P:P + C])
    assert torch.equal(dst[:, P:P + C], b) and torch.count_nonzero(dst[:, :P]) == 0
    # not vector-addressable (odd pitch): refused, the caller falls back to ATen
    odd = torch.randn(rows, C + 1, generator=g).to(dtype).to(DEV)[:, :C]
    assert not ops.rows_copy_add(odd, None, out) or (C + 1) % 8 == 0


# ------------------------------------------------------------------------------------------------
# BTI / TI loss
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def bti_gold():
    return np.load(os.path.join(H.GOLDEN, "bti_reference.npz"))


@pytest.mark.parametrize("case", BTI_CASES, ids=[c[0] for c in BTI_CASES])
@pytest.mark.parametrize("layout", ["ncdhw", "channels_last"])
def test_bti_loss_matches_reference_golden_and_oracle(case, layout, bti_gold):
    from nextou_b200 import ops
    from nextou_b200.losses import BTI_Loss, TI_Loss
    name, shape, nc, seed, conn, thick, kind = case
    logits, target = bti_case(shape, nc, seed)
    inc, exc = bti_interactions(kind, nc)
    dim = len(shape) - 1
    lg = logits.to(DEV)
    if layout == "channels_last":
        lg = ops.channels_last(lg)
    lg.requires_grad_(True)
    mod = BTI_Loss(dim=dim, connectivity=conn, inclusion=inc, exclusion=exc, min_thick=thick)
    val = mod(lg, target.to(DEV))
    assert val.dtype == torch.float64 and val.dim() == 0
    ref = float(bti_gold[f"{name}.bti.loss"])
    assert abs(val.item() - ref) <= 1e-9 * max(1.0, abs(ref))               # fp64 scalar vs the reference
    labels = ops.bti_labels(lg.detach())
    assert np.array_equal(labels.cpu().numpy(), torch.softmax(logits, 1).argmax(1).numpy().astype(np.uint8))   # BTI:132-134
    crit = mod.binary_topological_interaction_module(labels.unsqueeze(1))
    ref_crit = np.unpackbits(bti_gold[f"{name}.bti.crit"])[: labels.numel()].reshape(labels.shape)
    assert np.array_equal(crit[:, 0].cpu().numpy().astype(np.uint8), ref_crit)  # bit-exact critical map
    (val * 3.0).backward()
    gr = lg.grad.double().cpu() / 3.0
    assert abs(gr.sum().item() - float(bti_gold[f"{name}.bti.grad_sum"])) < 1e-3
    assert abs(gr.abs().sum().item() - float(bti_gold[f"{name}.bti.grad_abs_sum"])) < 1e-4 * float(bti_gold[f"{name}.bti.grad_abs_sum"]) + 1e-6
    probe = gr.reshape(-1)[:: max(1, gr.numel() // 4096)].float().numpy()
    assert np.allclose(probe, bti_gold[f"{name}.bti.grad_probe"], rtol=1e-5, atol=1e-6)
    if kind == "pairs":
        ti = TI_Loss(dim=dim, connectivity=conn, inclusion=inc, exclusion=exc, min_thick=thick)
        tv = ti(logits.to(DEV), target.to(DEV))
        assert abs(tv.item() - float(bti_gold[f"{name}.ti.loss"])) <= 1e-9 * max(1.0, abs(ref))


def test_bti_full_size_properties():
    """BASELINE config 4 size (14 x 64 x 224 x 192): map vs the C oracle bit-exact, loss vs the torch oracle, plus
    size-independent properties: swapping A and C leaves the map unchanged; a constant label volume has no
    critical voxel; the loss is linear in the batch mean."""
    from nextou_b200 import ops
    from nextou_b200.losses import BTI_Loss
    from oracle.ref_shims import SYNAPSE_EXCLUSION, make_tensors
    exc = make_tensors(SYNAPSE_EXCLUSION)
    shape = (1, 64, 224, 192)
    logits, target = bti_case(shape, 14, 21)
    mod = BTI_Loss(dim=3, connectivity=26, inclusion=[], exclusion=exc, min_thick=1)
    lg = logits.to(DEV).bfloat16()
    val = mod(lg, target.to(DEV))
    labels = ops.bti_labels(lg)
    lab_cpu = torch.softmax(lg.float().cpu(), 1).argmax(1)      # BTI:132-134
    assert np.array_equal(labels.cpu().numpy(), lab_cpu.numpy().astype(np.uint8))
    ma, mc, flags = mod.interaction_table()
    crit = ops.bti_critical_map(labels, ma, mc, flags, 26, 1)
    want = c_oracle.bti_critical(lab_cpu.numpy().astype(np.uint8), ma, mc, flags, 26, 1)
    assert np.array_equal(crit.cpu().numpy(), want)
    ref = TO.bti_loss(lg.float().cpu(), target, [], exc, 3, 26, 1)
    assert abs(val.item() - ref.item()) <= 1e-9 * abs(ref.item())
    swapped = ops.bti_critical_map(labels, mc, ma, flags, 26, 1)
    assert torch.equal(swapped, crit)
    const = ops.bti_critical_map(torch.full_like(labels, 3), ma, mc, flags, 26, 1)
    assert int(const.sum()) == 0


def _near_tie_logits(shape, nc, seed, scale=0.05):
    """fp32 logits whose two largest classes per voxel are closer than fp32 softmax can resolve: gaps of 0, 1 or 2 ulp at
    |x| ~ 0.15 (0 / 1.5e-8 / 3e-8 <= 2^-25), with the LARGER logit at the HIGHER class index half of the time, so that
    argmax(logits) != argmax(softmax(logits)) there (the reference takes the latter, bti_loss.py:132-134)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape[0], nc, *shape[1:], generator=g) * scale
    mx = x.max(1, keepdim=True).values
    vs = (shape[0], 1, *shape[1:])
    i1 = torch.randint(0, nc, vs, generator=g)
    i2 = (i1 + torch.randint(1, nc, vs, generator=g)) % nc
    top = mx + 0.01
    ulps = torch.randint(0, 3, vs, generator=g)
    hi = top.clone()
    for _ in range(2):
        hi = torch.where(ulps > 0, torch.nextafter(hi, torch.full_like(hi, 10.0)), hi)
        ulps = ulps - 1
    x.scatter_(1, i1, top)
    x.scatter_(1, i2, hi)
    return x


@pytest.mark.parametrize("fused", [False, True], ids=["bti_kernel", "dsloss_kernel"])
def test_bti_labels_are_argmax_of_softmax_on_near_ties(fused):
    """The discrete map of the reference is argmax(softmax(x)) (bti_loss.py:132-134).  fp32 softmax maps logits within
    2^-25 of the maximum to the same probability and torch.argmax then returns the lowest index — which is NOT
    argmax(x).  Inside that clear-tie zone the result does not depend on the exp / sum implementation, and the kernels
    must equal torch bit for bit.  (One to four ulp further out, torch's own CPU kernels disagree with each other —
    softmax over dim 1 of NCDHW vs. the last dim of channels-last gave different arg-maxima on 4 % of crafted voxels —
    so no implementation-independent answer exists there; the kernels evaluate exp(x - max) / sum in class order, fp32.)"""
    from nextou_b200 import ops
    shape, nc = (2, 6, 40, 48), 14
    x = _near_tie_logits(shape, nc, 11)
    want = torch.softmax(x, 1).argmax(1)
    naive = x.argmax(1)
    assert (want != naive).float().mean().item() > 0.15           # the case really separates the two definitions
    target = torch.randint(0, nc, (shape[0], 1, *shape[1:])).float()
    for layout in ("ncdhw", "channels_last"):
        lg = x.to(DEV)
        if layout == "channels_last":
            lg = ops.channels_last(lg)
        if not fused:
            got = ops.bti_labels(lg)
        else:
            from nextou_b200 import _lib
            import ctypes
            xx, (sb, sc, sv) = ops._prep_logits(lg)
            y = ops._prep_target(target.to(DEV))
            B, V = shape[0], xx[0, 0].numel()
            nblk = ctypes.c_int(0)
            ops.check(_lib.lib().nextou_dsloss_plan(ops.ll(V), B, ctypes.byref(nblk)), "plan")
            got = torch.empty((B, *shape[1:]), device=DEV, dtype=torch.uint8)
            ce = torch.empty((B, V), device=DEV, dtype=torch.float64)
            partial = torch.empty((B, nblk.value, 3 * nc + 1), device=DEV, dtype=torch.float64)
            sums = torch.empty((B, 3 * nc + 1), device=DEV, dtype=torch.float64)
            ops.check(_lib.lib().nextou_dsloss_stats(ops.ptr(xx), ops.dtype_code(xx), ops.ll(sb), ops.ll(sc), ops.ll(sv), B, nc,
                                                     ops.ll(V), ops.ptr(y), ops._TGT_CODE[y.dtype], ops.ptr(got), ops.ptr(ce),
                                                     ops.ptr(partial), ops.ptr(sums), ops.cstream()), "stats")
        assert np.array_equal(got.cpu().numpy(), want.numpy().astype(np.uint8)), layout
    # and the loss built on those labels equals the oracle's (which calls torch.softmax + argmax like the reference)
    from nextou_b200.losses import BTI_Loss
    _, exc = bti_interactions("synapse", nc)
    mod = BTI_Loss(dim=3, connectivity=26, inclusion=[], exclusion=exc, min_thick=1)
    val = mod(x.to(DEV), target.to(DEV))
    ref = TO.bti_loss(x, target, [], exc, 3, 26, 1)
    assert abs(val.item() - ref.item()) <= 1e-9 * max(1.0, abs(ref.item()))


def test_ti_loss_all_pairs_of_13_organs_78_interactions():
    """nnUNetTrainer_NexToU_TI builds every C(13, 2) = 78 pairwise exclusion of the Synapse organs (generate_combinations);
    the reference has no limit on the list length.  Map bit-exact vs the C oracle, loss vs the torch oracle."""
    from itertools import combinations
    from nextou_b200 import ops
    from nextou_b200.losses import TI_Loss
    nc = 14
    exc = [[torch.tensor(a), torch.tensor(b)] for a, b in combinations(range(1, nc), 2)]
    assert len(exc) == 78
    logits, target = bti_case((2, 12, 40, 36), nc, 9)
    mod = TI_Loss(dim=3, connectivity=26, inclusion=[], exclusion=exc, min_thick=1)
    val = mod(logits.to(DEV), target.to(DEV))
    ref = TO.bti_loss(logits, target, [], exc, 3, 26, 1)
    assert abs(val.item() - ref.item()) <= 1e-9 * max(1.0, abs(ref.item()))
    labels = ops.bti_labels(logits.to(DEV))
    ma, mc, flags = mod.interaction_table()
    crit = ops.bti_critical_map(labels, ma, mc, flags, 26, 1)
    want = c_oracle.bti_critical(labels.cpu().numpy(), ma, mc, flags, 26, 1)
    assert np.array_equal(crit.cpu().numpy(), want)
    # more than one launch chunk (> 128 distinct interactions): ordered pairs of 20 labels = 190, OR-accumulated
    ma2 = [1 << a for a, b in combinations(range(20), 2)]
    mc2 = [1 << b for a, b in combinations(range(20), 2)]
    lab20 = torch.randint(0, 20, (1, 8, 24, 32), dtype=torch.uint8)
    crit2 = ops.bti_critical_map(lab20.to(DEV), ma2, mc2, [0] * len(ma2), 26, 1)
    want2 = c_oracle.bti_critical(lab20.numpy(), ma2, mc2, [0] * len(ma2), 26, 1)
    assert np.array_equal(crit2.cpu().numpy(), want2)
    # editing an entry of the public interaction_list in place takes effect (the reference re-reads the list every call)
    mod.interaction_list[0][2] = torch.tensor(13)
    assert mod.interaction_table()[1][0] == 1 << 13


# ------------------------------------------------------------------------------------------------
# fused CE + soft Dice + BTI (csrc/dsloss.cu) vs the oracle's three-term composition
# ------------------------------------------------------------------------------------------------
def _oracle_compound(logits, target, exc, dim, conn, w_ti, batch_dice, do_bg, smooth=1e-5):
    import torch.nn.functional as F
    lg = logits.double().requires_grad_(True)
    p = torch.softmax(lg, 1)
    axes = tuple(range(2, p.ndim))
    onehot = torch.zeros_like(p).scatter_(1, target.long(), 1)
    first = 0 if do_bg else 1
    inter, pred, gt = (p * onehot).sum(axes)[:, first:], p.sum(axes)[:, first:], onehot.sum(axes)[:, first:]
    if batch_dice:
        inter, pred, gt = inter.sum(0), pred.sum(0), gt.sum(0)
    dc = -((2 * inter + smooth) / torch.clip(gt + pred + smooth, 1e-8)).mean()
    ce = F.cross_entropy(lg, target[:, 0].long())
    ti = TO.bti_loss(lg, target, [], exc, dim, conn, 1) if w_ti else 0.0
    total = ce + dc + w_ti * ti
    total.backward()
    return total.item(), lg.grad


@pytest.mark.parametrize("shape,nc,kind,batch_dice,do_bg,w_ti,dtype", [
    ((2, 10, 20, 24), 14, "synapse", True, False, 1e-6, torch.float32),
    ((2, 10, 20, 24), 14, "synapse", False, True, 1e-2, torch.float32),
    ((3, 33, 47), 5, "pairs", True, False, 1e-4, torch.float32),
    ((1, 9, 31, 30), 3, "pairs", False, False, 0.0, torch.float32),
    ((2, 12, 24, 20), 14, "synapse", True, False, 1e-2, torch.bfloat16),
])
@pytest.mark.parametrize("layout", ["ncdhw", "channels_last"])
def test_fused_seg_loss_matches_oracle_and_unfused(shape, nc, kind, batch_dice, do_bg, w_ti, dtype, layout):
    from nextou_b200 import ops
    from nextou_b200.losses import DC_and_CE_and_BTI_Loss
    logits, target = bti_case(shape, nc, 5)
    logits = logits.to(dtype)
    _, exc = bti_interactions(kind, nc)
    dim = len(shape) - 1
    conn = 26 if dim == 3 else 8
    mod = DC_and_CE_and_BTI_Loss({"batch_dice": batch_dice, "smooth": 1e-5, "do_bg": do_bg, "ddp": False}, {},
                                 {"dim": dim, "connectivity": conn, "inclusion": [], "exclusion": exc, "min_thick": 1},
                                 weight_ce=1, weight_dice=1, weight_ti=w_ti)
    lg = logits.to(DEV)
    if layout == "channels_last":
        lg = ops.channels_last(lg)
    lg.requires_grad_(True)
    assert mod._fusable(lg, target.to(DEV))
    n0 = ops._lib.lib().nextou_launch_count()
    val = mod(lg, target.to(DEV))
    (val * 2.0).backward()
    assert ops._lib.lib().nextou_launch_count() - n0 in (5, 8)       # stats + reduce + finish + scale + bwd (+ critical map + 2-stage sum)
    ref, ref_grad = _oracle_compound(logits.float(), target, exc, dim, conn, w_ti, batch_dice, do_bg)
    assert abs(val.item() - ref) <= 1e-5 * max(1.0, abs(ref))          # fp32 softmax / log, fp64 accumulation
    g = lg.grad.double().cpu() / 2.0
    scale = ref_grad.abs().max().item()
    tol = 1e-5 if dtype == torch.float32 else 8e-3                      # bf16 gradient storage: 2^-8 relative
    assert (g - ref_grad).abs().max().item() <= tol * scale
    if dtype != torch.float32:
        return
    # the unfused composition of the product (reference structure) agrees too
    mod.fused = False
    lg2 = lg.detach().clone().requires_grad_(True)
    val2 = mod(lg2, target.to(DEV))
    val2.backward()
    assert abs(val2.item() - val.item()) <= 1e-5 * max(1.0, abs(ref))
    assert (lg2.grad.double().cpu() - g).abs().max().item() <= max(tol, 2e-5) * scale


def test_fused_seg_loss_full_size_properties():
    """BASELINE patch (14 x 64 x 224 x 192): value vs the oracle, and gradient properties that hold at any size —
    every voxel's gradient sums to zero over classes (softmax Jacobian), and the loss is invariant to a per-voxel
    shift of all logits."""
    from nextou_b200.losses import DC_and_CE_and_BTI_Loss
    from oracle.ref_shims import SYNAPSE_EXCLUSION, make_tensors
    exc = make_tensors(SYNAPSE_EXCLUSION)
    logits, target = bti_case((1, 64, 224, 192), 14, 21)
    mod = DC_and_CE_and_BTI_Loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {},
                                 {"dim": 3, "connectivity": 26, "inclusion": [], "exclusion": exc, "min_thick": 1},
                                 weight_ce=1, weight_dice=1, weight_ti=1e-6)
    lg = logits.to(DEV).requires_grad_(True)
    val = mod(lg, target.to(DEV))
    val.backward()
    ref = TO.training_loss([logits, logits], [target, target], exc)     # weights (1, 0): one active scale
    assert abs(val.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert lg.grad.sum(1).abs().max().item() <= 1e-9
    shifted = mod(lg.detach() + torch.randn(1, 1, 64, 224, 192, device=DEV), target.to(DEV))
    assert abs(shifted.item() - val.item()) <= 1e-5 * abs(val.item())
