"""GPU parity of the tcgen05 GEMM engine (plain GEMM and implicit-GEMM convolution) against a CPU fp32 oracle
evaluated on the same bf16-rounded operands (products of bf16 values are exact in fp32, so only the accumulation
order and the final bf16 rounding differ)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("M,N,K,ld_extra", [(128, 16, 64, 0), (300, 132, 132, 4), (1000, 33, 66, 6), (86016 // 8, 264, 132, 4),
                                            (257, 528, 264, 0), (129, 1296, 324, 4), (64, 14, 33, 7), (5000, 324, 1296, 0),
                                            (168, 648, 648, 0)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32], ids=["obf16", "of32"])
def test_gemm_bf16_tn(M, N, K, ld_extra, out_dtype):
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    lda = ops.pad8(K) + (8 if ld_extra else 0)
    a_buf = torch.randn(M, lda, generator=g).bfloat16()
    b_buf = (torch.randn(N, ops.pad8(K), generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g)
    a_buf[:, K:] = float("nan")                             # strided views: lanes beyond K must never be read
    b_buf[:, K:] = float("nan")
    a, b = a_buf[:, :K], b_buf[:, :K]
    want = a.float() @ b.float().t() + bias
    out = ops.gemm_bf16_tn(a_buf.to(DEV)[:, :K], b_buf.to(DEV)[:, :K], bias.to(DEV), out_dtype=out_dtype)
    assert out.shape == (M, ops.pad8(N)) and out.dtype == out_dtype
    got = out.float().cpu()
    assert torch.count_nonzero(got[:, N:]) == 0             # channel padding is written as zeros
    tol = 1e-2 if out_dtype == torch.bfloat16 else 2e-5
    assert torch.allclose(got[:, :N], want, rtol=tol, atol=tol * want.abs().max().item()), _rel(got[:, :N], want)
    if out_dtype == torch.float32:
        assert _rel(got[:, :N], want) < 1e-6


@pytest.mark.parametrize("B,spatial,cin,cout,ks", [
    (1, (4, 16, 24), 33, 33, (1, 3, 3)), (1, (8, 12, 16), 66, 66, (3, 3, 3)), (2, (4, 7, 6), 324, 324, (3, 3, 3)),
    (1, (6, 10, 20), 132, 66, (3, 3, 3)), (1, (1, 20, 36), 16, 40, (1, 3, 3)), (1, (16, 28, 24), 264, 132, (3, 3, 3)),
    (1, (5, 9, 11), 8, 264, (3, 3, 3)), (2, (12, 20), 32, 48, (3, 3)), (1, (3, 18, 10), 64, 32, (1, 3, 1)),
    (1, (3, 18, 10), 64, 32, (1, 1, 3)), (1, (5, 18, 10), 64, 32, (3, 1, 1))])
@pytest.mark.parametrize("halo", [False, True], ids=["pertap", "halo"])
def test_conv_ndhwc_implicit_gemm(B, spatial, cin, cout, ks, halo):
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(cin * 7 + cout)
    dim = len(spatial)
    x = torch.randn(B, cin, *spatial, generator=g).bfloat16()
    w = (torch.randn(cout, cin, *ks, generator=g) / (cin * 9) ** 0.5).bfloat16()
    bias = torch.randn(cout, generator=g)
    conv = F.conv3d if dim == 3 else F.conv2d
    want = conv(x.float(), w.float(), bias, padding=[k // 2 for k in ks])
    ldx = ops.pad8(cin)
    tok = torch.zeros(B * x[0, 0].numel(), ldx, dtype=torch.bfloat16)
    tok[:, :cin] = x.permute(0, *range(2, 2 + dim), 1).reshape(-1, cin)
    if ldx > cin:
        tok[:, cin:] = float("nan")                         # padding lanes must never be read
    out = ops.conv_ndhwc_bf16(tok.to(DEV), B, spatial, cin, ops.pack_conv_weight(w).to(DEV), cout, ks, bias.to(DEV),
                              out_dtype=torch.float32, halo=halo)
    got = out.cpu()[:, :cout].reshape(B, *spatial, cout).permute(0, dim + 1, *range(1, dim + 1))
    assert torch.count_nonzero(out.cpu()[:, cout:]) == 0
    assert _rel(got, want) < 1e-5, _rel(got, want)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-4 * want.abs().max().item())
    # data gradient == the same kernel on dY with the flipped / transposed pack
    dy = torch.randn(B, cout, *spatial, generator=g).bfloat16()
    xg = x.float().requires_grad_(True)
    conv(xg, w.float(), None, padding=[k // 2 for k in ks]).backward(dy.float())
    dtok = torch.zeros(B * x[0, 0].numel(), ops.pad8(cout), dtype=torch.bfloat16)
    dtok[:, :cout] = dy.permute(0, *range(2, 2 + dim), 1).reshape(-1, cout)
    dx = ops.conv_ndhwc_bf16(dtok.to(DEV), B, spatial, cout, ops.pack_conv_weight(w, transpose_flip=True).to(DEV), cin, ks,
                             None, out_dtype=torch.float32, halo=halo)
    gotdx = dx.cpu()[:, :cin].reshape(B, *spatial, cin).permute(0, dim + 1, *range(1, dim + 1))
    assert _rel(gotdx, xg.grad) < 1e-5, _rel(gotdx, xg.grad)


@pytest.mark.parametrize("B,spatial,cin,cout,ks", [
    (1, (4, 16, 24), 33, 33, (1, 3, 3)), (1, (8, 12, 16), 66, 66, (3, 3, 3)), (2, (4, 7, 6), 324, 324, (3, 3, 3)),
    (1, (6, 10, 20), 132, 66, (3, 3, 3)), (1, (16, 28, 24), 264, 132, (3, 3, 3)), (1, (5, 9, 11), 8, 264, (3, 3, 3)),
    (2, (12, 20), 32, 48, (3, 3)), (1, (32, 56, 48), 132, 132, (1, 1, 1)), (1, (2, 7, 12), 648, 324, (1, 1, 1)),
    (1, (8, 64, 96), 1, 33, (1, 3, 3))])
@pytest.mark.parametrize("halo", [False, True], ids=["pertap", "halo"])
def test_conv_wgrad_mn_major(B, spatial, cin, cout, ks, halo):
    """Weight gradient through the MN-major tcgen05 path vs autograd of F.conv on the same bf16-valued operands."""
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(cin * 3 + cout)
    dim = len(spatial)
    x = torch.randn(B, cin, *spatial, generator=g).bfloat16()
    dy = torch.randn(B, cout, *spatial, generator=g).bfloat16()
    w = torch.zeros(cout, cin, *ks, requires_grad=True)
    conv = F.conv3d if dim == 3 else F.conv2d
    conv(x.float(), w, None, padding=[k // 2 for k in ks]).backward(dy.float())
    V = B * x[0, 0].numel()
    xt = torch.full((V, ops.pad8(cin)), float("nan"), dtype=torch.bfloat16)
    xt[:, :cin] = x.permute(0, *range(2, 2 + dim), 1).reshape(-1, cin)
    dt = torch.full((V, ops.pad8(cout)), float("nan"), dtype=torch.bfloat16)
    dt[:, :cout] = dy.permute(0, *range(2, 2 + dim), 1).reshape(-1, cout)
    dw = ops.conv_wgrad_bf16(dt.to(DEV)[:, :cout], xt.to(DEV)[:, :cin], B, spatial, cin, cout, ks, halo=halo)
    assert dw.shape == w.shape and dw.dtype == torch.float32
    assert _rel(dw.cpu(), w.grad) < 1e-5, _rel(dw.cpu(), w.grad)


def _tok_of(x, pad_nan=True):
    from nextou_b200 import ops
    B, C = x.shape[:2]
    dim = x.dim() - 2
    t = torch.full((B * x[0, 0].numel(), ops.pad8(C)), float("nan") if pad_nan else 0.0, dtype=torch.bfloat16)
    t[:, :C] = x.permute(0, *range(2, 2 + dim), 1).reshape(-1, C)
    return t


def _vol_of(tok, B, C, spatial):
    dim = len(spatial)
    return tok.float().cpu()[:, :C].reshape(B, *spatial, C).permute(0, dim + 1, *range(1, dim + 1))


@pytest.mark.parametrize("B,spatial,cin,cout,ks,stride", [
    (1, (8, 16, 24), 33, 66, (3, 3, 3), (1, 2, 2)),      # enc s1 conv0
    (1, (8, 12, 16), 66, 132, (3, 3, 3), (2, 2, 2)),     # enc s2
    (2, (4, 14, 12), 264, 324, (3, 3, 3), (2, 2, 2)),    # enc s4 (batch 2)
    (1, (8, 14, 12), 324, 324, (3, 3, 3), (2, 2, 2)),    # enc s5
    (1, (6, 10, 14), 16, 40, (3, 3, 3), (2, 2, 2)),      # even, small channels
    (1, (7, 9, 11), 24, 24, (3, 3, 3), (2, 2, 2)),       # odd extents
    (2, (16, 20), 33, 66, (3, 3), (2, 2)),               # 2-D
    (1, (8, 12, 12), 32, 32, (1, 1, 1), (2, 2, 2)),      # k < stride: parity classes without taps
])
def test_conv_strided_fwd_bwd(B, spatial, cin, cout, ks, stride):
    """Strided convolution: forward, data gradient (parity classes) and weight gradient (strided X box) through the
    autograd wrapper vs F.conv on the same bf16-valued operands."""
    from nextou_b200 import native
    g = torch.Generator().manual_seed(cin + 3 * cout)
    dim = len(spatial)
    pad = tuple((k - 1) // 2 for k in ks)
    x = torch.randn(B, cin, *spatial, generator=g).bfloat16()
    w = (torch.randn(cout, cin, *ks, generator=g) / (cin * 9) ** 0.5).bfloat16().float()
    bias = torch.randn(cout, generator=g)
    conv = F.conv3d if dim == 3 else F.conv2d
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    want = conv(xr, wr, br, stride=stride, padding=pad)
    dy = torch.randn(want.shape, generator=g).bfloat16()
    want.backward(dy.float())

    xt = _tok_of(x).to(DEV)[:, :cin].requires_grad_(True)
    wd, bd = w.to(DEV).requires_grad_(True), bias.to(DEV).requires_grad_(True)
    y, osp = native.conv_strided_tokens(xt, wd, bd, B, spatial, stride, pad)
    assert tuple(osp) == tuple(want.shape[2:])
    got = _vol_of(y.detach(), B, cout, osp)
    assert _rel(got, want.detach()) < 6e-3, _rel(got, want.detach())                 # bf16 output rounding
    y.backward(_tok_of(dy, pad_nan=False).to(DEV)[:, :cout])
    gx = _vol_of(xt.grad, B, cin, spatial)
    assert _rel(gx, xr.grad) < 6e-3, _rel(gx, xr.grad)
    assert _rel(wd.grad.cpu(), wr.grad) < 1e-4, _rel(wd.grad.cpu(), wr.grad)
    assert _rel(bd.grad.cpu(), br.grad) < 1e-4


@pytest.mark.parametrize("B,spatial,cin,cout,ks", [
    (1, (4, 7, 6), 324, 324, (2, 2, 2)), (1, (4, 14, 12), 324, 264, (2, 2, 2)), (2, (4, 6, 8), 264, 132, (2, 2, 2)),
    (1, (8, 12, 16), 132, 66, (2, 2, 2)), (1, (8, 12, 16), 66, 33, (1, 2, 2)), (2, (10, 12), 66, 33, (2, 2))])
def test_conv_transpose_fwd_bwd(B, spatial, cin, cout, ks):
    """kernel == stride transposed convolution (decoder up-sampling) vs F.conv_transpose."""
    from nextou_b200 import native
    g = torch.Generator().manual_seed(cin + 5 * cout)
    dim = len(spatial)
    x = torch.randn(B, cin, *spatial, generator=g).bfloat16()
    w = (torch.randn(cin, cout, *ks, generator=g) / cin ** 0.5).bfloat16().float()
    bias = torch.randn(cout, generator=g)
    convt = F.conv_transpose3d if dim == 3 else F.conv_transpose2d
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    want = convt(xr, wr, br, stride=ks)
    dy = torch.randn(want.shape, generator=g).bfloat16()
    want.backward(dy.float())

    xt = _tok_of(x).to(DEV)[:, :cin].requires_grad_(True)
    wd, bd = w.to(DEV).requires_grad_(True), bias.to(DEV).requires_grad_(True)
    y, osp = native.conv_transpose_tokens(xt, wd, bd, B, spatial)
    assert tuple(osp) == tuple(want.shape[2:])
    got = _vol_of(y.detach(), B, cout, osp)
    assert _rel(got, want.detach()) < 6e-3, _rel(got, want.detach())
    y.backward(_tok_of(dy, pad_nan=False).to(DEV)[:, :cout])
    gx = _vol_of(xt.grad, B, cin, spatial)
    assert _rel(gx, xr.grad) < 6e-3, _rel(gx, xr.grad)
    assert _rel(wd.grad.cpu(), wr.grad) < 1e-4, _rel(wd.grad.cpu(), wr.grad)
    assert _rel(bd.grad.cpu(), br.grad) < 1e-4


@pytest.mark.parametrize("R,Cc,ks,groups,conv,flip", [
    (66, 33, (3, 3, 3), 1, True, True), (33, 66, (1, 3, 3), 1, True, False), (324, 324, (2, 2, 2), 1, True, False),
    (132, 132, (1, 1, 1), 1, False, False), (264, 264, (1, 1, 1), 6, False, False), (14, 33, (1, 1, 1), 1, False, False),
    (528, 132, (1, 1), 1, False, False), (24, 24, (1, 1), 4, False, False)])
def test_pack_weight_pair(R, Cc, ks, groups, conv, flip):
    """The one-launch operand packs equal the ATen reshape / permute / pad / block_diag chains they replace."""
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(R + Cc)
    w = torch.randn(R, Cc // groups, *ks, generator=g)
    a, b = ops.pack_weight_pair(w.to(DEV), conv=conv, flip_b=flip, groups=groups)
    taps = 1
    for k in ks:
        taps *= k
    pad = (lambda c: (c + 63) // 64 * 64) if conv else ops.pad8
    dense = w
    if groups > 1:
        blocks = w.reshape(groups, R // groups, Cc // groups)
        dense = torch.block_diag(*blocks.unbind(0)).reshape(R, Cc, *ks)
    d3 = dense.reshape(R, Cc, taps)
    want_a = torch.zeros(R, taps, pad(Cc))
    want_a[:, :, :Cc] = d3.permute(0, 2, 1)
    want_b = torch.zeros(Cc, taps, pad(R))
    src = d3.flip(2) if flip else d3
    want_b[:, :, :R] = src.permute(1, 2, 0)
    assert torch.equal(a.float().cpu(), want_a.reshape(R, -1).bfloat16().float())
    assert torch.equal(b.float().cpu(), want_b.reshape(Cc, -1).bfloat16().float())
    if conv and groups == 1:        # same as the ATen reference packer
        assert torch.equal(a.cpu(), ops.pack_conv_weight(w))
        assert torch.equal(b.cpu(), ops.pack_conv_weight(w, transpose_flip=True) if flip else ops.pack_conv_weight(w, transpose=True))


def test_fork_tokens_sums_gradients_in_padded_layout():
    """ops.fork_tokens: two aliases whose gradients are summed by the package (padded pitch kept), equal to autograd's sum."""
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(3)
    rows, C = 500, 33
    buf = torch.randn(rows, ops.pad8(C), generator=g).bfloat16().to(DEV)
    x = buf[:, :C].requires_grad_(True)
    a, b = ops.fork_tokens(x)
    ga = torch.randn(rows, ops.pad8(C), generator=g).bfloat16().to(DEV)[:, :C]     # padded-pitch gradient (GEMM dgrad output)
    gb = torch.randn(rows, C, generator=g).bfloat16().to(DEV)                      # dense gradient (autograd add output)
    (a.float() * ga.float()).sum().backward(retain_graph=True)
    assert torch.equal(x.grad, ga)                                                # single consumer: passed through
    x.grad = None
    seen = []
    h = x.register_hook(lambda t: seen.append(t.stride(0)))                       # layout of the summed gradient itself
    ((a.float() * ga.float()).sum() + (b.float() * gb.float()).sum()).backward()
    h.remove()
    assert torch.equal(x.grad.float(), (ga.float() + gb.float()).bfloat16().float())
    assert seen and seen[0] % 8 == 0                                              # TMA-ready: no re-padding copy downstream
    x.grad = None
    a2, b2 = ops.fork_tokens(x)
    ((a2.float() * ga.float()).sum() + (b2.float() * ga.float()).sum()).backward()   # both padded: vectorised full-row add
    assert torch.equal(x.grad.float(), (2 * ga.float()).bfloat16().float())


@pytest.mark.parametrize("B,spatial,cin,cskip,ks", [(1, (4, 7, 6), 324, 324, (2, 2, 2)), (1, (8, 12, 16), 132, 66, (2, 2, 2)),
                                                    (2, (8, 12, 16), 66, 33, (1, 2, 2)), (1, (10, 12), 66, 33, (2, 2))])
def test_up_cat_matches_transpose_conv_plus_cat(B, spatial, cin, cskip, ks):
    """dense.up_cat == torch.cat((conv_transpose(low), skip), 1) up to the zero channel gap, forward and backward."""
    from nextou_b200 import dense, ops
    g = torch.Generator().manual_seed(cin + cskip)
    dim = len(spatial)
    osp = tuple(n * k for n, k in zip(spatial, ks))
    low = torch.randn(B, cin, *spatial, generator=g).bfloat16()
    skip = torch.randn(B, cskip, *osp, generator=g).bfloat16()
    tc = (torch.nn.ConvTranspose3d if dim == 3 else torch.nn.ConvTranspose2d)(cin, cskip, ks, ks, bias=True)
    with torch.no_grad():
        tc.weight.copy_(tc.weight.bfloat16().float())
    lr, sr = low.float().requires_grad_(True), skip.float().requires_grad_(True)
    want = torch.cat((tc(lr), sr), 1)
    dy = torch.randn(want.shape, generator=g).bfloat16()
    want.backward(dy.float())
    wg, bg = tc.weight.grad.clone(), tc.bias.grad.clone()
    tc.zero_grad()
    tcd = tc.to(DEV)
    ld = ops.channels_last(low.to(DEV)).requires_grad_(True)
    sd = ops.channels_last(skip.to(DEV)).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got, gap = dense.up_cat(ld, tcd, sd)
    ca, pa = gap
    assert ca == cskip and pa == ops.pad8(cskip) and got.shape[1] == pa + cskip
    gf = got.detach().float().cpu()
    assert float(gf[:, ca:pa].abs().max()) == 0.0 if pa > ca else True
    plain = torch.cat([gf[:, :ca], gf[:, pa:]], 1)
    assert _rel(plain, want.detach()) < 6e-3
    dyg = torch.zeros(B, pa + cskip, *osp)
    dyg[:, :ca], dyg[:, pa:] = dy[:, :ca].float(), dy[:, ca:].float()
    got.backward(dyg.to(DEV).to(got.dtype))
    assert _rel(ld.grad.float().cpu(), lr.grad) < 6e-3
    assert torch.equal(sd.grad.float().cpu(), dy[:, ca:].float())
    assert _rel(tcd.weight.grad.cpu(), wg) < 1e-4 and _rel(tcd.bias.grad.cpu(), bg) < 1e-4


@pytest.mark.parametrize("kind", ["conv3", "conv_strided", "linear", "grouped"])
def test_folded_eval_norm_matches_fp32_reference(kind):
    """Inference path: an eval-mode BatchNorm (+ LeakyReLU) folded into the conv / GEMM epilogue equals
    lrelu(bn_eval(layer(x))) in fp32 on the same bf16-valued operands (only the final bf16 rounding differs)."""
    from nextou_b200 import dense, ops
    g = torch.Generator().manual_seed(11)
    B, spatial = 1, (6, 10, 12)
    if kind == "conv3":
        layer = torch.nn.Conv3d(33, 66, 3, 1, 1, bias=True)
    elif kind == "conv_strided":
        layer = torch.nn.Conv3d(33, 66, 3, 2, 1, bias=True)
    elif kind == "linear":
        layer = torch.nn.Conv3d(132, 264, 1, bias=True)
    else:
        layer = torch.nn.Conv3d(264, 264, 1, bias=True, groups=6)
    cin, cout = layer.in_channels, layer.out_channels
    bn = torch.nn.BatchNorm3d(cout)
    with torch.no_grad():
        layer.weight.copy_(layer.weight.bfloat16().float())
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.uniform_(-0.5, 0.5, generator=g)
        bn.running_mean.uniform_(-0.3, 0.3, generator=g); bn.running_var.uniform_(0.5, 2.0, generator=g)
    bn.eval()
    x = torch.randn(B, cin, *spatial, generator=g).bfloat16()
    with torch.no_grad():
        want = F.leaky_relu(bn(layer(x.float())), 0.01)
    layer, bn = layer.to(DEV), bn.to(DEV)
    tok = _tok_of(x).to(DEV)[:, :cin]
    dense.stats.clear()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        if kind.startswith("conv"):
            y, osp = dense.conv_norm_act_tokens(tok, B, spatial, layer, bn, 0.01)
        else:
            y, osp = dense.linear_norm_act_tokens(tok, layer, bn, B, 0.01), spatial
    assert dense.stats["tcgen05.conv_folded_norm"] + dense.stats["tcgen05.linear_folded_norm"] == 1
    assert tuple(osp) == tuple(want.shape[2:])
    got = _vol_of(y, B, cout, osp)
    assert _rel(got, want) < 4e-3, _rel(got, want)
    # training mode (or autograd on) must NOT fold
    dense.stats.clear()
    bn.train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        if kind.startswith("conv"):
            dense.conv_norm_act_tokens(tok, B, spatial, layer, bn, 0.01)
        else:
            dense.linear_norm_act_tokens(tok, layer, bn, B, 0.01)
    assert dense.stats["tcgen05.conv_folded_norm"] + dense.stats["tcgen05.linear_folded_norm"] == 0


# ------------------------------------------------------------------------------------------------------
# full-resolution layers of 3d_fullres_nextou: the shapes bench.py runs — 21 504 bricks over the
# persistent CTAs, 80-byte row pitch at 33 channels, the [up | gap | skip] concat layout — against fp32 cuDNN (TF32 off)
# on the same bf16-valued operands, all on the GPU (the CPU would need minutes)
# ------------------------------------------------------------------------------------------------------
def _gpu_tok(x, pitch=None):
    """(B, C, *sp) bf16 CUDA tensor -> [rows, C] view over a [rows, pitch] buffer whose padding lanes hold NaN."""
    from nextou_b200 import ops
    B, C = x.shape[:2]
    dim = x.dim() - 2
    pitch = ops.pad8(C) if pitch is None else pitch
    t = torch.full((B * x[0, 0].numel(), pitch), float("nan"), dtype=torch.bfloat16, device=x.device)
    t[:, :C] = x.permute(0, *range(2, 2 + dim), 1).reshape(-1, C)
    return t[:, :C]


def _gpu_vol(tok, B, C, spatial):
    dim = len(spatial)
    return tok[:, :C].float().reshape(B, *spatial, C).permute(0, dim + 1, *range(1, dim + 1))


@pytest.mark.parametrize("spatial,cin,cout,ks,gap", [
    ((64, 224, 192), 33, 33, (1, 3, 3), False),     # enc s0 conv1 / dec st4 conv1
    ((64, 224, 192), 66, 33, (1, 3, 3), True),      # dec st4 conv0 on the [up 33 | zero gap 7 | skip 33] concat layout
    ((64, 112, 96), 66, 66, (3, 3, 3), False),      # enc s1 conv1 / dec st3 conv1
], ids=["s0_33to33_1x3x3", "dec4_66to33_gap_layout", "s1_66to66_3x3x3"])
def test_full_resolution_conv_layer_fwd_dgrad_wgrad(spatial, cin, cout, ks, gap):
    from nextou_b200 import native, ops
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device=DEV).manual_seed(cin * 131 + cout)
        x = torch.randn(1, cin, *spatial, generator=g, device=DEV).bfloat16()
        w = (torch.randn(cout, cin, *ks, generator=g, device=DEV) / (cin * 9) ** 0.5).bfloat16().float()
        bias = torch.randn(cout, generator=g, device=DEV)
        dy = torch.randn(1, cout, *spatial, generator=g, device=DEV).bfloat16()
        xr, wr = x.float().requires_grad_(True), w.clone().requires_grad_(True)
        want = F.conv3d(xr, wr, bias, padding=[k // 2 for k in ks])
        want.backward(dy.float())
        if gap:
            # physical input: the decoder's concatenation buffer (native.up_cat_tokens): up-sampled half, zero-filled gap up
            # to the next multiple of 8 channels, skip half; the weight gets matching zero columns (dense.conv_tokens)
            ca, pa = cin // 2, ops.pad8(cin // 2)
            xcat = torch.cat([x[:, :ca], torch.zeros(1, pa - ca, *spatial, device=DEV, dtype=torch.bfloat16), x[:, ca:]], 1)
            xt = _gpu_tok(xcat).requires_grad_(True)
            wd = w.clone().requires_grad_(True)
            wfull = torch.cat([wd[:, :ca], wd.new_zeros(cout, pa - ca, *ks), wd[:, ca:]], 1)
        else:
            xt = _gpu_tok(x).requires_grad_(True)
            wd = w.clone().requires_grad_(True)
            wfull = wd
        bd = bias.clone().requires_grad_(True)
        y = native.conv_tokens(xt, wfull, bd, 1, spatial)
        y.backward(_gpu_tok(dy))
        torch.cuda.synchronize()
        got = _gpu_vol(y.detach(), 1, cout, spatial)
        assert _rel(got, want.detach()) < 4e-3, _rel(got, want.detach())                    # bf16 output rounding (2^-9 rms)
        assert (got - want.detach()).abs().max().item() <= 2 ** -7 * want.abs().max().item()
        gx = _gpu_vol(xt.grad, 1, xt.shape[1], spatial)
        if gap:
            assert torch.count_nonzero(gx[:, ca:pa]) == 0             # gap lanes: data gradient through zero weight columns
            gx = torch.cat([gx[:, :ca], gx[:, pa:]], 1)
        assert _rel(gx, xr.grad) < 4e-3, _rel(gx, xr.grad)
        assert _rel(wd.grad, wr.grad) < 2e-4, _rel(wd.grad, wr.grad)                         # fp32 accumulation over 0.7-2.75 M voxels
        bias_ref = dy.float().sum((0, 2, 3, 4))
        assert _rel(bd.grad, bias_ref) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("B,spatial,cin,cout,ks", [
    (1, (8, 32, 40), 1, 33, (1, 3, 3)),       # the first convolution of 3d_fullres_nextou (one CT modality)
    (2, (24, 20), 1, 8, (3, 3)),               # 2-D, batch 2
    (1, (4, 16, 24), 4, 33, (1, 1, 3)),        # four MR modalities, 3 taps
    (1, (4, 16, 24), 2, 16, (1, 1, 1)),
    (1, (5, 9, 11), 3, 40, (3, 1, 1))])
def test_small_cin_convolution_fwd_dgrad_wgrad(B, spatial, cin, cout, ks):
    """csrc/conv_small.cu (CUDA-core streaming kernels for Cin <= 4) through the autograd wrapper vs F.conv on the same
    bf16-valued operands; the data gradient (only needed when the layer is not the first one) takes the tensor-core path."""
    from nextou_b200 import _lib, native, ops
    g = torch.Generator().manual_seed(cin * 11 + cout)
    dim = len(spatial)
    conv = F.conv3d if dim == 3 else F.conv2d
    x = torch.randn(B, cin, *spatial, generator=g).bfloat16()
    w = (torch.randn(cout, cin, *ks, generator=g) / (cin * 3) ** 0.5).bfloat16().float()
    bias = torch.randn(cout, generator=g)
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    want = conv(xr, wr, br, padding=[k // 2 for k in ks])
    dy = torch.randn(want.shape, generator=g).bfloat16()
    want.backward(dy.float())
    assert ops.conv_small_supported(cin, cout, ks, w.to(DEV))
    xt = x.permute(0, *range(2, 2 + dim), 1).reshape(-1, cin).contiguous().to(DEV).requires_grad_(True)     # unpadded rows
    wd, bd = w.to(DEV).requires_grad_(True), bias.to(DEV).requires_grad_(True)
    n0 = _lib.launch_count()
    y = native.conv_tokens(xt, wd, bd, B, spatial)
    assert _lib.launch_count() - n0 == 1                                   # no pack kernel, no padded copy of the image
    got = _vol_of(y.detach(), B, cout, spatial)
    assert _rel(got, want.detach()) < 4e-3, _rel(got, want.detach())
    y.backward(_tok_of(dy, pad_nan=False).to(DEV)[:, :cout])
    assert _rel(wd.grad.cpu(), wr.grad) < 1e-5, _rel(wd.grad.cpu(), wr.grad)
    assert _rel(bd.grad.cpu(), br.grad) < 1e-4
    gx = _vol_of(xt.grad, B, cin, spatial)
    assert _rel(gx, xr.grad) < 6e-3, _rel(gx, xr.grad)
