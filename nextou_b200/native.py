"""Autograd wrappers around the tcgen05 GEMM engine (csrc/gemm_tcgen05.cu) for bf16 token-major activations.

forward  : hand-written tcgen05 kernels (TMA -> swizzled smem -> tcgen05.mma -> TMEM -> epilogue).
backward : data gradients through the same kernels (a GEMM / convolution with the transposed / flipped weight
           pack); weight gradients through the MN-major tcgen05 kernel (voxel axis = K, split over CTAs).
Token matrices are [rows, C] views with a row pitch that is a multiple of 8 elements (channel padding, don't-care).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops


def _f32(t):
    return None if t is None else t.detach().float()


class _Owner:
    """Carries the nn.Parameter a weight argument is (a view of) through autograd.Function.apply as a NON-tensor argument, so
    that its persistent operand packs (ops.PackEntry) can be found without making it a second differentiable input."""
    __slots__ = ("p",)

    def __init__(self, p):
        self.p = p if isinstance(p, torch.nn.Parameter) else None


def _unwrap(owner):
    return None if owner is None else owner.p


def _fork_wgrad(ctx, tensor):
    """A side-stream launcher for the layer's weight-gradient kernel when its data gradient is computed as well (the two
    kernels are independent: they become parallel branches of the step graph), else None."""
    from ._lib import KernelTimers
    if ops.OVERLAP_WGRAD and ctx.needs_input_grad[0] and ctx.needs_input_grad[1] and not KernelTimers.enabled:
        return ops.side_launch(tensor.device)       # (per-kernel event timing wants every kernel alone on the GPU)
    return None


class _PendingWgrads:
    """Weight-gradient kernels of the running backward pass that the main stream has not waited for yet."""
    keep = []          # tensors / closures the side-stream kernels still read or write
    task = None        # autograd graph task whose end-of-backward callback is queued


def _flush_pending(device):
    torch.cuda.current_stream(device).wait_stream(ops.side_launch(device).side)
    _PendingWgrads.keep.clear()        # only now may the caching allocator hand these blocks to main-stream tensors
    _PendingWgrads.task = None


def _data_parallel() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _complete_wgrad(side, param, finish, keep):
    """Finish a side-stream weight gradient: `finish()` turns the kernel's raw fp32 accumulator into the gradient tensor
    (layout copy, gap / group slicing).

    Default: join (the main stream waits for the side stream), then finish on the main stream.
    Deferred (ops.DEFER_WGRAD_JOIN): when the parameter has no gradient yet and no post-accumulate hooks, autograd's
    AccumulateGrad takes the returned tensor as param.grad without launching a kernel, so nothing on the main stream reads
    it before the backward pass ends.  Then `finish` runs on the SIDE stream as well, the operands (`keep`) stay referenced,
    and one join at the end of the backward pass (engine callback) orders everything before the optimizer.  The main stream
    goes straight on to the previous layer's normalisation backward, which is HBM bound and co-resides with the
    tensor-bound weight-gradient CTAs.  Gradient accumulation (param.grad already set), DistributedDataParallel and foreign
    hooks take the default path; parallel.GradientAllReducer's hook is stream-safe (it copies behind the side stream)."""
    hooks = getattr(param, "_post_accumulate_grad_hooks", None) if param is not None else None
    ours = bool(hooks) and all(getattr(h, "_nextou_stream_safe", False) for h in hooks.values())
    if ours:
        hooks = None       # parallel.GradientAllReducer: bookkeeping on the host, its copies run behind the side stream
    elif _data_parallel():
        # a process group without our reducer on this parameter: torch's DistributedDataParallel hooks the AccumulateGrad NODE
        # (invisible from here) and copies the gradient into its bucket on the main stream right away -> join per layer
        hooks = True
    if not (ops.DEFER_WGRAD_JOIN and param is not None and param.grad is None and not hooks and not torch.is_grad_enabled()):
        side.join()
        return finish()
    with torch.cuda.stream(side.side):
        # the layout copy into the parameter's own (contiguous) layout happens HERE, on the side stream: a strided view would
        # be copied by AccumulateGrad on the main stream, before the kernel has written it
        dw = finish().contiguous()
    if dw.numel() != param.numel() or not param.is_contiguous():     # AccumulateGrad would copy it on the main stream: wait first
        side.join()
        return dw
    device = dw.device
    task = torch._C._current_graph_task_id()
    if _PendingWgrads.task != task:
        if _PendingWgrads.keep:            # a backward pass that raised before its callback ran: settle it now
            _flush_pending(device)
        _PendingWgrads.task = task
        torch.autograd.Variable._execution_engine.queue_callback(lambda: _flush_pending(device))
    _PendingWgrads.keep.append((keep, finish))
    return dw


class _LinearTokens(torch.autograd.Function):
    """y[T, N] = x[T, K] @ w2d[N, K]^T + b  with w2d an fp32 / bf16 master weight."""

    @staticmethod
    def forward(ctx, x, w2d, bias, groups=1, owner=None):
        xb = ops.tma_ready_bf16(x)
        N = w2d.shape[0]
        K = w2d.shape[1] * groups
        # one launch packs the forward operand [N, K] and the data-gradient operand [K, N] (bf16, padded pitches); a grouped
        # layer (TN:85) becomes the block-diagonal dense operand, which keeps the output token-major without a permute.
        # The packs persist on the parameter (`owner`) until it changes (ops.pack_weight_pair).
        wp, wt = ops.pack_weight_pair(w2d, conv=False, groups=groups, want_b=ctx.needs_input_grad[0], owner=_unwrap(owner))
        y = ops.gemm_bf16_tn(xb, wp[:, :K], bias, n=N)[:, :N]
        ctx.save_for_backward(xb, w2d)
        ctx.wt = wt
        ctx.owner = owner
        ctx.groups = groups
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, w2d = ctx.saved_tensors
        g = ctx.groups
        N = w2d.shape[0]
        K = w2d.shape[1] * g
        dyb = ops.tma_ready_bf16(dy)
        dx = dw = db = None
        side = _fork_wgrad(ctx, dyb)
        if ctx.needs_input_grad[1]:
            # dW = dY^T X through the MN-major tcgen05 weight-gradient kernel (a 1x1 "convolution" over T voxels)
            dw = ops.conv_wgrad_bf16(dyb, xb, 1, (xb.shape[0],), K, N, (1,), side=side)
        if ctx.needs_input_grad[0]:
            dx = ops.gemm_bf16_tn(dyb, ctx.wt[:, :N], None, n=K)[:, :K]          # dX = dY W
        if ctx.needs_input_grad[1]:
            raw = dw

            def finish():
                d = (raw() if side is not None else raw).reshape(N, K)
                if g > 1:   # diagonal blocks of the dense gradient -> (Cout, Cin / groups)
                    d = d.view(g, N // g, g, K // g).diagonal(dim1=0, dim2=2).permute(2, 0, 1).reshape(N, K // g)
                return d.to(w2d.dtype)
            dw = _complete_wgrad(side, _unwrap(ctx.owner), finish, (dyb, xb)) if side is not None else finish()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum_tokens(dyb).to(ctx.bias_dtype)
        return dx, dw, db, None, None


def linear_tokens(x_tok: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """1x1 convolution on token rows through the tcgen05 GEMM; weight: (Cout, Cin, 1, ...)."""
    return _LinearTokens.apply(x_tok, weight.reshape(weight.shape[0], -1), bias, 1, _Owner(weight))


def grouped_linear_tokens(x_tok: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], groups: int):
    """Grouped 1x1 convolution: the groups are the diagonal blocks of one [Cout, Cin] operand (TN:85)."""
    return _LinearTokens.apply(x_tok, weight.reshape(weight.shape[0], -1), bias, groups, _Owner(weight))


class _ConvTokens(torch.autograd.Function):
    """Stride-1 'same' convolution on a token-major bf16 volume (implicit GEMM).  gap = (lo, hi): the input rows carry hi - lo
    extra zero channels after logical channel lo (the decoder's [up | gap | skip] concatenation buffer, ED:322); the weight
    packs get matching zero columns, the weight gradient drops them again."""

    @staticmethod
    def forward(ctx, x, weight, bias, batch, spatial, gap=None, owner=None):
        cout, cin = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        gapw = 0 if gap is None else gap[1] - gap[0]
        # the network's first convolution (Cin = image modalities): K = taps x Cin is far below one tensor-core slab
        ctx.small = gap is None and ops.conv_small_supported(cin, cout, ks, weight)
        ctx.owner = owner
        if ctx.small:      # scalar loads: no 16-byte row pitch needed, so the image is not copied into a channel-padded buffer
            xb = x if (x.dtype == torch.bfloat16 and (x.stride(1) == 1 or x.shape[1] == 1)) else x.to(torch.bfloat16).contiguous()
        else:
            xb = ops.tma_ready_bf16(x)
        if ctx.small:
            y = ops.conv_small_fwd(xb, batch, spatial, weight, bias)[:, :cout]
            wt = None
        else:
            wp, wt = ops.pack_weight_pair(weight, conv=True, flip_b=True, want_b=ctx.needs_input_grad[0], gap=gap, owner=_unwrap(owner))
            y = ops.conv_ndhwc_bf16(xb, batch, spatial, cin + gapw, wp, cout, ks, bias)[:, :cout]
        ctx.save_for_backward(xb, weight)
        ctx.wt = wt            # data-gradient operator: [Cin (+ gap), flipped taps * cout_pad]
        ctx.meta = (batch, tuple(spatial), bias is not None, None if bias is None else bias.dtype, gap)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, weight = ctx.saved_tensors
        batch, spatial, has_bias, bdt, gap = ctx.meta
        cout, cin = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        cin_p = cin + (0 if gap is None else gap[1] - gap[0])        # physical input channels
        dyb = ops.tma_ready_bf16(dy)
        dx = dw = db = None
        side = None if ctx.small else _fork_wgrad(ctx, dyb)
        if ctx.needs_input_grad[1]:
            dw = ops.conv_small_wgrad(dyb, xb, batch, spatial, cin, cout, ks) if ctx.small else \
                ops.conv_wgrad_bf16(dyb, xb, batch, spatial, cin_p, cout, ks, side=side)
        if ctx.needs_input_grad[0]:
            wt = ctx.wt
            if wt is None:     # small-Cin layer whose input wants a gradient (never the image itself): tensor-core data gradient
                _, wt = ops.pack_weight_pair(weight, conv=True, flip_b=True, want_b=True, owner=_unwrap(ctx.owner))
            dx = ops.conv_ndhwc_bf16(dyb, batch, spatial, cout, wt, cin_p, ks, None)[:, :cin_p]
        if ctx.needs_input_grad[1]:
            raw = dw

            def finish():
                d = raw() if side is not None else raw
                if gap is not None and gap[1] > gap[0]:
                    d = torch.cat([d[:, :gap[0]], d[:, gap[1]:]], 1)
                return d.to(weight.dtype)
            dw = _complete_wgrad(side, _unwrap(ctx.owner), finish, (dyb, xb)) if side is not None else finish()
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum_tokens(dyb).to(bdt)
        return dx, dw, db, None, None, None, None


def conv_tokens(x_tok, weight, bias, batch: int, spatial: Sequence[int], gap=None):
    return _ConvTokens.apply(x_tok, weight, bias, batch, tuple(spatial), None if gap is None else tuple(gap), _Owner(weight))


def _out_spatial(spatial, ks, stride, padding):
    return tuple((n + 2 * p - k) // s + 1 for n, k, s, p in zip(spatial, ks, stride, padding))


class _ConvStridedTokens(torch.autograd.Function):
    """Strided convolution (the down-sampling conv of an encoder stage) on a token-major bf16 volume: forward = per-tap
    implicit GEMM with a strided TMA gather, data gradient = one stride-1 sub-convolution per output parity class,
    weight gradient = MN-major tcgen05 kernel with a strided X box."""

    @staticmethod
    def forward(ctx, x, weight, bias, batch, spatial, stride, padding, owner=None):
        xb = ops.tma_ready_bf16(x)
        cout, cin = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        wp, wt = ops.pack_weight_pair(weight, conv=True, flip_b=False, want_b=ctx.needs_input_grad[0], owner=_unwrap(owner))
        y, _ = ops.conv_strided_fwd_bf16(xb, batch, spatial, cin, wp, cout, ks, stride, padding, bias)
        ctx.save_for_backward(xb, weight)
        ctx.wt = wt
        ctx.owner = owner
        ctx.meta = (batch, tuple(spatial), tuple(stride), tuple(padding), bias is not None, None if bias is None else bias.dtype)
        return y[:, :cout]

    @staticmethod
    def backward(ctx, dy):
        xb, weight = ctx.saved_tensors
        batch, spatial, stride, padding, has_bias, bdt = ctx.meta
        cout, cin = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        osp = _out_spatial(spatial, ks, stride, padding)
        dyb = ops.tma_ready_bf16(dy)
        dx = dw = db = None
        side = _fork_wgrad(ctx, dyb)
        if ctx.needs_input_grad[1]:
            dw = ops.conv_strided_wgrad_bf16(dyb, xb, batch, osp, spatial, cin, cout, ks, stride, padding, side=side)
        if ctx.needs_input_grad[0]:
            dx = ops.conv_strided_dgrad_bf16(dyb, batch, osp, cout, ctx.wt, cin, ks, stride, padding, spatial)[:, :cin]
        if ctx.needs_input_grad[1]:
            raw = dw

            def finish():
                return (raw() if side is not None else raw).permute(0, 2, 1).reshape(cout, cin, *ks).to(weight.dtype)
            dw = _complete_wgrad(side, _unwrap(ctx.owner), finish, (dyb, xb)) if side is not None else finish()
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum_tokens(dyb).to(bdt)
        return dx, dw, db, None, None, None, None, None


def conv_strided_tokens(x_tok, weight, bias, batch: int, spatial: Sequence[int], stride, padding):
    """-> (token rows [rows_out, Cout], output spatial shape)."""
    ks = tuple(weight.shape[2:])
    y = _ConvStridedTokens.apply(x_tok, weight, bias, batch, tuple(spatial), tuple(stride), tuple(padding), _Owner(weight))
    return y, _out_spatial(spatial, ks, stride, padding)


class _ConvTransposeTokens(torch.autograd.Function):
    """kernel == stride transposed convolution (decoder up-sampling, ED:273-276): every input voxel owns a disjoint
    kd x kh x kw block of output voxels, so the forward is the strided data-gradient kernel (one tap per parity class), the
    data gradient a strided convolution of dY and the weight gradient the strided weight-gradient kernel with the roles
    of the operands exchanged.  weight: (Cin, Cout, *k) as nn.ConvTranspose stores it."""

    @staticmethod
    def forward(ctx, x, weight, bias, batch, spatial, owner=None):
        xb = ops.tma_ready_bf16(x)
        cin, cout = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        osp = tuple(n * k for n, k in zip(spatial, ks))
        zero = (0,) * len(ks)
        # A = [Cin][tap][Cout pad] (operand of the data gradient), Bt = [Cout][tap][Cin pad] (operand of the forward)
        wa, wb = ops.pack_weight_pair(weight, conv=True, flip_b=False, owner=_unwrap(owner))
        y = ops.convtranspose_fwd_bf16(xb, batch, spatial, cin, wb, cout, ks, bias)
        ctx.save_for_backward(xb, weight)
        ctx.wa = wa
        ctx.meta = (batch, tuple(spatial), osp, bias is not None, None if bias is None else bias.dtype)
        return y[:, :cout]

    @staticmethod
    def backward(ctx, dy):
        xb, weight = ctx.saved_tensors
        batch, spatial, osp, has_bias, bdt = ctx.meta
        cin, cout = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        zero = (0,) * len(ks)
        dyb = ops.tma_ready_bf16(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx, _ = ops.conv_strided_fwd_bf16(dyb, batch, osp, cout, ctx.wa, cin, ks, ks, zero, None)   # (Cin, Cout, k) as a conv
            dx = dx[:, :cin]
        if ctx.needs_input_grad[1]:
            dw = ops.conv_strided_wgrad_bf16(xb, dyb, batch, spatial, osp, cout, cin, ks, ks, zero)   # [Cin][tap][Cout]
            dw = dw.permute(0, 2, 1).reshape(cin, cout, *ks).to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum_tokens(dyb).to(bdt)
        return dx, dw, db, None, None, None


def conv_transpose_tokens(x_tok, weight, bias, batch: int, spatial: Sequence[int]):
    """-> (token rows [rows_out, Cout], output spatial shape) of a kernel == stride transposed convolution."""
    ks = tuple(weight.shape[2:])
    y = _ConvTransposeTokens.apply(x_tok, weight, bias, batch, tuple(spatial), _Owner(weight))
    return y, tuple(n * k for n, k in zip(spatial, ks))


class _UpCatTokens(torch.autograd.Function):
    """torch.cat((ConvTranspose(low), skip), 1) of a decoder stage (ED:321-322) in ONE buffer: the transposed convolution
    writes its Cout channels (+ zero padding up to a multiple of 8) straight into the first columns of the concatenation
    rows, the skip channels are copied behind them at the 16-byte aligned column pad8(Cout).  The logical result has
    pad8(Cout) + Cskip channels (a zero gap after the up-sampled half; the consuming convolution's weight gets matching zero
    columns, dense.conv_tokens), so both halves of the incoming gradient are TMA-ready views: no copy in the backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, skip, batch, spatial, owner=None):
        xb = ops.tma_ready_bf16(x)
        cin, cout = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        osp = tuple(n * k for n, k in zip(spatial, ks))
        zero = (0,) * len(ks)
        cb = skip.shape[1]
        pa = ops.pad8(cout)
        rows = skip.shape[0]
        buf = torch.empty((rows, pa + ops.pad8(cb)), device=x.device, dtype=torch.bfloat16)
        wa, wb = ops.pack_weight_pair(weight, conv=True, flip_b=False, owner=_unwrap(owner))
        ops.convtranspose_fwd_bf16(xb, batch, spatial, cin, wb, cout, ks, bias, out=buf, store_cols=pa)
        if not ops.rows_copy_add(skip, None, buf[:, pa:pa + cb]):      # skip half behind the up-sampled half (ED:322)
            buf[:, pa:pa + cb].copy_(skip)
        ctx.save_for_backward(xb, weight)
        ctx.wa = wa
        ctx.owner = owner
        ctx.meta = (batch, tuple(spatial), osp, bias is not None, None if bias is None else bias.dtype, pa, cb, skip.dtype)
        return buf[:, :pa + cb]

    @staticmethod
    def backward(ctx, g):
        xb, weight = ctx.saved_tensors
        batch, spatial, osp, has_bias, bdt, pa, cb, sdt = ctx.meta
        cin, cout = weight.shape[:2]
        ks = tuple(weight.shape[2:])
        zero = (0,) * len(ks)
        dyb = ops.tma_ready_bf16(g[:, :cout])
        dx = dw = db = dskip = None
        side = _fork_wgrad(ctx, dyb)
        if ctx.needs_input_grad[1]:
            dw = ops.conv_strided_wgrad_bf16(xb, dyb, batch, spatial, osp, cout, cin, ks, ks, zero, side=side)
        if ctx.needs_input_grad[0]:
            dx, _ = ops.conv_strided_fwd_bf16(dyb, batch, osp, cout, ctx.wa, cin, ks, ks, zero, None)
            dx = dx[:, :cin]
        if ctx.needs_input_grad[1]:
            raw = dw

            def finish():
                return (raw() if side is not None else raw).permute(0, 2, 1).reshape(cin, cout, *ks).to(weight.dtype)
            dw = _complete_wgrad(side, _unwrap(ctx.owner), finish, (dyb, xb)) if side is not None else finish()
        if has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum_tokens(dyb).to(bdt)
        if ctx.needs_input_grad[3]:
            dskip = g[:, pa:pa + cb]
            if dskip.dtype != sdt:
                dskip = dskip.to(sdt)
        return dx, dw, db, dskip, None, None, None


def up_cat_tokens(x_tok, weight, bias, skip_tok, batch: int, spatial: Sequence[int]):
    """-> (token rows [rows_out, pad8(Cout) + Cskip], output spatial shape, (Cout, pad8(Cout)) gap descriptor)."""
    ks = tuple(weight.shape[2:])
    y = _UpCatTokens.apply(x_tok, weight, bias, skip_tok, batch, tuple(spatial), _Owner(weight))
    cout = weight.shape[1]
    return y, tuple(n * k for n, k in zip(spatial, ks)), (cout, ops.pad8(cout))
