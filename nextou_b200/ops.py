"""Tensor-level entry points of the CUDA hot path (thin wrappers over the C-ABI, include/nextou_b200.h).

Everything here works on *token-major* 2-D views ``[rows, C]`` (row = voxel / token, channels contiguous,
``stride(0)`` = row pitch) of channels-last activations; `as_tokens` gives that view without a copy when the
tensor already is channels-last.  There is no CPU or PyTorch fallback: a non-CUDA tensor or a missing
``libnextou_b200.so`` raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import NextouError, cf, check, cstream, dtype_code, ll, ptr


# ----------------------------------------------------------------------------------------------
# layout helpers
# ----------------------------------------------------------------------------------------------
def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise NextouError("nextou_b200 ops need CUDA tensors (there is no CPU fallback path)")


def channels_last(x: torch.Tensor) -> torch.Tensor:
    """Return x (N, C, *spatial) in channels-last physical layout (no copy if it already is)."""
    fmt = torch.channels_last_3d if x.dim() == 5 else torch.channels_last
    return x.contiguous(memory_format=fmt)


def pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def _token_pitch(x: torch.Tensor):
    """Row pitch (elements) if the logical (N, C, *spatial) tensor is physically token-major — channels unit-stride,
    voxels in (n, *spatial) order at a constant pitch >= C (plain channels-last or channel-padded) — else None."""
    C = x.shape[1]
    if C > 1 and x.stride(1) != 1:
        return None
    pitch = x.stride(-1)
    if pitch < C:
        return None
    expect = pitch
    for n, st in zip(reversed(x.shape[2:]), reversed(x.stride()[2:])):
        if n != 1 and st != expect:
            return None
        expect *= n
    if x.shape[0] != 1 and x.stride(0) != expect:
        return None
    return pitch


def _as_tokens_view(x: torch.Tensor) -> torch.Tensor:
    C = x.shape[1]
    T = x.numel() // C
    pitch = _token_pitch(x)
    if pitch is None:
        x = channels_last(x)
        pitch = _token_pitch(x)
        if pitch is None:  # degenerate shapes where memory_format is ambiguous
            return x.permute(0, *range(2, x.dim()), 1).reshape(T, C)
    return x.as_strided((T, C), (pitch, 1), x.storage_offset())


def _from_tokens_view(tok: torch.Tensor, batch: int, spatial: Sequence[int]) -> torch.Tensor:
    C = tok.shape[1]
    if tok.stride(1) != 1 and C > 1:
        tok = tok.contiguous()
    pitch = tok.stride(0)
    strides = [1] * (2 + len(spatial))
    acc = pitch
    for i in range(len(spatial) - 1, -1, -1):
        strides[2 + i] = acc
        acc *= spatial[i]
    strides[0] = acc
    return tok.as_strided((batch, C, *spatial), tuple(strides), tok.storage_offset())


class _AsTokens(torch.autograd.Function):
    """Re-view only: the backward is the inverse re-view (never torch's generic as_strided backward, which
    zero-fills the whole base buffer and scatters into it)."""

    @staticmethod
    def forward(ctx, x):
        ctx.meta = (x.shape[0], tuple(x.shape[2:]))
        return _as_tokens_view(x)

    @staticmethod
    def backward(ctx, g):
        return _from_tokens_view(g, *ctx.meta)


class _FromTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, batch, spatial):
        return _from_tokens_view(tok, batch, spatial)

    @staticmethod
    def backward(ctx, g):
        return _as_tokens_view(g), None, None


def as_tokens(x: torch.Tensor) -> torch.Tensor:
    """(N, C, *spatial) -> [N*prod(spatial), C] view with strides (pitch, 1); copies only if x is not token-major."""
    if x.requires_grad and torch.is_grad_enabled():
        return _AsTokens.apply(x)
    return _as_tokens_view(x)


def from_tokens(tok: torch.Tensor, batch: int, spatial: Sequence[int]) -> torch.Tensor:
    """[rows, C] token rows (strides (pitch, 1)) -> logical (N, C, *spatial) view, physically token-major."""
    if tok.requires_grad and torch.is_grad_enabled():
        return _FromTokens.apply(tok, batch, tuple(spatial))
    return _from_tokens_view(tok, batch, spatial)


def _rows_in_bounds(t: torch.Tensor, pitch: int) -> bool:
    """True if the full [rows, pitch] rectangle under the [rows, C] view `t` lies inside its storage."""
    need = t.storage_offset() + t.shape[0] * pitch
    return need <= t.untyped_storage().nbytes() // t.element_size()


def full_rows(t: torch.Tensor) -> torch.Tensor:
    """The [rows, pitch] matrix (real + padding channels) under a [rows, C] token view, or None if unavailable."""
    pitch = t.stride(0)
    if t.dim() != 2 or t.stride(1) != 1 or pitch < t.shape[1]:
        return None
    if pitch == t.shape[1]:
        return t
    if not _rows_in_bounds(t, pitch):
        return None
    return t.as_strided((t.shape[0], pitch), (pitch, 1), t.storage_offset())


def padded_like(rows: int, C: int, dtype, device) -> torch.Tensor:
    """New [rows, C] view over a [rows, pad8(C)] buffer (padding channels are don't-care)."""
    return torch.empty((rows, pad8(C)), device=device, dtype=dtype)[:, :C]


def tma_ready_bf16(tok: torch.Tensor) -> torch.Tensor:
    """[rows, C] bf16 view usable as a TMA operand (pitch % 8 == 0, 16-byte aligned); copies into a padded buffer
    only when the given view is not."""
    if (tok.dtype == torch.bfloat16 and tok.dim() == 2 and tok.stride(1) == 1 and tok.stride(0) % 8 == 0
            and tok.data_ptr() % 16 == 0):
        return tok
    out = padded_like(tok.shape[0], tok.shape[1], torch.bfloat16, tok.device)
    out.copy_(tok)
    return out


def _tok2d(t: torch.Tensor) -> torch.Tensor:
    if t.dim() != 2 or t.stride(1) != 1:
        t = t.reshape(-1, t.shape[-1]).contiguous()
    return t


def _work_dtype(t: torch.Tensor) -> torch.Tensor:
    """fp32 and bf16 are native; fp16 (nnU-Net's default autocast dtype) is widened to bf16-range-safe fp32."""
    if t.dtype in (torch.float32, torch.bfloat16):
        return t
    return t.float()


# ----------------------------------------------------------------------------------------------
# kNN graph  (torch_edge.py:139-163)
# ----------------------------------------------------------------------------------------------
def knn_normalize(tok: torch.Tensor, graphs: int, n_per_graph: int, row_map: Optional[torch.Tensor] = None,
                  normalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """L2-normalise token rows and transpose: returns (xn [graphs, C, ldn] fp32, sq [graphs, n] fp32)."""
    tok = _tok2d(_work_dtype(tok))
    _need_cuda(tok, row_map)
    C = tok.shape[1]
    ldn = (n_per_graph + 3) // 4 * 4
    xn = torch.empty((graphs, C, ldn), device=tok.device, dtype=torch.float32)
    sq = torch.empty((graphs, n_per_graph), device=tok.device, dtype=torch.float32)
    if row_map is not None:
        assert row_map.dtype == torch.int32 and row_map.numel() == graphs * n_per_graph
    else:
        assert tok.shape[0] == graphs * n_per_graph, (tok.shape, graphs, n_per_graph)
    check(_lib.lib().nextou_knn_normalize(ptr(tok), dtype_code(tok), ll(tok.stride(0)), ll(n_per_graph * tok.stride(0)),
                                          ptr(row_map), graphs, n_per_graph, C, int(normalize), ptr(xn), ldn, ptr(sq),
                                          cstream()),
          "nextou_knn_normalize")
    return xn, sq


def knn_topk(xn, sqx, yn=None, sqy=None, relpos: Optional[torch.Tensor] = None, k: int = 9, dilation: int = 1,
             want_i32: bool = True):
    """Fused distance + top-k.  Returns (idx int64 [graphs, N, k], idx32 int32 or None)."""
    if yn is None:
        yn, sqy = xn, sqx
    _need_cuda(xn, yn, relpos)
    B, C, ldn = xn.shape
    N = sqx.shape[1]
    M = sqy.shape[1]
    ldm = yn.shape[2]
    if relpos is not None:
        relpos = relpos.reshape(-1, relpos.shape[-1])
        if relpos.shape != (N, M):
            raise NextouError(f"relative_pos shape {tuple(relpos.shape)} != ({N}, {M})")
        relpos = relpos.contiguous().float()
    out = torch.empty((B, N, k), device=xn.device, dtype=torch.int64)
    out32 = torch.empty((B, N, k), device=xn.device, dtype=torch.int32) if want_i32 else None
    # algorithmic traffic / work of this launch: operands once + relpos once + int64 indices
    nbytes = 4 * B * C * (N + (M if yn is not xn else 0)) + (4 * N * M if relpos is not None else 0) + 8 * B * N * k
    L = _lib.lib()
    L.nextou_knn_topk_workspace_bytes.restype = ctypes.c_size_t
    ws_bytes = int(L.nextou_knn_topk_workspace_bytes(B, N, M))
    ws = torch.empty(ws_bytes, device=xn.device, dtype=torch.uint8) if ws_bytes else None
    with _lib.timed("knn_topk", nbytes, 2 * B * N * M * C):
        check(L.nextou_knn_topk_ws(ptr(xn), ptr(sqx), ldn, ptr(yn), ptr(sqy), ldm, ptr(relpos), B, N, M, C, k, dilation,
                                   ptr(out), ptr(out32), ptr(ws), ctypes.c_size_t(ws_bytes), cstream()), "nextou_knn_topk_ws")
    return out, out32


def knn_graph(x_tok, graphs: int, n: int, y_tok=None, m: Optional[int] = None, relpos=None, k: int = 9, dilation: int = 1,
              x_row_map=None, y_row_map=None, normalize: bool = True):
    """DenseDilatedKnnGraph on token-major inputs (deterministic `[::dilation]` branch, TE:133)."""
    with torch.no_grad():
        xn, sqx = knn_normalize(x_tok, graphs, n, x_row_map, normalize)
        if y_tok is None:
            return knn_topk(xn, sqx, None, None, relpos, k, dilation)
        yn, sqy = knn_normalize(y_tok, graphs, m, y_row_map, normalize)
        return knn_topk(xn, sqx, yn, sqy, relpos, k, dilation)


# ----------------------------------------------------------------------------------------------
# MRConv message passing (NexToU_Encoder_Decoder.py:401-409)
# ----------------------------------------------------------------------------------------------
class _MRConvGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_tok, y_tok, idx32, qmap, ymap, n, m):
        self_graph = y_tok is None
        x_tok = _tok2d(_work_dtype(x_tok))
        y2 = x_tok if self_graph else _tok2d(_work_dtype(y_tok)).to(x_tok.dtype)
        _need_cuda(x_tok, y2, idx32)
        R = idx32.shape[0] * idx32.shape[1]
        k = idx32.shape[2]
        C = x_tok.shape[1]
        out = torch.empty((x_tok.shape[0], 2 * C), device=x_tok.device, dtype=x_tok.dtype)
        arg = torch.empty((x_tok.shape[0], C), device=x_tok.device, dtype=torch.uint8)
        if qmap is None and x_tok.shape[0] != R:
            raise NextouError(f"mrconv: {x_tok.shape[0]} query rows but {R} index rows")
        check(_lib.lib().nextou_mrconv_gather_fwd(ptr(x_tok), ll(x_tok.stride(0)), ptr(y2), ll(y2.stride(0)),
                                                  dtype_code(x_tok), C, ptr(idx32), k, ptr(qmap), ptr(ymap), ll(R), n, m,
                                                  ptr(out), ll(out.stride(0)), ptr(arg), cstream()),
              "nextou_mrconv_gather_fwd")
        ctx.save_for_backward(idx32, arg, qmap, ymap)
        ctx.meta = (n, m, x_tok.shape, None if self_graph else y2.shape, x_tok.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx32, arg, qmap, ymap = ctx.saved_tensors
        n, m, xshape, yshape, dt = ctx.meta
        dout = _tok2d(dout.to(dt))
        C = xshape[1]
        dx = torch.zeros(xshape, device=dout.device, dtype=torch.float32)
        dy = dx if yshape is None else torch.zeros(yshape, device=dout.device, dtype=torch.float32)
        R = idx32.shape[0] * idx32.shape[1]
        check(_lib.lib().nextou_mrconv_gather_bwd(ptr(dout), ll(dout.stride(0)), dtype_code(dout), C, ptr(idx32),
                                                  idx32.shape[2], ptr(arg), ptr(qmap), ptr(ymap), ll(R), n, m, ptr(dx),
                                                  ll(dx.stride(0)), ptr(dy), ll(dy.stride(0)), cstream()),
              "nextou_mrconv_gather_bwd")
        return dx.to(dt), (None if yshape is None else dy.to(dt)), None, None, None, None, None


def mrconv_gather(x_tok, idx32, n: int, m: int, y_tok=None, q_row_map=None, y_row_map=None) -> torch.Tensor:
    """[rows, C] -> [rows, 2C] = interleave(x, max_j(y[nbr_j] - x)).  idx32: int32 [graphs, n, k]."""
    return _MRConvGather.apply(x_tok, y_tok, idx32.contiguous(), q_row_map, y_row_map, n, m)


# ----------------------------------------------------------------------------------------------
# pooling on token-major volumes (NexToU_Encoder_Decoder.py:511-512, 524-528, 536-549)
# ----------------------------------------------------------------------------------------------
def _geom(batch, spatial, pool):
    sp = list(spatial)
    pl = list(pool)
    if len(sp) == 2:
        sp, pl = [1] + sp, [1] + pl
    return (batch, *sp, *pl)


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, batch, spatial, pool):
        tok = _tok2d(_work_dtype(tok))
        _need_cuda(tok)
        B, D, H, W, pd, ph, pw = _geom(batch, spatial, pool)
        C = tok.shape[1]
        P = B * (D // pd) * (H // ph) * (W // pw)
        out = torch.empty((P, C), device=tok.device, dtype=tok.dtype)
        arg = torch.empty((P, C), device=tok.device, dtype=torch.uint8)
        check(_lib.lib().nextou_maxpool3d_fwd(ptr(tok), dtype_code(tok), ll(tok.stride(0)), C, B, D, H, W, pd, ph, pw,
                                              ptr(out), ll(out.stride(0)), ptr(arg), cstream()), "nextou_maxpool3d_fwd")
        ctx.save_for_backward(arg)
        ctx.meta = (B, D, H, W, pd, ph, pw, C, tok.dtype)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, dout, _darg):
        (arg,) = ctx.saved_tensors
        B, D, H, W, pd, ph, pw, C, dt = ctx.meta
        dout = _tok2d(dout.to(dt))
        dx = torch.empty((B * D * H * W, C), device=dout.device, dtype=dt)
        check(_lib.lib().nextou_maxpool3d_bwd(ptr(dout), dtype_code(dout), ll(dout.stride(0)), ptr(arg), C, B, D, H, W, pd,
                                              ph, pw, ptr(dx), ll(dx.stride(0)), cstream()), "nextou_maxpool3d_bwd")
        return dx, None, None, None


class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, batch, spatial, pool):
        tok = _tok2d(_work_dtype(tok))
        _need_cuda(tok)
        B, D, H, W, pd, ph, pw = _geom(batch, spatial, pool)
        C = tok.shape[1]
        P = B * (D // pd) * (H // ph) * (W // pw)
        out = torch.empty((P, C), device=tok.device, dtype=tok.dtype)
        check(_lib.lib().nextou_avgpool3d_fwd(ptr(tok), dtype_code(tok), ll(tok.stride(0)), C, B, D, H, W, pd, ph, pw,
                                              ptr(out), ll(out.stride(0)), cstream()), "nextou_avgpool3d_fwd")
        ctx.meta = (B, D, H, W, pd, ph, pw, C, tok.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, D, H, W, pd, ph, pw, C, dt = ctx.meta
        dout = _tok2d(dout.to(dt))
        dx = torch.empty((B * D * H * W, C), device=dout.device, dtype=dt)
        check(_lib.lib().nextou_avgpool3d_bwd(ptr(dout), dtype_code(dout), ll(dout.stride(0)), C, B, D, H, W, pd, ph, pw,
                                              ptr(dx), ll(dx.stride(0)), cstream()), "nextou_avgpool3d_bwd")
        return dx, None, None, None


class _MaxUnpool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, arg, batch, spatial, pool):
        g = _tok2d(_work_dtype(g))
        _need_cuda(g, arg)
        B, D, H, W, pd, ph, pw = _geom(batch, spatial, pool)
        C2, Carg = g.shape[1], arg.shape[1]
        out = torch.empty((B * D * H * W, C2), device=g.device, dtype=g.dtype)
        check(_lib.lib().nextou_maxunpool3d_fwd(ptr(g), dtype_code(g), ll(g.stride(0)), ptr(arg), Carg, C2, B, D, H, W, pd,
                                                ph, pw, ptr(out), ll(out.stride(0)), cstream()), "nextou_maxunpool3d_fwd")
        ctx.save_for_backward(arg)
        ctx.meta = (B, D, H, W, pd, ph, pw, C2, Carg, g.dtype, g.shape[0])
        return out

    @staticmethod
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        B, D, H, W, pd, ph, pw, C2, Carg, dt, P = ctx.meta
        dout = _tok2d(dout.to(dt))
        dg = torch.empty((P, C2), device=dout.device, dtype=dt)
        check(_lib.lib().nextou_maxunpool3d_bwd(ptr(dout), dtype_code(dout), ll(dout.stride(0)), ptr(arg), Carg, C2, B, D,
                                                H, W, pd, ph, pw, ptr(dg), ll(dg.stride(0)), cstream()),
              "nextou_maxunpool3d_bwd")
        return dg, None, None, None, None


def maxpool_tokens(tok, batch, spatial, pool):
    """Non-overlapping max pool; spatial = full-resolution (D,H,W) or (H,W).  Returns (pooled, uint8 child arg)."""
    return _MaxPool.apply(tok, batch, tuple(spatial), tuple(pool))


def avgpool_tokens(tok, batch, spatial, pool):
    return _AvgPool.apply(tok, batch, tuple(spatial), tuple(pool))


def maxunpool_tokens(g, arg, batch, spatial, pool):
    """Scatter pooled rows back to full resolution: channel j uses arg[:, j % arg.shape[1]] (ED:536)."""
    return _MaxUnpool.apply(g, arg, batch, tuple(spatial), tuple(pool))


# ----------------------------------------------------------------------------------------------
# BTI / TI loss (loss/bti_loss.py:76-145)
# ----------------------------------------------------------------------------------------------
_TGT_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.int64: 2, torch.uint8: 3, torch.int32: 4}


def _prep_logits(x):
    """logits (b, c, *spatial) -> (tensor, (batch, class, voxel) element strides); copies only if the spatial
    dims do not flatten to one stride (NCDHW-contiguous and channels-last both flatten)."""
    x = _work_dtype(x)
    try:
        xv = x.view(x.shape[0], x.shape[1], -1)
    except RuntimeError:
        x = x.contiguous()
        xv = x.view(x.shape[0], x.shape[1], -1)
    return x, tuple(xv.stride())


def _prep_target(y):
    y = y.reshape(y.shape[0], -1)
    if y.dtype not in _TGT_CODE:
        y = y.float()
    return y.contiguous()


def bti_labels(logits: torch.Tensor) -> torch.Tensor:
    """argmax over classes as uint8 (b, *spatial) (bti_loss.py:132-134)."""
    _need_cuda(logits)
    x, (sb, sc, sv) = _prep_logits(logits)
    B, NC = x.shape[:2]
    V = x[0, 0].numel()
    labels = torch.empty((B, *x.shape[2:]), device=x.device, dtype=torch.uint8)
    check(_lib.lib().nextou_bti_argmax_ce(ptr(x), dtype_code(x), ll(sb), ll(sc), ll(sv), B, NC, ll(V), ptr(None), 0,
                                          ptr(labels), ptr(None), cstream()), "nextou_bti_argmax_ce")
    return labels


def bti_critical_map(labels: torch.Tensor, mask_a, mask_c, inclusion, connectivity: int, min_thick: int = 1) -> torch.Tensor:
    """uint8 critical-voxel map from uint8 labels (b, *spatial) and the interaction table (host int lists)."""
    _need_cuda(labels)
    labels = labels.contiguous()
    dim = labels.dim() - 1
    B = labels.shape[0]
    D, H, W = (1, *labels.shape[1:]) if dim == 2 else labels.shape[1:]
    n = len(mask_a)
    A = (ctypes.c_uint32 * max(n, 1))(*[int(v) & 0xFFFFFFFF for v in mask_a])
    Cm = (ctypes.c_uint32 * max(n, 1))(*[int(v) & 0xFFFFFFFF for v in mask_c])
    inc = (ctypes.c_uint8 * max(n, 1))(*[1 if v else 0 for v in inclusion])
    crit = torch.empty_like(labels)
    check(_lib.lib().nextou_bti_critical_map(ptr(labels), B, D, H, W, dim, A, Cm, inc, n, connectivity, min_thick,
                                             ptr(crit), cstream()), "nextou_bti_critical_map")
    return crit


class _BTILoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, table, connectivity, min_thick):
        _need_cuda(logits, target)
        x, (sb, sc, sv) = _prep_logits(logits)
        y = _prep_target(target)
        B, NC = x.shape[:2]
        V = x[0, 0].numel()
        L = _lib.lib()
        labels = torch.empty((B, *x.shape[2:]), device=x.device, dtype=torch.uint8)
        ce = torch.empty((B, V), device=x.device, dtype=torch.float64)
        check(L.nextou_bti_argmax_ce(ptr(x), dtype_code(x), ll(sb), ll(sc), ll(sv), B, NC, ll(V), ptr(y), _TGT_CODE[y.dtype],
                                     ptr(labels), ptr(ce), cstream()), "nextou_bti_argmax_ce")
        crit = bti_critical_map(labels, *table, connectivity, min_thick)
        L.nextou_bti_masked_sum_workspace_bytes.restype = ctypes.c_size_t
        ws = torch.empty(L.nextou_bti_masked_sum_workspace_bytes(B) // 8, device=x.device, dtype=torch.float64)
        out = torch.empty((), device=x.device, dtype=torch.float64)
        check(L.nextou_bti_masked_sum(ptr(ce), ptr(crit), B, ll(V), ptr(ws), ptr(out), cstream()), "nextou_bti_masked_sum")
        ctx.save_for_backward(x, y, crit)
        ctx.meta = (sb, sc, sv, B, NC, V, logits.dtype)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, y, crit = ctx.saved_tensors
        sb, sc, sv, B, NC, V, in_dtype = ctx.meta
        g = gout.to(torch.float64).contiguous()
        dx = torch.empty_like(x)  # preserves strides
        if dx.stride() != x.stride():
            dx = torch.empty_strided(x.shape, x.stride(), device=x.device, dtype=x.dtype)
        check(_lib.lib().nextou_bti_ce_bwd(ptr(x), dtype_code(x), ll(sb), ll(sc), ll(sv), B, NC, ll(V), ptr(y),
                                           _TGT_CODE[y.dtype], ptr(crit), ptr(g), ptr(dx), ll(sb), ll(sc), ll(sv), cstream()),
              "nextou_bti_ce_bwd")
        return dx.to(in_dtype), None, None, None, None


def bti_loss(logits, target, mask_a, mask_c, inclusion, connectivity: int, min_thick: int = 1) -> torch.Tensor:
    """fp64 scalar: mean_b sum_v CE(logits, target)[b, v] * critical[b, v]  (bti_loss.py:141-143)."""
    return _BTILoss.apply(logits, target, (list(mask_a), list(mask_c), list(inclusion)), connectivity, min_thick)


SEG_LOSS_CLASSES = (2, 3, 4, 5, 6, 7, 8, 14, 16, 19)   # instantiated class counts of csrc/dsloss.cu


def _empty_like_strided(x: torch.Tensor) -> torch.Tensor:
    """Uninitialised tensor with x's shape and strides whose storage also covers the channel padding of the LAST token row
    (torch.empty_strided stops at the last addressed element, which makes the full [rows, pitch] rectangle unaddressable
    and forces the consumers of the gradient to copy it)."""
    need = 1 + sum((n - 1) * st for n, st in zip(x.shape, x.stride()))
    pitch = min((st for st in x.stride() if st > 1), default=1)
    need = (need + pitch - 1) // pitch * pitch
    return torch.empty(need, device=x.device, dtype=x.dtype).as_strided(x.shape, x.stride())


class _SegLoss(torch.autograd.Function):
    """w_ce * CE + w_dice * SoftDice + w_ti * (B)TI of one deep-supervision scale from two passes over the logits
    (csrc/dsloss.cu).  cfg = (w_ce, w_dice, w_ti, batch_dice, do_bg, smooth, ddp, table | None, connectivity, min_thick)."""

    @staticmethod
    def forward(ctx, logits, target, cfg):
        _need_cuda(logits, target)
        w_ce, w_dice, w_ti, batch_dice, do_bg, smooth, ddp, table, connectivity, min_thick = cfg
        x, (sb, sc, sv) = _prep_logits(logits)
        y = _prep_target(target)
        B, NC = x.shape[:2]
        V = x[0, 0].numel()
        L = _lib.lib()
        dev = x.device
        nblk = ctypes.c_int(0)
        check(L.nextou_dsloss_plan(ll(V), B, ctypes.byref(nblk)), "nextou_dsloss_plan")
        width = 3 * NC + 1
        labels = torch.empty((B, *x.shape[2:]), device=dev, dtype=torch.uint8)
        ce = torch.empty((B, V), device=dev, dtype=torch.float64)
        partial = torch.empty((B, nblk.value, width), device=dev, dtype=torch.float64)
        sums = torch.empty((B, width), device=dev, dtype=torch.float64)
        check(L.nextou_dsloss_stats(ptr(x), dtype_code(x), ll(sb), ll(sc), ll(sv), B, NC, ll(V), ptr(y), _TGT_CODE[y.dtype],
                                    ptr(labels), ptr(ce), ptr(partial), ptr(sums), cstream()), "nextou_dsloss_stats")
        crit = None
        ti = None
        if w_ti != 0:
            crit = bti_critical_map(labels, *table, connectivity, min_thick)
            L.nextou_bti_masked_sum_workspace_bytes.restype = ctypes.c_size_t
            ws = torch.empty(L.nextou_bti_masked_sum_workspace_bytes(B) // 8, device=dev, dtype=torch.float64)
            ti = torch.empty((), device=dev, dtype=torch.float64)
            check(L.nextou_bti_masked_sum(ptr(ce), ptr(crit), B, ll(V), ptr(ws), ptr(ti), cstream()), "nextou_bti_masked_sum")
        # per-class algebra on [B, NC] doubles (MemoryEfficientSoftDiceLoss.forward; CrossEntropyLoss mean reduction): one launch
        pooled = None
        grad_world = 1.0
        if batch_dice and ddp and torch.distributed.is_available() and torch.distributed.is_initialized():
            pooled = sums[:, :3 * NC].sum(0)
            torch.distributed.all_reduce(pooled)
            # upstream gathers with AllGatherGrad, whose backward all-reduces (SUM) the incoming gradient: every rank's
            # loss depends on every rank's sums, so the local derivative is world_size x d(dice)/d(sums); DDP's gradient
            # averaging then yields the gradient of the global-batch Dice (same as torch.distributed.nn all_gather)
            grad_world = float(torch.distributed.get_world_size())
        total = torch.empty((), device=dev, dtype=torch.float64)
        coef = torch.empty((2, B, NC), device=dev, dtype=torch.float64)
        cd = ctypes.c_double
        check(L.nextou_dsloss_finish(ptr(sums), ptr(pooled), B, NC, ll(V), cd(w_ce), cd(w_dice), cd(w_ti), ptr(ti), int(batch_dice),
                                     int(do_bg), cd(smooth), cd(grad_world), ptr(total), ptr(coef), cstream()), "nextou_dsloss_finish")
        ctx.save_for_backward(x, y, crit if crit is not None else labels, coef)
        ctx.meta = (sb, sc, sv, B, NC, V, logits.dtype, crit is not None, w_ce, w_ti)
        return total

    @staticmethod
    def backward(ctx, gout):
        x, y, crit, coef = ctx.saved_tensors
        sb, sc, sv, B, NC, V, in_dtype, has_crit, w_ce, w_ti = ctx.meta
        g = gout.to(torch.float64).contiguous()
        L = _lib.lib()
        coef32 = torch.empty((2, B, NC), device=x.device, dtype=torch.float32)
        scal = torch.empty(2, device=x.device, dtype=torch.float32)
        cd = ctypes.c_double
        check(L.nextou_dsloss_scale(ptr(coef), ptr(g), B, NC, ll(V), cd(w_ce), cd(w_ti), ptr(coef32), ptr(scal), cstream()),
              "nextou_dsloss_scale")
        dx = _empty_like_strided(x)
        check(L.nextou_dsloss_bwd(ptr(x), dtype_code(x), ll(sb), ll(sc), ll(sv), B, NC, ll(V), ptr(y),
                                  _TGT_CODE[y.dtype], ptr(crit if has_crit else None), ptr(coef32[0]), ptr(coef32[1]),
                                  ptr(scal), ptr(dx), ll(sb), ll(sc), ll(sv), cstream()), "nextou_dsloss_bwd")
        return dx.to(in_dtype), None, None


def seg_loss(logits, target, w_ce, w_dice, w_ti, batch_dice, do_bg, smooth, ddp, table, connectivity, min_thick):
    """Fused CE + soft Dice + (B)TI for one output scale -> fp64 scalar (compound_bti_loss.py:33-61)."""
    return _SegLoss.apply(logits, target, (float(w_ce), float(w_dice), float(w_ti), bool(batch_dice), bool(do_bg), float(smooth),
                                           bool(ddp), table, connectivity, min_thick))


# ----------------------------------------------------------------------------------------------
# tcgen05 GEMM engine (csrc/gemm_tcgen05.cu)
# ----------------------------------------------------------------------------------------------
def gemm_bf16_tn(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, n: Optional[int] = None,
                 out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None,
                 slope: float = 1.0) -> torch.Tensor:
    """out[M, ldc] = a[M, K] @ b[N, K]^T (+ bias); a, b bf16 with unit inner stride and row pitches that are multiples of
    8 elements; returns the PADDED [M, pad8(N)] matrix (columns >= N are zero), or writes into `out`."""
    _need_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0] if n is None else n
    if out is None:
        out = torch.empty((M, pad8(N)), device=a.device, dtype=out_dtype)
    bias32 = None if bias is None else bias.detach().float().contiguous()
    scale32 = None if scale is None else scale.detach().float().contiguous()
    with _lib.timed("gemm_tcgen05", 2 * M * (K + N) + 2 * N * K, 2 * M * N * K):
        check(_lib.lib().nextou_gemm_bf16_tn_affine(ptr(a), ll(a.stride(0)), ptr(b), ll(b.stride(0)), ptr(out), ll(out.stride(0)),
                                                    M, N, K, ptr(scale32), ptr(bias32), cf(slope), dtype_code(out), cstream()),
              "nextou_gemm_bf16_tn")
    return out


def pack_conv_weight(w: torch.Tensor, transpose_flip: bool = False, transpose: bool = False) -> torch.Tensor:
    """(Cout, Cin, *k) conv weight -> bf16 [Cout, taps * cin_pad] (taps in (kd, kh, kw) order, Cin zero-padded to a
    multiple of 64).  transpose_flip=True packs the data-gradient operator of a stride-1 convolution: [Cin, flipped
    taps * cout_pad]; transpose=True the un-flipped one the strided data-gradient kernel takes."""
    nd = w.dim() - 2
    if transpose_flip:
        w = w.transpose(0, 1).flip(dims=tuple(range(2, 2 + nd)))
    elif transpose:
        w = w.transpose(0, 1)
    co, ci = w.shape[:2]
    taps = 1
    for s in w.shape[2:]:
        taps *= s
    ci_pad = (ci + 63) // 64 * 64
    wp = w.reshape(co, ci, taps).permute(0, 2, 1)
    wp = torch.nn.functional.pad(wp, (0, ci_pad - ci))
    return wp.reshape(co, taps * ci_pad).to(torch.bfloat16).contiguous()


class PackEntry:
    """Persistent bf16 operand packs of ONE master weight in ONE layout (kept on the Parameter object, so their lifetime and
    identity are the parameter's).  `version` is the parameter's in-place version counter at packing time: any in-place
    update by torch (optimizer.step, load_state_dict, DDP broadcast) invalidates the packs; nextou_b200.optim.FusedSGD
    updates the weights AND rewrites the packs in one pass, so they stay valid from step to step."""
    __slots__ = ("a", "b", "version", "ptr", "meta")

    def __init__(self, meta):
        self.a = self.b = None
        self.version = self.ptr = -1
        self.meta = meta          # (R, Cc, taps, groups, flip_b, pa, pb, gap_lo, gap_hi)


PACK_CACHE = True   # False: re-pack on every call (A/B measurements)


def pack_weight_pair(w: torch.Tensor, conv: bool, flip_b: bool = False, groups: int = 1, want_b: bool = True,
                     gap: Optional[Tuple[int, int]] = None, owner: Optional[torch.Tensor] = None):
    """Master weight (R, Cc/groups, *k) -> (A [R, taps*pa] bf16, Bt [Cc, taps*pb] bf16 | None) in ONE kernel launch
    (csrc/pool.cu::pack_weight_kernel): A is the forward operand, Bt the data-gradient operand (taps flipped when
    flip_b).  conv=True pads the channel axes to multiples of 64 (taps are concatenated along K), else to 8 (TMA pitch).
    gap=(lo, hi): zero input channels [lo, hi) are inserted (the decoder's [up | gap | skip] concatenation layout).
    owner: the nn.Parameter `w` is (a view of): its packs are kept on it and re-used until the parameter changes."""
    _need_cuda(w)
    w = _work_dtype(w.detach()).contiguous()
    R, cpg = w.shape[:2]
    Cc = cpg * groups
    taps = 1
    for k in w.shape[2:]:
        taps *= k
    glo, ghi = (0, 0) if gap is None else (int(gap[0]), int(gap[1]))
    Cp = Cc + ghi - glo
    pad = (lambda c: (c + 63) // 64 * 64) if conv else pad8
    pa, pb = pad(Cp), pad(R)
    entry = None
    if PACK_CACHE and owner is not None and w.dtype == torch.float32 and w.data_ptr() == owner.data_ptr():
        cache = owner.__dict__.setdefault("_nextou_packs", {})
        key = (conv, bool(flip_b), groups, glo, ghi)
        entry = cache.get(key)
        if entry is None:
            entry = cache[key] = PackEntry((R, Cc, taps, groups, int(flip_b), pa, pb, glo, ghi))
        if (entry.version == owner._version and entry.ptr == w.data_ptr() and entry.a is not None
                and (entry.b is not None or not want_b)):
            return entry.a, (entry.b if want_b else None)
    if entry is not None and entry.a is not None and entry.ptr == w.data_ptr():
        a = entry.a                                   # re-pack in place: FusedSGD / CUDA graphs hold these addresses
        b = entry.b if entry.b is not None else (torch.empty((Cp, taps * pb), device=w.device, dtype=torch.bfloat16) if want_b else None)
    else:
        a = torch.empty((R, taps * pa), device=w.device, dtype=torch.bfloat16)
        b = torch.empty((Cp, taps * pb), device=w.device, dtype=torch.bfloat16) if want_b else None
    check(_lib.lib().nextou_pack_weight_gap(ptr(w), dtype_code(w), R, Cc, taps, groups, int(flip_b), glo, ghi, ptr(a), pa, ptr(b), pb,
                                            cstream()), "nextou_pack_weight")
    if entry is not None:
        entry.a, entry.b, entry.version, entry.ptr = a, b, owner._version, w.data_ptr()
    return a, (b if want_b else None)


CONV_HALO = True  # use the halo-reuse kernel (csrc/conv_tcgen05.cu) whenever kh, kw are 1 or 3


def conv_ndhwc_bf16(x_tok: torch.Tensor, batch: int, spatial: Sequence[int], cin: int, wpack: torch.Tensor, cout: int,
                    ksize: Sequence[int], bias: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16,
                    halo: Optional[bool] = None, scale: Optional[torch.Tensor] = None, slope: float = 1.0) -> torch.Tensor:
    """Stride-1 'same' convolution on a token-major bf16 volume [B*prod(spatial), ldx] -> padded [.., pad8(cout)]."""
    _need_cuda(x_tok, wpack)
    assert x_tok.dtype == torch.bfloat16 and x_tok.stride(1) == 1 and wpack.dtype == torch.bfloat16
    sp = list(spatial)
    ks = list(ksize)
    if len(sp) == 2:
        sp, ks = [1] + sp, [1] + ks
    D, H, W = sp
    out = torch.empty((batch * D * H * W, pad8(cout)), device=x_tok.device, dtype=out_dtype)
    bias32 = None if bias is None else bias.detach().float().contiguous()
    scale32 = None if scale is None else scale.detach().float().contiguous()
    use_halo = (CONV_HALO if halo is None else halo) and ks[1] in (1, 3) and ks[2] in (1, 3)
    V = batch * D * H * W
    # algorithmic work / traffic: 2*V*Cin*Cout*taps flops; input + output + weights once, bf16
    flops = 2 * V * cin * cout * ks[0] * ks[1] * ks[2]
    nbytes = 2 * V * (cin + cout) + 2 * cin * cout * ks[0] * ks[1] * ks[2]
    L = _lib.lib()
    with _lib.timed("conv_halo_tcgen05" if use_halo else "conv_pertap_tcgen05", nbytes, flops):
        if use_halo:
            check(L.nextou_conv3d_ndhwc_halo_fwd_affine(ptr(x_tok), ll(x_tok.stride(0)), batch, D, H, W, cin, ptr(wpack), cout,
                                                        ks[0], ks[1], ks[2], ptr(scale32), ptr(bias32), cf(slope), ptr(out),
                                                        ll(out.stride(0)), dtype_code(out), cstream()),
                  "nextou_conv3d_ndhwc_halo_fwd")
        else:
            check(L.nextou_conv3d_ndhwc_strided_fwd_affine(ptr(x_tok), ll(x_tok.stride(0)), batch, D, H, W, cin, ptr(wpack), cout,
                                                           ks[0], ks[1], ks[2], 1, 1, 1, ks[0] // 2, ks[1] // 2, ks[2] // 2,
                                                           ptr(scale32), ptr(bias32), cf(slope), ptr(out), ll(out.stride(0)),
                                                           dtype_code(out), cstream()), "nextou_conv3d_ndhwc_fwd")
    return out


def _k3(spatial, ksize):
    sp, ks = list(spatial), list(ksize)
    while len(sp) < 3:
        sp, ks = [1] + sp, [1] + ks
    return sp, ks


def conv_small_supported(cin: int, cout: int, ksize: Sequence[int], weight: torch.Tensor) -> bool:
    """True for the network's first convolution (image modalities -> 33 channels): csrc/conv_small.cu covers it."""
    _, ks = _k3([1] * len(ksize), ksize)
    return bool(weight.dtype == torch.float32 and
                _lib.lib().nextou_conv3d_small_cin_supported(int(cin), int(cout), ks[0], ks[1], ks[2], ll(pad8(cout))))


def conv_small_fwd(x_tok: torch.Tensor, batch: int, spatial: Sequence[int], weight: torch.Tensor, bias: Optional[torch.Tensor]):
    """Stride-1 'same' convolution with Cin <= 4 (csrc/conv_small.cu) on bf16 token rows -> padded [rows, pad8(cout)] bf16."""
    _need_cuda(x_tok, weight)
    assert x_tok.dtype == torch.bfloat16 and (x_tok.stride(1) == 1 or x_tok.shape[1] == 1)
    cout, cin = weight.shape[:2]
    sp, ks = _k3(spatial, weight.shape[2:])
    V = batch * sp[0] * sp[1] * sp[2]
    out = torch.empty((V, pad8(cout)), device=x_tok.device, dtype=torch.bfloat16)
    w32 = weight.detach().contiguous()
    b32 = None if bias is None else bias.detach().float().contiguous()
    taps = ks[0] * ks[1] * ks[2]
    with _lib.timed("conv_small", 2 * V * (cin + cout) + 4 * cin * cout * taps, 2 * V * cin * cout * taps):
        check(_lib.lib().nextou_conv3d_small_cin_fwd(ptr(x_tok), ll(x_tok.stride(0)), batch, *sp, cin, ptr(w32), cout, *ks, ptr(b32),
                                                     ptr(out), ll(out.stride(0)), cstream()), "nextou_conv3d_small_cin_fwd")
    return out


def conv_small_wgrad(dy_tok: torch.Tensor, x_tok: torch.Tensor, batch: int, spatial: Sequence[int], cin: int, cout: int,
                     ksize: Sequence[int]) -> torch.Tensor:
    """fp32 weight gradient (Cout, Cin, *ksize) of the small-Cin convolution."""
    _need_cuda(dy_tok, x_tok)
    assert dy_tok.dtype == torch.bfloat16 and x_tok.dtype == torch.bfloat16 and dy_tok.stride(0) % 8 == 0
    sp, ks = _k3(spatial, ksize)
    taps = ks[0] * ks[1] * ks[2]
    dw = torch.zeros((cout, taps, cin), device=x_tok.device, dtype=torch.float32)
    V = batch * sp[0] * sp[1] * sp[2]
    with _lib.timed("conv_small", 2 * V * (cin + cout) + 4 * cin * cout * taps, 2 * V * cin * cout * taps):
        check(_lib.lib().nextou_conv3d_small_cin_wgrad(ptr(dy_tok), ll(dy_tok.stride(0)), ptr(x_tok), ll(x_tok.stride(0)), batch, *sp,
                                                       cin, cout, *ks, ptr(dw), cin, cstream()), "nextou_conv3d_small_cin_wgrad")
    return dw.permute(0, 2, 1).reshape(cout, cin, *ksize)


def _geom3(spatial, *lists):
    """Left-pad spatial / kernel / stride / padding lists to 3-D."""
    sp = list(spatial)
    out = [list(v) for v in lists]
    while len(sp) < 3:
        sp = [1] + sp
        out = [[1 if i < 2 else 0] + v for i, v in enumerate(out)]   # kernel 1, stride 1, padding 0
    return (sp, *out)


def conv_strided_fwd_bf16(x_tok: torch.Tensor, batch: int, spatial: Sequence[int], cin: int, wpack: torch.Tensor, cout: int,
                          ksize: Sequence[int], stride: Sequence[int], padding: Sequence[int],
                          bias: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16,
                          scale: Optional[torch.Tensor] = None, slope: float = 1.0):
    """Strided convolution on a token-major bf16 volume -> (padded [rows_out, pad8(cout)] matrix, output spatial shape)."""
    _need_cuda(x_tok, wpack)
    assert x_tok.dtype == torch.bfloat16 and x_tok.stride(1) == 1 and wpack.dtype == torch.bfloat16
    sp, ks, st, pd = _geom3(spatial, ksize, stride, padding)
    osp = [(sp[i] + 2 * pd[i] - ks[i]) // st[i] + 1 for i in range(3)]
    V = batch * osp[0] * osp[1] * osp[2]
    out = torch.empty((V, pad8(cout)), device=x_tok.device, dtype=out_dtype)
    bias32 = None if bias is None else bias.detach().float().contiguous()
    taps = ks[0] * ks[1] * ks[2]
    with _lib.timed("conv_pertap_tcgen05", 2 * batch * sp[0] * sp[1] * sp[2] * cin + 2 * V * cout + 2 * cin * cout * taps,
                    2 * V * cin * cout * taps):
        scale32 = None if scale is None else scale.detach().float().contiguous()
        check(_lib.lib().nextou_conv3d_ndhwc_strided_fwd_affine(ptr(x_tok), ll(x_tok.stride(0)), batch, *sp, cin, ptr(wpack), cout,
                                                                *ks, *st, *pd, ptr(scale32), ptr(bias32), cf(slope), ptr(out),
                                                                ll(out.stride(0)), dtype_code(out), cstream()),
              "nextou_conv3d_ndhwc_strided_fwd")
    return out, tuple(osp[3 - len(spatial):])


def conv_strided_dgrad_bf16(dy_tok: torch.Tensor, batch: int, out_spatial: Sequence[int], cout: int, wpack_t: torch.Tensor,
                            cin: int, ksize: Sequence[int], stride: Sequence[int], padding: Sequence[int],
                            in_spatial: Sequence[int], bias: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16,
                            out: Optional[torch.Tensor] = None, store_cols: Optional[int] = None) -> torch.Tensor:
    """Data gradient of a strided convolution / forward of a kernel == stride transposed convolution:
    dy [rows_out, cout] on `out_spatial` -> padded [rows_in, pad8(cin)] on `in_spatial` (or written into `out`)."""
    _need_cuda(dy_tok, wpack_t)
    assert dy_tok.dtype == torch.bfloat16 and dy_tok.stride(1) == 1 and wpack_t.dtype == torch.bfloat16
    osp, ks, st, pd = _geom3(out_spatial, ksize, stride, padding)
    isp = list(in_spatial)
    while len(isp) < 3:
        isp = [1] + isp
    V = batch * isp[0] * isp[1] * isp[2]
    if out is None:
        out = torch.empty((V, pad8(cin)), device=dy_tok.device, dtype=out_dtype)
    bias32 = None if bias is None else bias.detach().float().contiguous()
    taps = ks[0] * ks[1] * ks[2]
    Vo = batch * osp[0] * osp[1] * osp[2]
    with _lib.timed("conv_pertap_tcgen05", 2 * Vo * cout + 2 * V * cin + 2 * cin * cout * taps, 2 * Vo * cin * cout * taps):
        cols = int(out.stride(0)) if store_cols is None else int(store_cols)
        check(_lib.lib().nextou_conv3d_ndhwc_strided_dgrad_cols(ptr(dy_tok), ll(dy_tok.stride(0)), batch, *osp, cout, ptr(wpack_t),
                                                                cin, *ks, *st, *pd, ptr(bias32), ptr(out), ll(out.stride(0)), cols,
                                                                *isp, dtype_code(out), cstream()),
              "nextou_conv3d_ndhwc_strided_dgrad")
    return out


def convtranspose_fwd_bf16(x_tok: torch.Tensor, batch: int, spatial: Sequence[int], cin: int, wpack_t: torch.Tensor, cout: int,
                           ksize: Sequence[int], bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                           store_cols: Optional[int] = None) -> torch.Tensor:
    """Forward of a kernel == stride transposed convolution (ED:273-276, 321): x [rows_in, cin] on `spatial` -> padded
    [rows_out, pad8(cout)] on spatial * ksize (or written into the first `store_cols` columns of `out`).  One scatter GEMM over the
    input voxels (csrc/gemm_tcgen05.cu) when the class segment fits a tile, else one per-tap launch per output parity class."""
    _need_cuda(x_tok, wpack_t)
    assert x_tok.dtype == torch.bfloat16 and x_tok.stride(1) == 1 and wpack_t.dtype == torch.bfloat16
    sp, ks = _k3(spatial, ksize)
    osp = [n * k for n, k in zip(sp, ks)]
    Vo = batch * osp[0] * osp[1] * osp[2]
    if out is None:
        out = torch.empty((Vo, pad8(cout)), device=x_tok.device, dtype=torch.bfloat16)
    cols = int(out.stride(0)) if store_cols is None else int(store_cols)
    L = _lib.lib()
    if not L.nextou_convtranspose_scatter_fwd_supported(cout, cols, *ks) or out.dtype != torch.bfloat16:
        zero = (0,) * len(ksize)
        return conv_strided_dgrad_bf16(x_tok, batch, spatial, cin, wpack_t, cout, ksize, ksize, zero,
                                       tuple(n * k for n, k in zip(spatial, ksize)), bias, out=out, store_cols=store_cols)
    bias32 = None if bias is None else bias.detach().float().contiguous()
    V = batch * sp[0] * sp[1] * sp[2]
    taps = ks[0] * ks[1] * ks[2]
    with _lib.timed("convtranspose_scatter", 2 * V * cin + 2 * Vo * cout + 2 * cin * cout * taps, 2 * V * cin * cout * taps):
        check(L.nextou_convtranspose_scatter_fwd(ptr(x_tok), ll(x_tok.stride(0)), batch, *sp, cin, ptr(wpack_t), cout, *ks, ptr(bias32),
                                                 ptr(out), ll(out.stride(0)), cols, cstream()), "nextou_convtranspose_scatter_fwd")
    return out


def conv_strided_wgrad_bf16(dense_tok: torch.Tensor, strided_tok: torch.Tensor, batch: int, dense_spatial: Sequence[int],
                            strided_spatial: Sequence[int], c_strided: int, c_dense: int, ksize: Sequence[int],
                            stride: Sequence[int], padding: Sequence[int], side: Optional["side_launch"] = None):
    """fp32 [c_dense, taps, c_strided] = sum_i dense[i][m] * strided[i*s + tap - pad][n] (csrc/gemm_tcgen05.cu).
    With `side`: launched on the side stream, returns a finisher (see conv_wgrad_bf16)."""
    _need_cuda(dense_tok, strided_tok)
    assert dense_tok.dtype == torch.bfloat16 and strided_tok.dtype == torch.bfloat16
    dsp, ks, st, pd = _geom3(dense_spatial, ksize, stride, padding)
    ssp = list(strided_spatial)
    while len(ssp) < 3:
        ssp = [1] + ssp
    taps = ks[0] * ks[1] * ks[2]
    cs = (c_strided + 3) // 4 * 4
    # split-K accumulator: allocated on the current stream, zero-filled on the side stream when there is one
    dw = (torch.empty if side is not None else torch.zeros)((c_dense, taps, cs), device=dense_tok.device, dtype=torch.float32)
    V = batch * dsp[0] * dsp[1] * dsp[2]
    L = _lib.lib()
    # down-sampling 3x3 convolutions with a small Cin: halo reuse over the four parity planes of the input (csrc/conv_tcgen05.cu)
    planes = bool(L.nextou_conv3d_ndhwc_planes_wgrad_supported(c_strided, *ks, *st, *pd)) and min(ssp[1], ssp[2]) >= 2
    with (side if side is not None else _null_ctx()):
        if side is not None:
            dw.zero_()
        with _lib.timed("wgrad_planes_tcgen05" if planes else "wgrad_tcgen05",
                        2 * V * c_dense + 2 * batch * ssp[0] * ssp[1] * ssp[2] * c_strided + 4 * c_dense * c_strided * taps,
                        2 * V * c_dense * c_strided * taps):
            if planes:
                check(L.nextou_conv3d_ndhwc_planes_wgrad(ptr(dense_tok), ll(dense_tok.stride(0)), ptr(strided_tok),
                                                         ll(strided_tok.stride(0)), batch, *dsp, *ssp, c_strided, c_dense, ks[0],
                                                         st[0], pd[0], ptr(dw), cs, cstream()), "nextou_conv3d_ndhwc_planes_wgrad")
            else:
                check(L.nextou_conv3d_ndhwc_strided_wgrad(ptr(dense_tok), ll(dense_tok.stride(0)), ptr(strided_tok),
                                                          ll(strided_tok.stride(0)), batch, *dsp, *ssp, c_strided, c_dense,
                                                          *ks, *st, *pd, ptr(dw), cs, cstream()),
                      "nextou_conv3d_ndhwc_strided_wgrad")
    finish = lambda: dw[:, :, :c_strided]
    return finish if side is not None else finish()


OVERLAP_WGRAD = True   # run a layer's weight-gradient kernel on a side stream, concurrently with its data-gradient kernel
# ... and do not make the main stream wait for it at the end of the layer's backward when autograd will merely take the gradient
# tensor as param.grad (no kernel reads it before the end of the backward pass): native._complete_wgrad
DEFER_WGRAD_JOIN = os.environ.get("NEXTOU_DEFER_WGRAD_JOIN", "1") != "0"
_SIDE_STREAMS = {}


class side_launch:
    """Fork / join helper: kernel launches inside the `with` block go to a per-device side stream that first waits for the
    current stream; `join()` makes the current stream wait for them.  Nothing may be ALLOCATED inside the block (the
    caching allocator ties a block to the stream it was allocated on): callers allocate outputs before entering.  Inside a
    CUDA-graph capture the fork / join become graph edges, so the two kernels are parallel branches of the step graph."""

    def __init__(self, device):
        self.cur = torch.cuda.current_stream(device)
        key = (device.index if device.index is not None else torch.cuda.current_device())
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = torch.cuda.Stream(device)
        self.side = _SIDE_STREAMS[key]
        self._ctx = None

    def __enter__(self):
        self.side.wait_stream(self.cur)
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._ctx.__exit__(*exc)
        return False

    def join(self):
        self.cur.wait_stream(self.side)


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def conv_wgrad_bf16(dy_tok: torch.Tensor, x_tok: torch.Tensor, batch: int, spatial: Sequence[int], cin: int, cout: int,
                    ksize: Sequence[int], halo: Optional[bool] = None, side: Optional[side_launch] = None):
    """fp32 weight gradient (Cout, Cin, *ksize) of a stride-1 'same' convolution (1x1: ksize of ones) from bf16
    token-major dY [rows, Cout] and X [rows, Cin] (csrc/gemm_tcgen05.cu, MN-major tcgen05 operands).
    With `side` the kernel is launched on the side stream and a FINISHER is returned: call side.join(), then the finisher."""
    _need_cuda(dy_tok, x_tok)
    assert dy_tok.dtype == torch.bfloat16 and x_tok.dtype == torch.bfloat16
    sp = list(spatial)
    ks = list(ksize)
    while len(sp) < 3:
        sp, ks = [1] + sp, [1] + ks
    D, H, W = sp
    taps = ks[0] * ks[1] * ks[2]
    cs = (cin + 3) // 4 * 4                                     # 16-byte rows: the split-K reduction uses vector reds
    # split-K accumulator: allocated on the current stream, zero-filled on the side stream when there is one
    dw = (torch.empty if side is not None else torch.zeros)((cout, taps, cs), device=x_tok.device, dtype=torch.float32)
    use_halo = (CONV_HALO if halo is None else halo) and ks[1] in (1, 3) and ks[2] in (1, 3) and taps > 1
    fn = _lib.lib().nextou_conv3d_ndhwc_halo_wgrad if use_halo else _lib.lib().nextou_conv3d_ndhwc_wgrad
    V = batch * D * H * W
    with (side if side is not None else _null_ctx()):
        if side is not None:
            dw.zero_()
        with _lib.timed("wgrad_halo_tcgen05" if use_halo else "wgrad_tcgen05", 2 * V * (cin + cout) + 4 * cin * cout * taps,
                        2 * V * cin * cout * taps):
            check(fn(ptr(dy_tok), ll(dy_tok.stride(0)), ptr(x_tok), ll(x_tok.stride(0)), batch, D, H, W, cin, cout, ks[0], ks[1],
                     ks[2], ptr(dw), cs, cstream()), "nextou_conv3d_ndhwc_wgrad")
    finish = lambda: dw[:, :, :cin].permute(0, 2, 1).reshape(cout, cin, *ksize)
    return finish if side is not None else finish()


# ----------------------------------------------------------------------------------------------
# batch / instance norm (+ LeakyReLU) on token-major matrices (TR:54-55, TN:32-51)
# The kernels stream the physical [rows, pitch] matrix; padding channels (pitch > C) are extra don't-care lanes.
# ----------------------------------------------------------------------------------------------
def _physical_rows(t: torch.Tensor):
    """[rows, C] token view -> (physical [rows, pitch] matrix, C).  Copies only if the rows are not addressable."""
    t = _tok2d(_work_dtype(t))
    full = full_rows(t)
    if full is None:
        t = t.contiguous()
        full = t
    return full, t.shape[1]


def _rows_like(t: torch.Tensor, rows: int, C: int, pitch: int, dtype) -> torch.Tensor:
    """Physical [rows, pitch] matrix whose first C columns equal `t` (a view when t already has that layout)."""
    t = _tok2d(t)
    if t.dtype == dtype and t.stride(0) == pitch:
        full = full_rows(t)
        if full is not None:
            return full
    buf = torch.empty((rows, pitch), device=t.device, dtype=dtype)
    buf[:, :C].copy_(t)
    return buf


def _pad_vec(v: Optional[torch.Tensor], n: int, value: float):
    if v is None:
        return None
    v = v.detach().float().contiguous()
    return v if v.numel() == n else torch.nn.functional.pad(v, (0, n - v.numel()), value=value)


def _norm_partial(C, rows, instances, device):
    nblk = ctypes.c_int(0)
    check(_lib.lib().nextou_norm_plan(C, ll(rows), instances, ctypes.byref(nblk)), "nextou_norm_plan")
    return torch.empty(instances * nblk.value * 2 * C, device=device, dtype=torch.float32)


class _DxColsum:
    """Hand-over of the per-channel column sums that the normalisation backward computes while it writes dx, to the
    backward of the producing convolution / linear layer, whose bias gradient they are.  Holds ONE entry (the most recent
    normalisation backward) together with a reference to dx, so the address cannot be recycled while the entry is live."""
    entry = None

    @classmethod
    def put(cls, dx_full, sums):
        cls.entry = (dx_full, sums)

    @classmethod
    def take(cls, tok):
        e, cls.entry = cls.entry, None
        if e is None:
            return None
        dx, sums = e
        if (tok.data_ptr() == dx.data_ptr() and tok.dtype == dx.dtype and tok.dim() == 2 and tok.shape[0] == dx.shape[0]
                and tok.stride(0) == dx.stride(0) and tok.stride(1) == 1 and tok.shape[1] <= dx.shape[1]):
            s = sums[0] if sums.shape[0] == 1 else sums.sum(0)
            return s[:tok.shape[1]]
        return None


class _NormAct(torch.autograd.Function):
    """Train-mode normalisation with batch statistics + optional LeakyReLU (slope 1.0 = none)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, slope, instances, tracked=None, residual=None):
        xf, C = _physical_rows(x)
        _need_cuda(xf)
        T, P = xf.shape                      # P = physical row pitch >= C
        rows = T // instances
        assert rows * instances == T
        L = _lib.lib()
        f32 = lambda v: None if v is None else v.detach().float().contiguous()   # no copy for fp32 parameters / buffers
        g32, b32 = f32(gamma), f32(beta)
        rm = rv = None
        if running_mean is not None:
            assert running_mean.dtype == torch.float32 and running_var.dtype == torch.float32
            rm, rv = running_mean, running_var
        partial = _norm_partial(P, rows, instances, xf.device)
        mean = torch.empty(instances * P, device=xf.device, dtype=torch.float32)
        invstd = torch.empty_like(mean)
        check(L.nextou_norm_stats_tracked(ptr(xf), dtype_code(xf), P, C, ll(rows), instances, cf(eps), ptr(partial), ptr(mean),
                                          ptr(invstd), ptr(rm), ptr(rv), cf(momentum if momentum is not None else 0.0),
                                          ptr(tracked), cstream()), "nextou_norm_stats_tracked")
        y = torch.empty_like(xf)
        rf = None
        if residual is not None:      # shortcut of the residual block, in the same physical layout as x
            rf = full_rows(_tok2d(residual))
            if rf is None or rf.shape != xf.shape or rf.dtype != xf.dtype:
                rf = _rows_like(residual, T, C, P, xf.dtype)
        check(L.nextou_norm_apply_res(ptr(xf), dtype_code(xf), P, C, ll(rows), instances, ptr(mean), ptr(invstd), ptr(g32),
                                      ptr(b32), cf(slope), ptr(rf), ptr(y), cstream()), "nextou_norm_apply")
        ctx.save_for_backward(xf, mean, invstd, g32, b32)
        ctx.meta = (C, P, rows, instances, slope, gamma is not None, None if gamma is None else gamma.dtype)
        ctx.has_residual = residual is not None
        return y[:, :C]

    @staticmethod
    def backward(ctx, dy):
        xf, mean, invstd, g32, b32 = ctx.saved_tensors
        C, P, rows, instances, slope, affine, pdt = ctx.meta
        dyf = _rows_like(dy, xf.shape[0], C, P, xf.dtype)
        partial = _norm_partial(P, rows, instances, xf.device)
        sums = torch.empty(instances * 2 * P, device=xf.device, dtype=torch.float32)
        dx = torch.empty_like(xf)
        dxsum = torch.empty((instances, P), device=xf.device, dtype=torch.float32)
        check(_lib.lib().nextou_norm_bwd_colsum(ptr(xf), ptr(dyf), dtype_code(xf), P, C, ll(rows), instances, ptr(mean),
                                                ptr(invstd), ptr(g32), ptr(b32), cf(slope), ptr(partial), ptr(sums), ptr(dx),
                                                ptr(dxsum), cstream()), "nextou_norm_bwd_colsum")
        # the column sums of dx are the bias gradient of the layer that produced x: colsum_tokens() picks them up
        _DxColsum.put(dx, dxsum)
        dgamma = dbeta = None
        if affine:
            s = sums.view(instances, 2, P)
            s = s[0] if instances == 1 else s.sum(0)
            dbeta, dgamma = s[0, :C].to(pdt), s[1, :C].to(pdt)
        return dx[:, :C], dgamma, dbeta, None, None, None, None, None, None, None, (dy if ctx.has_residual else None)


def sync_moments(sums: torch.Tensor, count: int, eps: float, group=None):
    """Cross-rank batch statistics of a SyncBatchNorm layer (upstream nnU-Net converts every BatchNorm under DDP, SURVEY.md
    §8e).  sums: fp32 [2, P] = this rank's (sum x, sum x^2) per channel; count: this rank's rows.  One all-reduce of
    [2P + 1] values; returns (mean [P], invstd [P], unbiased variance [P], total row count as a python int is NOT needed on
    the host: the count travels in the same buffer and stays on the device) -> (mean, invstd, var_unbiased, n_total tensor)."""
    import torch.distributed as dist
    P = sums.shape[1]
    buf = torch.cat([sums.reshape(-1).double(), torch.full((1,), float(count), device=sums.device, dtype=torch.float64)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n = buf[2 * P]
    mean = buf[:P] / n
    var = (buf[P:2 * P] / n - mean * mean).clamp_min(0.0)
    invstd = torch.rsqrt(var + eps)
    unbiased = var * (n / (n - 1.0).clamp_min(1.0))
    return mean.float(), invstd.float(), unbiased.float(), n


_EQUAL_ROWS = {}


def _equal_rows(rows: int, group, world: int) -> bool:
    """True when every rank of `group` normalises the same number of rows (nnU-Net DDP splits the batch unevenly when the
    batch size is not divisible by the world size).  One tiny all-gather per (group, rows), remembered afterwards."""
    import torch.distributed as dist
    key = (id(group), rows, world)
    hit = _EQUAL_ROWS.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            raise NextouError("SyncBatchNorm: run one eager step before capturing (row counts are compared across ranks once)")
        mine = torch.tensor([rows], device="cuda", dtype=torch.int64)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        hit = _EQUAL_ROWS[key] = all(int(t) == rows for t in every)
    return hit


# SyncBatchNorm statistics over NVLink peer memory (csrc/syncnorm.cu); False / NEXTOU_PEER_EXCHANGE=0: NCCL all-reduce (A/B)
PEER_EXCHANGE = os.environ.get("NEXTOU_PEER_EXCHANGE", "1") != "0"


def _peer_exchange(group, device):
    if not PEER_EXCHANGE:
        return None
    from .parallel import PeerExchange
    return PeerExchange.get(group, device)


class _SyncNormAct(torch.autograd.Function):
    """Train-mode SyncBatchNorm (+ LeakyReLU): statistics over the rows of ALL ranks.  Forward: local per-CTA partial sums ->
    one exchange kernel over NVLink peer memory (sum over ranks, mean / invstd / running statistics) -> normalise; backward:
    local partials of (sum dy', sum dy' xhat) -> one exchange kernel -> dx with the global sums and row count
    (torch.nn.SyncBatchNorm semantics: d gamma / d beta stay local, the gradient all-reduce averages them).  Without peer
    memory the sums travel through NCCL all-reduces instead."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, slope, tracked, group, world, residual=None):
        xf, C = _physical_rows(x)
        _need_cuda(xf)
        T, P = xf.shape
        L = _lib.lib()
        f32 = lambda v: None if v is None else v.detach().float().contiguous()
        g32, b32 = f32(gamma), f32(beta)
        partial = _norm_partial(P, T, 1, xf.device)
        px = _peer_exchange(group, xf.device)
        key = (gamma.data_ptr() if gamma is not None else id(ctx), P)
        if px is not None:
            nblk = ctypes.c_int(0)
            check(L.nextou_norm_partial_stats(ptr(xf), dtype_code(xf), P, ll(T), ptr(partial), ctypes.byref(nblk), cstream()),
                  "nextou_norm_partial_stats")
            off, ep = px.slot((key, "fwd"), P, False)
            mean = torch.empty(P, device=xf.device, dtype=torch.float32)
            invstd = torch.empty_like(mean)
            n = torch.empty((), device=xf.device, dtype=torch.float64)
            track = running_mean is not None
            check(L.nextou_sync_norm_finalize(ptr(partial), nblk.value, P, C, ll(T), cf(eps), ptr(px.peer_base), ll(off), px.rank,
                                              px.world, ptr(ep), ptr(mean), ptr(invstd), ptr(running_mean if track else None),
                                              ptr(running_var if track else None), cf(momentum), ptr(tracked if track else None),
                                              ptr(n), cstream()), "nextou_sync_norm_finalize")
        else:
            sums = torch.empty(2 * P, device=xf.device, dtype=torch.float32)
            check(L.nextou_colsum(ptr(xf), dtype_code(xf), P, ll(T), ptr(partial), ptr(sums), cstream()), "nextou_colsum")
            mean, invstd, unbiased, n = sync_moments(sums.view(2, P), T, eps, group)
            if running_mean is not None:
                running_mean.mul_(1.0 - momentum).add_(mean[:C], alpha=momentum)
                running_var.mul_(1.0 - momentum).add_(unbiased[:C], alpha=momentum)
                if tracked is not None:
                    tracked.add_(1)
        y = torch.empty_like(xf)
        rf = None
        if residual is not None:      # shortcut of the residual block, in the same physical layout as x
            rf = full_rows(_tok2d(residual))
            if rf is None or rf.shape != xf.shape or rf.dtype != xf.dtype:
                rf = _rows_like(residual, T, C, P, xf.dtype)
        check(L.nextou_norm_apply_res(ptr(xf), dtype_code(xf), P, C, ll(T), 1, ptr(mean), ptr(invstd), ptr(g32), ptr(b32),
                                      cf(slope), ptr(rf), ptr(y), cstream()), "nextou_norm_apply")
        ctx.has_residual = residual is not None
        # backward divides the all-reduced sums by the number of rows normalised together.  Equal row counts on every rank
        # (one patch each: checked ONCE per (group, rows) on the host) -> T * world; else the true count from the exchange
        fix = None if _equal_rows(T, group, world) else ((T * world) / n).float().reshape(1)
        ctx.save_for_backward(xf, mean, invstd, g32, b32, fix)
        ctx.meta = (C, P, T, slope, gamma is not None, None if gamma is None else gamma.dtype, group, world, key)
        return y[:, :C]

    @staticmethod
    def backward(ctx, dy):
        import torch.distributed as dist
        xf, mean, invstd, g32, b32, fix = ctx.saved_tensors
        C, P, T, slope, affine, pdt, group, world, key = ctx.meta
        L = _lib.lib()
        dyf = _rows_like(dy, T, C, P, xf.dtype)
        partial = _norm_partial(P, T, 1, xf.device)
        px = _peer_exchange(group, xf.device)
        if px is not None:
            nblk = ctypes.c_int(0)
            check(L.nextou_norm_bwd_partial(ptr(xf), ptr(dyf), dtype_code(xf), P, C, ll(T), ptr(mean), ptr(invstd), ptr(g32), ptr(b32),
                                            cf(slope), ptr(partial), ctypes.byref(nblk), cstream()), "nextou_norm_bwd_partial")
            off, ep = px.slot((key, "bwd"), P, True)
            local = torch.empty(2 * P, device=xf.device, dtype=torch.float32)
            sums = torch.empty(2 * P, device=xf.device, dtype=torch.float32)
            check(L.nextou_sync_norm_bwd_finalize(ptr(partial), nblk.value, P, ptr(px.peer_base), ll(off), px.rank, px.world, ptr(ep),
                                                  ptr(local), ptr(sums), cstream()), "nextou_sync_norm_bwd_finalize")
            local = local.view(2, P)
            dgamma = dbeta = None
            if affine:
                dbeta, dgamma = local[0, :C].to(pdt), local[1, :C].to(pdt)
        else:
            sums = torch.empty(2 * P, device=xf.device, dtype=torch.float32)
            check(L.nextou_norm_bwd_reduce(ptr(xf), ptr(dyf), dtype_code(xf), P, C, ll(T), 1, ptr(mean), ptr(invstd), ptr(g32),
                                           ptr(b32), cf(slope), ptr(partial), ptr(sums), cstream()), "nextou_norm_bwd_reduce")
            local = sums.view(2, P)
            dgamma = dbeta = None
            if affine:
                dbeta, dgamma = local[0, :C].clone().to(pdt), local[1, :C].clone().to(pdt)   # `sums` is reduced in place below
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        if fix is not None:
            sums.mul_(fix)      # uneven batch split: sums / n_true == (sums * T * world / n_true) / (T * world)
        dx = torch.empty_like(xf)
        dxsum = torch.empty((1, P), device=xf.device, dtype=torch.float32)
        check(L.nextou_norm_bwd_apply(ptr(xf), ptr(dyf), dtype_code(xf), P, C, ll(T), 1, ll(T * world), ptr(mean), ptr(invstd),
                                      ptr(g32), ptr(b32), cf(slope), ptr(sums), ptr(partial), ptr(dx), ptr(dxsum), cstream()),
              "nextou_norm_bwd_apply")
        _DxColsum.put(dx, dxsum)
        return dx[:, :C], dgamma, dbeta, None, None, None, None, None, None, None, None, (dy if ctx.has_residual else None)


def sync_norm_act_tokens(x_tok, gamma, beta, running_mean, running_var, momentum, eps, slope, num_batches_tracked, group,
                         world: int, residual=None):
    """SyncBatchNorm (+ LeakyReLU) (+ residual shortcut) over the rows of every rank of `group`."""
    return _SyncNormAct.apply(x_tok, gamma, beta, running_mean, running_var, float(momentum), float(eps), float(slope),
                              num_batches_tracked, group, int(world), residual)


def norm_act_tokens(x_tok, gamma, beta, running_mean=None, running_var=None, momentum=0.1, eps=1e-5, slope=1.0,
                    instances=1, num_batches_tracked=None, residual=None):
    """Batch norm (instances=1) / instance norm (instances=batch) with batch statistics, + LeakyReLU(slope).
    num_batches_tracked (int64 0-d CUDA tensor) is incremented by the statistics kernel."""
    if num_batches_tracked is not None:
        assert num_batches_tracked.dtype == torch.int64 and num_batches_tracked.is_cuda
    return _NormAct.apply(x_tok, gamma, beta, running_mean, running_var, momentum, eps, float(slope), int(instances),
                          num_batches_tracked, residual)


def affine_act_tokens(x_tok, scale, shift, slope=1.0):
    """Eval-mode batch norm: y = lrelu(x * scale[c] + shift[c]) (no autograd: inference only)."""
    if torch.is_grad_enabled() and x_tok.requires_grad:
        raise NextouError("affine_act_tokens (eval-mode norm) does not implement a backward pass")
    xf, C = _physical_rows(x_tok)
    _need_cuda(xf)
    P = xf.shape[1]
    y = torch.empty_like(xf)
    # (named: a temporary passed straight to ptr() is freed before the launch and its block re-used by the next temporary)
    sc, sh = _pad_vec(scale, P, 0.0), _pad_vec(shift, P, 0.0)
    check(_lib.lib().nextou_affine_act(ptr(xf), dtype_code(xf), P, ll(xf.shape[0]), ptr(sc), ptr(sh), cf(slope), ptr(y), cstream()),
          "nextou_affine_act")
    return y[:, :C]


# ----------------------------------------------------------------------------------------------
# residual add / channel concat that keep the channel-padded token layout (ED:322, 389, 817, 932)
# ----------------------------------------------------------------------------------------------
class _AddTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        fa, fb = full_rows(_tok2d(a)), full_rows(_tok2d(b))
        if fa is not None and fb is not None and fa.shape == fb.shape and fa.dtype == fb.dtype:
            return (fa + fb)[:, :a.shape[1]]                   # one pass over the physical rows, layout preserved
        out = padded_like(a.shape[0], a.shape[1], torch.promote_types(a.dtype, b.dtype), a.device)
        torch.add(a, b, out=out)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add_tokens(a, b):
    return _AddTokens.apply(a, b)


def _rows_view_ok(t: torch.Tensor, cols: int, per: int) -> bool:
    """True if the [rows, cols] rectangle starting at t (cols >= t.shape[1], the channel padding included) can be moved with
    16-byte vectors: unit inner stride, pitch / base aligned, rectangle inside the row pitch and inside the storage."""
    if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % per or t.stride(0) < cols or t.data_ptr() % 16:
        return False
    need = t.storage_offset() + (t.shape[0] - 1) * t.stride(0) + cols
    return need <= t.untyped_storage().nbytes() // t.element_size()


def rows_copy_add(a: torch.Tensor, b: Optional[torch.Tensor], out: torch.Tensor) -> bool:
    """out[:, :C] = a (+ b) for [rows, C] token views of different pitch / column offset through csrc/pool.cu's vector kernel;
    the padding lanes up to the next multiple of 16 bytes are moved along (don't-care).  Returns False (nothing done) when a
    view is not vector-addressable: the caller falls back to the strided ATen op."""
    if a.dtype != out.dtype or (b is not None and b.dtype != a.dtype) or a.dtype not in (torch.bfloat16, torch.float32):
        return False
    per = 8 if a.dtype == torch.bfloat16 else 4
    C = a.shape[1]
    cols = (C + per - 1) // per * per
    if out.shape != a.shape or (b is not None and b.shape != a.shape) or a.shape[0] == 0:
        return False
    if not all(_rows_view_ok(t, cols, per) for t in (a, out) + ((b,) if b is not None else ())):
        return False
    _need_cuda(a, b, out)
    check(_lib.lib().nextou_rows_copy_add(ptr(a), ll(a.stride(0)), ptr(b), ll(0 if b is None else b.stride(0)), ptr(out),
                                          ll(out.stride(0)), ll(a.shape[0]), cols, dtype_code(a), cstream()), "nextou_rows_copy_add")
    return True


def _sum_grads(g1: torch.Tensor, g2: torch.Tensor) -> torch.Tensor:
    """g1 + g2 for two [rows, C] token gradients, keeping the channel-padded layout: one vectorised pass over the physical
    rows when both share it, else one strided add INTO a padded buffer (autograd's own accumulation would produce an
    unpadded tensor that every TMA consumer has to copy again)."""
    a, b = _tok2d(g1), _tok2d(g2)
    if a.dtype != b.dtype:
        b = b.to(a.dtype)
    fa, fb = full_rows(a), full_rows(b)
    if fa is not None and fb is not None and fa.shape == fb.shape and fa.shape[1] % 8 == 0:
        return (fa + fb)[:, :a.shape[1]]
    out = padded_like(a.shape[0], a.shape[1], a.dtype, a.device)
    if not rows_copy_add(a, b, out):          # e.g. the skip half of a concatenation gradient (pitch 80) + a pitch-40 gradient
        torch.add(a, b, out=out)
    return out


class _ForkTokens(torch.autograd.Function):
    """Explicit fan-out of a token tensor that is consumed twice (residual shortcut + branch, U-Net skip + next stage):
    the two incoming gradients are summed HERE, in the padded token layout, instead of by autograd's accumulation."""

    @staticmethod
    def forward(ctx, tok):
        ctx.set_materialize_grads(False)
        return tok.view_as(tok), tok.view_as(tok)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None:
            return g2
        if g2 is None:
            return g1
        return _sum_grads(g1, g2)


def fork_tokens(tok: torch.Tensor):
    """-> two aliases of `tok` whose gradients are summed by nextou_b200 (see _ForkTokens); a no-op without autograd."""
    if tok.requires_grad and torch.is_grad_enabled():
        return _ForkTokens.apply(tok)
    return tok, tok


def fork(x: torch.Tensor):
    """fork_tokens for a logical (N, C, *spatial) activation."""
    if not (x.requires_grad and torch.is_grad_enabled()):
        return x, x
    B, spatial = x.shape[0], tuple(x.shape[2:])
    a, b = fork_tokens(as_tokens(x))
    return from_tokens(a, B, spatial), from_tokens(b, B, spatial)


class _CatTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ca, cb = a.shape[1], b.shape[1]
        out = padded_like(a.shape[0], ca + cb, a.dtype, a.device)
        out[:, :ca].copy_(a)
        if not rows_copy_add(b, None, out[:, ca:]):
            out[:, ca:].copy_(b)
        ctx.split = ca
        return out

    @staticmethod
    def backward(ctx, g):
        return g[:, :ctx.split], g[:, ctx.split:]


def cat_tokens(a, b):
    """[rows, Ca], [rows, Cb] -> [rows, Ca + Cb] (torch.cat((up, skip), 1) of the decoder, ED:322), padded layout."""
    return _CatTokens.apply(a, b.to(a.dtype))


def colsum_tokens(tok: torch.Tensor) -> torch.Tensor:
    """fp32 [C] column sums of a [rows, C] token view (bias gradients): taken from the normalisation backward that just
    wrote `tok` when there is one, else one streaming pass (csrc/norm.cu)."""
    hit = _DxColsum.take(tok)
    if hit is not None:
        return hit
    xf, C = _physical_rows(tok)
    _need_cuda(xf)
    T, P = xf.shape
    partial = _norm_partial(P, T, 1, xf.device)
    sums = torch.empty(2 * P, device=xf.device, dtype=torch.float32)
    check(_lib.lib().nextou_colsum(ptr(xf), dtype_code(xf), P, ll(T), ptr(partial), ptr(sums), cstream()), "nextou_colsum")
    return sums[:C]
