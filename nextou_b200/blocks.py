"""NexToU building blocks — API mirror of the reference's network_architecture/NexToU_Encoder_Decoder.py (ED).

Class names, constructor signatures, attribute names (hence state_dict keys) and forward semantics follow the
reference; the execution does not.  Per-token operators (1x1 conv, BatchNorm statistics, residuals) are
invariant to the order of the tokens, so a Swin grapher never rolls, partitions or reverses the volume
(ED:634-693, 784, 813): the shifted-window structure only exists as an int32 row map consumed by the kNN and
message-passing kernels.  The N x M distance matrix and the (B, C, N, k) gathered tensors of the reference are
never materialised (csrc/knn.cu, csrc/mrconv.cu); pooling / unpooling run on the channels-last volume
(csrc/pool.cu).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Type, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import dense, graph, ops
from .conv_blocks import (StackedConvBlocks, convert_conv_op_to_dim, get_matching_convtransp, get_matching_pool_op,
                          maybe_convert_scalar_to_list)
from .layers import BasicConv, act_layer
from .pos_embed import relative_pos_parameter


def _ndim(conv_op) -> int:
    if conv_op == nn.Conv2d:
        return 2
    if conv_op == nn.Conv3d:
        return 3
    raise NotImplementedError("conv operation [%s] is not found" % conv_op)


def _prod(v) -> int:
    out = 1
    for a in v:
        out *= int(a)
    return out


class OptInit:
    """Hard-wired GNN hyper-parameters (ED:17-32)."""

    def __init__(self, drop_path_rate=0., pool_op_kernel_sizes_len=4):
        self.pool_op_kernel_sizes_len = pool_op_kernel_sizes_len
        self.conv = 'mr'
        self.act = 'leakyrelu'
        self.norm = 'instance'
        self.bias = True
        self.dropout = 0.0
        self.use_dilation = True
        self.epsilon = 0.2
        self.use_stochastic = True
        self.drop_path = drop_path_rate
        self.blocks = [1] * pool_op_kernel_sizes_len
        self.reduce_ratios = [16, 8, 4, 2] + [1] * (pool_op_kernel_sizes_len - 4)


class DropPath(nn.Module):
    """Stochastic depth (timm's DropPath; never instantiated by NexToU because drop_path is 0, ED:382/721/843)."""

    def __init__(self, drop_prob: float = 0.):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _residual(mod, out_tok, short_tok, B, spatial):
    """drop_path(out) + x (ED:389, 817, 932).  drop_path is the identity for NexToU (rate 0): then the add runs on the
    physical token rows and keeps the channel-padded layout; a real DropPath goes through the logical view.  `short_tok` is
    the shortcut alias of the block input from ops.fork_tokens (the branch consumed the other one)."""
    if isinstance(mod.drop_path, nn.Identity):
        return ops.from_tokens(ops.add_tokens(out_tok, short_tok), B, spatial)
    return mod.drop_path(ops.from_tokens(out_tok, B, spatial)) + ops.from_tokens(short_tok, B, spatial)


def _fc_bn(seq: nn.Sequential, tok: torch.Tensor, batch: int, act_slope=None, residual=None) -> torch.Tensor:
    """nn.Sequential(1x1 conv, norm) as used for fc1 / fc2 everywhere (ED:373-381, 710-720, 833-842), on token rows:
    a GEMM followed by the fused norm (+ LeakyReLU) (+ residual shortcut) kernel."""
    return dense.linear_norm_act_tokens(tok, seq[0], seq[1], batch, act_slope, residual)


def _fc_bn_residual(mod, seq, tok, short_tok, B, spatial):
    """x + drop_path(BN(fc2(.)))  (ED:389, 817, 932): with drop_path = identity (NexToU: rate 0) the shortcut add is fused
    into the norm-apply kernel; a real DropPath goes through the logical view."""
    if isinstance(mod.drop_path, nn.Identity):
        return ops.from_tokens(_fc_bn(seq, tok, B, residual=short_tok), B, spatial)
    return _residual(mod, _fc_bn(seq, tok, B), short_tok, B, spatial)


class FFN(nn.Module):
    """x + BN(fc2(act(BN(fc1 x))))  (ED:368-390)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act='relu', drop_path=0.0,
                 conv_op=nn.Conv3d, norm_op=nn.BatchNorm3d, norm_op_kwargs=None):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        norm_op_kwargs = norm_op_kwargs or {}
        self.fc1 = nn.Sequential(conv_op(in_features, hidden_features, 1, stride=1, padding=0),
                                 norm_op(hidden_features, **norm_op_kwargs))
        self.act = act_layer(act)
        self.fc2 = nn.Sequential(conv_op(hidden_features, out_features, 1, stride=1, padding=0),
                                 norm_op(out_features, **norm_op_kwargs))
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward(self, x):
        B, spatial = x.shape[0], tuple(x.shape[2:])
        tok, short = ops.fork_tokens(ops.as_tokens(x))
        if isinstance(self.act, nn.LeakyReLU):
            h = _fc_bn(self.fc1, tok, B, self.act.negative_slope)
        else:
            h = self.act(_fc_bn(self.fc1, tok, B))
        return _fc_bn_residual(self, self.fc2, h, short, B, spatial)


# ------------------------------------------------------------------------------------------------------
# graph convolutions
# ------------------------------------------------------------------------------------------------------
class MRConv(nn.Module):
    """Max-relative graph convolution on (B, C, N, 1) tensors with an explicit edge_index (ED:392-418)."""

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True, conv_op=nn.Conv3d,
                 dropout_op=nn.Dropout3d):
        super().__init__()
        self.conv_op = conv_op
        self.nn = BasicConv([in_channels * 2, out_channels], act=act, norm=norm, bias=bias, drop=0., conv_op=conv_op,
                            dropout_op=dropout_op)

    def forward(self, x, edge_index, y=None):
        ndim = _ndim(self.conv_op)
        B, C, N, _ = x.shape
        idx32 = edge_index[0].to(torch.int32)
        xt = x.squeeze(-1).transpose(1, 2).reshape(B * N, C)
        if y is None:
            feat = ops.mrconv_gather(xt, idx32, N, N)
        else:
            M = y.shape[2]
            feat = ops.mrconv_gather(xt, idx32, N, M, y_tok=y.squeeze(-1).transpose(1, 2).reshape(B * M, C))
        feat = feat.reshape(B, N, 2 * C).transpose(1, 2).unsqueeze(-1)
        if ndim == 3:
            feat = feat.unsqueeze(4)
        return self.nn(feat)


class GraphConv(nn.Module):
    """Static graph convolution wrapper; only conv='mr' exists (ED:420-432)."""

    def __init__(self, in_channels, out_channels, conv='edge', act='relu', norm=None, bias=True, conv_op=nn.Conv3d,
                 dropout_op=nn.Dropout3d):
        super().__init__()
        if conv == 'mr':
            self.gconv = MRConv(in_channels, out_channels, act, norm, bias, conv_op, dropout_op)
        else:
            raise NotImplementedError('conv:{} is not supported'.format(conv))

    def forward(self, x, edge_index, y=None):
        return self.gconv(x, edge_index, y)


def _dyn_graph_features(mod, q_tok, graphs, n, y_tok, m, relative_pos, q_row_map=None):
    """kNN (no grad) + message passing on token-major rows; returns [rows, 2C] in the row order of q_tok."""
    k, d = mod.k, mod.d
    knn_mod = mod.dilated_knn_graph
    cols = graph.draw_stochastic_columns(k, d, knn_mod.stochastic, knn_mod.epsilon, mod.training)
    forced = getattr(mod, "forced_nn_idx", None)
    if forced is not None:
        # diagnostics only (tools/dist_checks.py): neighbour lists recorded from another run of the same network, so that two
        # runs whose statistics differ in the last bit do not diverge through a flipped near-tie (chaos note)
        idx32 = forced
    elif cols is not None and d > 1:
        _, idx32 = ops.knn_graph(q_tok, graphs, n, y_tok, m, relpos=relative_pos, k=k * d, dilation=1,
                                 x_row_map=q_row_map, y_row_map=q_row_map if y_tok is None else None)
        idx32 = idx32[:, :, cols.to(idx32.device)].contiguous()
    else:
        # with dilation 1 the random branch only permutes the k neighbours: max aggregation is invariant to it
        _, idx32 = ops.knn_graph(q_tok, graphs, n, y_tok, m, relpos=relative_pos, k=k, dilation=d,
                                 x_row_map=q_row_map, y_row_map=q_row_map if y_tok is None else None)
    mod.last_nn_idx = idx32  # (graphs, n, k) int32, kept for tests / inspection
    return ops.mrconv_gather(q_tok, idx32, n, m if y_tok is not None else n, y_tok=y_tok, q_row_map=q_row_map,
                             y_row_map=q_row_map if y_tok is None else None)


class DyGraphConv(GraphConv):
    """Dynamic graph convolution: kNN on the current features, then MRConv (ED:434-474)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv='edge', act='relu', norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, conv_op=nn.Conv3d, dropout_op=nn.Dropout3d):
        super().__init__(in_channels, out_channels, conv, act, norm, bias, conv_op, dropout_op)
        self.k = kernel_size
        self.d = dilation
        self.r = r
        self.dilated_knn_graph = graph.DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)
        self.conv_op = conv_op
        self.dropout_op = dropout_op
        self.ndim = _ndim(conv_op)
        self.avg_pool = F.avg_pool2d if self.ndim == 2 else F.avg_pool3d

    def forward_tokens(self, tok, B, spatial, relative_pos=None, row_map=None, graphs=None, n=None):
        """[B*prod(spatial), C] -> [.., 2C] token rows.  With `row_map` the rows are regrouped into `graphs` graphs of
        `n` tokens each by pure indexing (shifted windows)."""
        if graphs is None:
            graphs, n = B, _prod(spatial)
        y_tok, m = None, n
        if self.r > 1:
            if row_map is not None:
                raise NotImplementedError("reduce ratio r > 1 inside windows (never built by NexToU, ED:1003)")
            y_tok = ops.avgpool_tokens(tok, B, spatial, (self.r,) * self.ndim)
            m = y_tok.shape[0] // B
        feat = _dyn_graph_features(self, tok, graphs, n, y_tok, m, relative_pos, q_row_map=row_map)
        return self.gconv.nn.forward_tokens(feat, B)

    def forward(self, x, relative_pos=None):
        B, spatial = x.shape[0], tuple(x.shape[2:])
        return ops.from_tokens(self.forward_tokens(ops.as_tokens(x), B, spatial, relative_pos), B, spatial)


def _pool_size_for(img_shape, img_min_shape):
    """[2,..] if the stage has more tokens than prod(4 * min_shape), else [1,..] (ED:490-503, 845-858)."""
    n = _prod(img_shape)
    n_small = _prod([h * 4 for h in img_min_shape])
    if n > n_small:
        return [2 if h % 2 == 0 else 1 for h in img_shape]
    return [1 for _ in img_shape]


class PoolDyGraphConv(GraphConv):
    """max-pool the queries, avg-pool the candidates, kNN + MRConv, max-unpool (ED:476-551)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv='edge', act='relu', norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, conv_op=nn.Conv3d, dropout_op=nn.Dropout3d,
                 img_shape=None, img_min_shape=None):
        super().__init__(in_channels, out_channels, conv, act, norm, bias, conv_op, dropout_op)
        self.k = kernel_size
        self.d = dilation
        self.r = r
        self.dilated_knn_graph = graph.DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)
        self.conv_op = conv_op
        self.dropout_op = dropout_op
        self.ndim = _ndim(conv_op)
        self.pool_size = _pool_size_for(img_shape, img_min_shape)
        pool_cls = get_matching_pool_op(conv_op, pool_type='max')
        unpool_cls = nn.MaxUnpool2d if self.ndim == 2 else nn.MaxUnpool3d
        self.avg_pool = F.avg_pool2d if self.ndim == 2 else F.avg_pool3d
        # parameter-free; kept so the module tree prints like the reference.  forward() uses csrc/pool.cu
        self.max_pool_input = pool_cls(self.pool_size, stride=self.pool_size, return_indices=True)
        self.max_unpool_output = unpool_cls(self.pool_size, stride=self.pool_size)

    def forward_tokens(self, tok, B, spatial, relative_pos=None):
        """[B*prod(spatial), C] -> [B*prod(spatial), 2C] token rows (un-pooled back to full resolution)."""
        pooled = any(p > 1 for p in self.pool_size)
        if pooled:
            q_tok, arg = ops.maxpool_tokens(tok, B, spatial, self.pool_size)
            q_spatial = tuple(s // p for s, p in zip(spatial, self.pool_size))
        else:
            q_tok, arg, q_spatial = tok, None, spatial
        n = _prod(q_spatial)
        y_tok, m = None, n
        if self.r > 1:
            y_tok = ops.avgpool_tokens(q_tok, B, q_spatial, (self.r,) * self.ndim)
            m = y_tok.shape[0] // B
        feat = _dyn_graph_features(self, q_tok, B, n, y_tok, m, relative_pos)
        g = self.gconv.nn.forward_tokens(feat, B)
        if pooled:
            g = ops.maxunpool_tokens(g, arg, B, spatial, self.pool_size)
        return g

    def forward(self, x, relative_pos=None):
        B, spatial = x.shape[0], tuple(x.shape[2:])
        return ops.from_tokens(self.forward_tokens(ops.as_tokens(x), B, spatial, relative_pos), B, spatial)


def _make_relative_pos(in_channels, n, r, ndim):
    return relative_pos_parameter(in_channels, n, n // (r ** ndim), ndim)


def _resized_relative_pos(relative_pos, n_now, n_built, r, ndim):
    """Bicubic resize when the token count differs from construction time (ED:744-763, 882-901)."""
    if relative_pos is None or n_now == n_built:
        return relative_pos
    return F.interpolate(relative_pos.unsqueeze(0), size=(n_now, n_now // (r ** ndim)), mode="bicubic").squeeze(0)


class Grapher(nn.Module):
    """Plain ViG grapher (fc1 -> graph conv -> fc2 + residual).  Unused by NexToU, kept for API parity (ED:553-632)."""

    def __init__(self, in_channels, kernel_size=9, dilation=1, conv='edge', act='relu', norm=None, bias=True,
                 stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0, relative_pos=False, conv_op=nn.Conv3d,
                 norm_op=nn.BatchNorm3d, dropout_op=nn.Dropout3d):
        super().__init__()
        self.channels = in_channels
        self.n = n
        self.r = r
        self.conv_op = conv_op
        self.ndim = _ndim(conv_op)
        self.fc1 = nn.Sequential(conv_op(in_channels, in_channels, 1, stride=1, padding=0), norm_op(in_channels))
        self.graph_conv = DyGraphConv(in_channels, in_channels * 2, kernel_size, dilation, conv, act, norm, bias,
                                      stochastic, epsilon, r, conv_op, dropout_op)
        self.fc2 = nn.Sequential(conv_op(in_channels * 2, in_channels, 1, stride=1, padding=0), norm_op(in_channels))
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.relative_pos = _make_relative_pos(in_channels, n, r, self.ndim) if relative_pos else None

    def forward(self, x):
        B, spatial = x.shape[0], tuple(x.shape[2:])
        tok, short = ops.fork_tokens(ops.as_tokens(x))
        h = _fc_bn(self.fc1, tok, B)
        rp = _resized_relative_pos(self.relative_pos, _prod(spatial), self.n, self.r, self.ndim)
        return _fc_bn_residual(self, self.fc2, self.graph_conv.forward_tokens(h, B, spatial, rp), short, B, spatial)


# ------------------------------------------------------------------------------------------------------
# windows
# ------------------------------------------------------------------------------------------------------
def window_partition(x, window_size):
    """(B, C, *S) -> (B * nW, C, *window)  (ED:634-660).  API helper (a strided copy); the hot path uses row maps."""
    if x.dim() not in (4, 5):
        raise NotImplementedError('len(x.shape) [%d] is equal to 4 or 5' % x.dim())
    B, C = x.shape[:2]
    S = x.shape[2:]
    d = len(S)
    v = x.reshape(B, C, *[t for s, w in zip(S, window_size) for t in (s // w, w)])
    v = v.permute(0, *[2 + 2 * i for i in range(d)], 1, *[3 + 2 * i for i in range(d)])
    return v.reshape(-1, C, *window_size)


def window_reverse(windows, window_size, size_tuple):
    """inverse of window_partition (ED:662-693)."""
    if windows.dim() not in (4, 5):
        raise NotImplementedError('len(x.shape) [%d] is equal to 4 or 5' % windows.dim())
    d = len(size_tuple)
    C = windows.shape[1]
    grid = [s // w for s, w in zip(size_tuple, window_size)]
    B = windows.shape[0] // _prod(grid)
    v = windows.reshape(B, *grid, C, *window_size)
    v = v.permute(0, d + 1, *[t for i in range(d) for t in (1 + i, d + 2 + i)])
    return v.reshape(B, C, *size_tuple)


def shifted_window_row_map(batch: int, spatial: Sequence[int], window: Sequence[int], shift: Sequence[int],
                           device) -> torch.Tensor:
    """int32 [batch * nW * n]: row (in the natural channels-last voxel order) of token t of window w of
    `window_partition(torch.roll(x, -shift))` — i.e. roll + partition as pure index math (ED:784-790)."""
    d = len(spatial)
    axes = [torch.arange(s) for s in spatial]
    pos = torch.meshgrid(*axes, indexing="ij")                       # coordinates in the rolled volume
    flat = torch.zeros(tuple(spatial), dtype=torch.long)
    for a in range(d):
        flat = flat * spatial[a] + (pos[a] + shift[a]) % spatial[a]   # rolled[p] = x[(p + shift) mod S]
    v = flat.reshape(*[t for s, w in zip(spatial, window) for t in (s // w, w)])
    v = v.permute(*[2 * i for i in range(d)], *[2 * i + 1 for i in range(d)]).reshape(-1)
    V = _prod(spatial)
    rows = (torch.arange(batch).view(-1, 1) * V + v.view(1, -1)).reshape(-1)
    return rows.to(torch.int32).to(device)


class SwinGrapher(nn.Module):
    """Shifted-window grapher: kNN graph inside each window, no attention mask (ED:695-818)."""

    def __init__(self, in_channels, img_shape, kernel_size=9, dilation=1, conv='edge', act='relu', norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0, relative_pos=False,
                 conv_op=nn.Conv3d, norm_op=nn.BatchNorm3d, norm_op_kwargs=None, dropout_op=nn.Dropout3d,
                 window_size=[3, 6, 6], shift_size=[0, 0, 0]):
        super().__init__()
        norm_op_kwargs = norm_op_kwargs or {}
        self.channels = in_channels
        self.r = r
        self.conv_op = conv_op
        self.ndim = _ndim(conv_op)
        self.img_shape = img_shape
        self.window_size = window_size
        self.shift_size = shift_size
        self.fc1 = nn.Sequential(conv_op(in_channels, in_channels, 1, stride=1, padding=0),
                                 norm_op(in_channels, **norm_op_kwargs))
        norm = 'batch'  # the reference overrides the requested norm here (ED:714)
        self.graph_conv = DyGraphConv(in_channels, in_channels * 2, kernel_size, dilation, conv, act, norm, bias,
                                      stochastic, epsilon, r, conv_op, dropout_op)
        self.fc2 = nn.Sequential(conv_op(in_channels * 2, in_channels, 1, stride=1, padding=0),
                                 norm_op(in_channels, **norm_op_kwargs))
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.n = _prod(window_size)
        self.relative_pos = _make_relative_pos(in_channels, self.n, r, self.ndim) if relative_pos else None
        self._row_maps = {}

    def _row_map(self, batch, spatial, device):
        key = (batch, tuple(spatial), str(device))
        if key not in self._row_maps:
            shift = self.shift_size if max(self.shift_size) > 0 else [0] * self.ndim
            self._row_maps[key] = shifted_window_row_map(batch, spatial, self.window_size, shift, device)
        return self._row_maps[key]

    def forward(self, x):
        B = x.shape[0]
        spatial = tuple(x.shape[2:])
        assert spatial == tuple(self.img_shape), "input features has wrong size"
        if self.r != 1:
            raise NotImplementedError("SwinGrapher with reduce ratio r > 1 (never built by NexToU, ED:1003)")
        # fc1 (1x1 conv + BN) is order-invariant over tokens: run it on the un-partitioned volume
        tok, short = ops.fork_tokens(ops.as_tokens(x))
        h = _fc_bn(self.fc1, tok, B)
        row_map = self._row_map(B, spatial, x.device)
        n_windows = B * _prod(spatial) // self.n
        g = self.graph_conv.forward_tokens(h, B, spatial, self.relative_pos, row_map=row_map, graphs=n_windows, n=self.n)
        return _fc_bn_residual(self, self.fc2, g, short, B, spatial)


class PoolGrapher(nn.Module):
    """Pooled global grapher (ED:820-933)."""

    def __init__(self, in_channels, img_shape, kernel_size=9, dilation=1, conv='edge', act='relu', norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0, relative_pos=False,
                 conv_op=nn.Conv3d, norm_op=nn.BatchNorm3d, norm_op_kwargs=None, dropout_op=nn.Dropout3d,
                 img_min_shape=None):
        super().__init__()
        norm_op_kwargs = norm_op_kwargs or {}
        self.channels = in_channels
        self.r = r
        self.conv_op = conv_op
        self.ndim = _ndim(conv_op)
        self.img_shape = img_shape
        self.fc1 = nn.Sequential(conv_op(in_channels, in_channels, 1, stride=1, padding=0),
                                 norm_op(in_channels, **norm_op_kwargs))
        self.graph_conv = PoolDyGraphConv(in_channels, in_channels * 2, kernel_size, dilation, conv, act, norm, bias,
                                          stochastic, epsilon, r, conv_op, dropout_op, img_shape=img_shape,
                                          img_min_shape=img_min_shape)
        self.fc2 = nn.Sequential(conv_op(in_channels * 2, in_channels, 1, stride=1, padding=0),
                                 norm_op(in_channels, **norm_op_kwargs))
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.pool_size = _pool_size_for(img_shape, img_min_shape)
        self.n = _prod(img_shape) // _prod(self.pool_size)
        self.relative_pos = _make_relative_pos(in_channels, self.n, r, self.ndim) if relative_pos else None

    def forward(self, x):
        B, spatial = x.shape[0], tuple(x.shape[2:])
        tok, short = ops.fork_tokens(ops.as_tokens(x))
        h = _fc_bn(self.fc1, tok, B)
        n_now = _prod([s // p for s, p in zip(spatial, self.pool_size)])
        rp = _resized_relative_pos(self.relative_pos, n_now, self.n, self.r, self.ndim)
        return _fc_bn_residual(self, self.fc2, self.graph_conv.forward_tokens(h, B, spatial, rp), short, B, spatial)


# ------------------------------------------------------------------------------------------------------
# per-stage block containers
# ------------------------------------------------------------------------------------------------------
def _k_schedule(img_min_shape, n_levels, ndim):
    """Neighbour counts per GNN level and the dilation cap (ED:960-987, 1040-1067)."""
    n_min = _prod(img_min_shape)
    max_num = int(n_min // ndim)
    max_k = min([2, 4, 8, 16, 32], key=lambda v: abs(v - max_num))
    min_k = max_num // (2 ** ndim)
    k_list = [min(min_k, max_k), min(min_k * 2, max_k), min(min_k * 2, max_k), min(min_k * 4, max_k),
              min(min_k * 8, max_k)]
    if n_levels >= 5:
        k_list = k_list + [min(min_k * 16, max_k)] * (n_levels - 5)
    else:
        k_list = k_list[0:n_levels]
    max_dilation = n_min // max(k_list)
    return k_list, max_dilation


class _GNNBlocks(nn.Module):
    def _common(self, opt, index, conv_op):
        self.n_blocks = sum(opt.blocks)
        dpr = [v.item() for v in torch.linspace(0, opt.drop_path, self.n_blocks)]
        first = sum(opt.blocks[0:index])
        idx_list = [first + j for j in range(opt.blocks[index])]
        ndim = _ndim(conv_op)
        k_list, max_dilation = _k_schedule(opt.img_min_shape, opt.pool_op_kernel_sizes_len, ndim)
        return dpr, idx_list, k_list, max_dilation, ndim

    def forward(self, x):
        return self.blocks(x)


class SwinGNNBlocks(_GNNBlocks):
    """[SwinGrapher, FFN] x blocks[index]  (ED:935-1013)."""

    def __init__(self, channels, img_shape, index, opt=None, conv_op=nn.Conv3d, norm_op=nn.BatchNorm3d,
                 norm_op_kwargs=None, dropout_op=nn.Dropout3d, **kwargs):
        super().__init__()
        dpr, idx_list, k_list, max_dilation, ndim = self._common(opt, index, conv_op)
        window = opt.img_min_shape
        blocks = []
        for idx in idx_list:
            shift = [w // 2 for w in window]
            blocks.append(nn.Sequential(
                SwinGrapher(channels, img_shape, k_list[index], min(idx // 4 + 1, max_dilation), opt.conv, opt.act,
                            opt.norm, opt.bias, opt.use_stochastic, opt.epsilon, 1, n=_prod(window),
                            drop_path=dpr[idx], relative_pos=True, conv_op=conv_op, norm_op=norm_op,
                            norm_op_kwargs=norm_op_kwargs, dropout_op=dropout_op, window_size=window,
                            shift_size=shift),
                FFN(channels, channels * 4, act=opt.act, drop_path=dpr[idx], conv_op=conv_op, norm_op=norm_op,
                    norm_op_kwargs=norm_op_kwargs)))
        self.blocks = nn.Sequential(*blocks)


class PoolGNNBlocks(_GNNBlocks):
    """[PoolGrapher, FFN] x blocks[index]  (ED:1015-1092)."""

    def __init__(self, channels, img_shape, index, stage_num, opt=None, conv_op=nn.Conv3d, norm_op=nn.BatchNorm3d,
                 norm_op_kwargs=None, dropout_op=nn.Dropout3d, **kwargs):
        super().__init__()
        dpr, idx_list, k_list, max_dilation, ndim = self._common(opt, index, conv_op)
        blocks = []
        for idx in idx_list:
            blocks.append(nn.Sequential(
                PoolGrapher(channels, img_shape, k_list[index + stage_num], min(idx // 4 + 1, max_dilation), opt.conv,
                            opt.act, opt.norm, opt.bias, opt.use_stochastic, opt.epsilon,
                            opt.reduce_ratios[index + stage_num], n=opt.n_size_list[index + stage_num],
                            drop_path=dpr[idx], relative_pos=True, conv_op=conv_op, norm_op=norm_op,
                            norm_op_kwargs=norm_op_kwargs, dropout_op=dropout_op, img_min_shape=opt.img_min_shape),
                FFN(channels, channels * 4, act=opt.act, drop_path=dpr[idx], conv_op=conv_op, norm_op=norm_op,
                    norm_op_kwargs=norm_op_kwargs)))
        self.blocks = nn.Sequential(*blocks)


# ------------------------------------------------------------------------------------------------------
# encoder / decoder
# ------------------------------------------------------------------------------------------------------
def _stage_shapes(patch_size, strides, conv_op):
    """Feature-map shape of every stage (ED:70-99, 223-252)."""
    if conv_op not in (nn.Conv2d, nn.Conv3d):
        raise ValueError("unknown convolution dimensionality, conv op: %s" % str(conv_op))
    ndim = _ndim(conv_op)
    cur = [int(v) for v in patch_size[:ndim]]
    shapes = [tuple(cur)]
    for st in strides[1:]:
        cur = [c // int(s) for c, s in zip(cur, st)]
        shapes.append(tuple(cur))
    return shapes, [_prod(s) for s in shapes]


class NexToU_Encoder(nn.Module):
    """conv stages, then [conv, PoolGNNBlocks, SwinGNNBlocks] stages (ED:34-185)."""

    def __init__(self, input_channels: int, patch_size: List[int], n_stages: int,
                 features_per_stage: Union[int, List[int], Tuple[int, ...]], conv_op: Type[nn.Module],
                 kernel_sizes: Union[int, List[int], Tuple[int, ...]], strides: Union[int, List[int], Tuple[int, ...]],
                 n_conv_per_stage: Union[int, List[int], Tuple[int, ...]], conv_bias: bool = False,
                 norm_op=None, norm_op_kwargs: dict = None, dropout_op=None, dropout_op_kwargs: dict = None,
                 nonlin=None, nonlin_kwargs: dict = None, return_skips: bool = False, nonlin_first: bool = False,
                 pool: str = 'conv'):
        super().__init__()
        if isinstance(kernel_sizes, int):
            kernel_sizes = [kernel_sizes] * n_stages
        if isinstance(features_per_stage, int):
            features_per_stage = [features_per_stage] * n_stages
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * n_stages
        if isinstance(strides, int):
            strides = [strides] * n_stages
        assert len(kernel_sizes) == n_stages, "kernel_sizes must have as many entries as we have resolution stages (n_stages)"
        assert len(n_conv_per_stage) == n_stages, "n_conv_per_stage must have as many entries as we have resolution stages (n_stages)"
        assert len(features_per_stage) == n_stages, "features_per_stage must have as many entries as we have resolution stages (n_stages)"
        assert len(strides) == n_stages, "strides must have as many entries as we have resolution stages (n_stages)"

        img_shape_list, n_size_list = _stage_shapes(patch_size, strides, conv_op)
        self.opt = OptInit(pool_op_kernel_sizes_len=len(strides))
        self.opt.img_min_shape = img_shape_list[-1]
        self.opt.n_size_list = n_size_list
        self.n_swin_gnn_stages = 0
        self.no_pool_gnn_stage_num = n_stages - 4
        self.n_conv_stages = self.no_pool_gnn_stage_num - self.n_swin_gnn_stages

        common = (conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs, nonlin_first)
        gnn_kw = dict(opt=self.opt, conv_op=conv_op, norm_op=norm_op, norm_op_kwargs=norm_op_kwargs, dropout_op=dropout_op)
        stages = []
        for s in range(n_stages):
            mods = []
            if pool in ('max', 'avg'):
                st = strides[s]
                if (isinstance(st, int) and st != 1) or (isinstance(st, (tuple, list)) and any(i != 1 for i in st)):
                    mods.append(get_matching_pool_op(conv_op, pool_type=pool)(kernel_size=st, stride=st))
                conv_stride = 1
            elif pool == 'conv':
                conv_stride = strides[s]
            else:
                raise RuntimeError()
            cin, cout = input_channels, features_per_stage[s]
            if s < self.n_conv_stages:
                mods.append(StackedConvBlocks(n_conv_per_stage[s], conv_op, cin, cout, kernel_sizes[s], conv_stride, *common))
            elif s < self.no_pool_gnn_stage_num:  # unreachable while n_swin_gnn_stages == 0 (ED:106, 128-133)
                mods.append(nn.Sequential(
                    StackedConvBlocks(n_conv_per_stage[s] - 1, conv_op, cin, cout, kernel_sizes[s], conv_stride, *common),
                    SwinGNNBlocks(cout, img_shape_list[s], s - self.n_conv_stages, **gnn_kw)))
            else:
                mods.append(nn.Sequential(
                    StackedConvBlocks(n_conv_per_stage[s] - 1, conv_op, cin, cout, kernel_sizes[s], conv_stride, *common),
                    PoolGNNBlocks(cout, img_shape_list[s], s - self.no_pool_gnn_stage_num, self.no_pool_gnn_stage_num, **gnn_kw),
                    SwinGNNBlocks(cout, img_shape_list[s], s - self.n_conv_stages, **gnn_kw)))
            stages.append(nn.Sequential(*mods))
            input_channels = cout

        self.stages = nn.Sequential(*stages)
        self.output_channels = features_per_stage
        self.strides = [maybe_convert_scalar_to_list(conv_op, i) for i in strides]
        self.return_skips = return_skips
        # stored for the decoder
        self.conv_op = conv_op
        self.norm_op = norm_op
        self.norm_op_kwargs = norm_op_kwargs
        self.nonlin = nonlin
        self.nonlin_kwargs = nonlin_kwargs
        self.dropout_op = dropout_op
        self.dropout_op_kwargs = dropout_op_kwargs
        self.conv_bias = conv_bias
        self.kernel_sizes = kernel_sizes

    def forward(self, x):
        ret = []
        last = len(self.stages) - 1
        for i, stage in enumerate(self.stages):
            x = stage(x)
            if self.return_skips and i < last:
                x, skip = ops.fork(x)       # next stage + decoder skip: gradients are summed in the padded token layout
                ret.append(skip)
            else:
                ret.append(x)
        return ret if self.return_skips else ret[-1]

    def compute_conv_feature_map_size(self, input_size):
        output = np.int64(0)
        for s in range(len(self.stages)):
            if isinstance(self.stages[s], nn.Sequential):
                for sq in self.stages[s]:
                    if hasattr(sq, 'compute_conv_feature_map_size'):
                        output += self.stages[s][-1].compute_conv_feature_map_size(input_size)
            else:
                output += self.stages[s].compute_conv_feature_map_size(input_size)
            input_size = [i // j for i, j in zip(input_size, self.strides[s])]
        return output


class NexToU_Decoder(nn.Module):
    """transposed conv -> concat skip -> [conv (+ GNN blocks)] -> 1x1 segmentation head, per stage (ED:187-366)."""

    def __init__(self, encoder: NexToU_Encoder, patch_size: List[int], strides, num_classes: int,
                 n_conv_per_stage: Union[int, Tuple[int, ...], List[int]], deep_supervision, nonlin_first: bool = False):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.encoder = encoder
        self.num_classes = num_classes
        n_enc = len(encoder.output_channels)
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * (n_enc - 1)
        assert len(n_conv_per_stage) == n_enc - 1, "n_conv_per_stage must have as many entries as we have " \
                                                   "resolution stages - 1 (n_stages in encoder - 1), here: %d" % n_enc
        transpconv_op = get_matching_convtransp(conv_op=encoder.conv_op)
        img_shape_list, n_size_list = _stage_shapes(patch_size, strides, encoder.conv_op)
        self.opt = OptInit(pool_op_kernel_sizes_len=len(strides))
        self.opt.img_min_shape = img_shape_list[-1]
        self.opt.n_size_list = n_size_list
        self.n_swin_gnn_stages = 0
        self.no_pool_gnn_stage_num = n_enc - 4
        self.n_conv_stages = self.no_pool_gnn_stage_num - self.n_swin_gnn_stages

        common = (encoder.conv_bias, encoder.norm_op, encoder.norm_op_kwargs, encoder.dropout_op,
                  encoder.dropout_op_kwargs, encoder.nonlin, encoder.nonlin_kwargs, nonlin_first)
        gnn_kw = dict(opt=self.opt, conv_op=encoder.conv_op, norm_op=encoder.norm_op,
                      norm_op_kwargs=encoder.norm_op_kwargs, dropout_op=encoder.dropout_op)
        stages, transpconvs, seg_layers = [], [], []
        for s in range(1, n_enc):
            below = encoder.output_channels[-s]
            skip = encoder.output_channels[-(s + 1)]
            stride = encoder.strides[-s]
            transpconvs.append(transpconv_op(below, skip, stride, stride, bias=encoder.conv_bias))
            level = n_enc - (s + 1)  # encoder stage whose resolution this decoder stage works at
            ks = encoder.kernel_sizes[-(s + 1)]
            if s < (n_enc - self.no_pool_gnn_stage_num):
                stages.append(nn.Sequential(
                    StackedConvBlocks(n_conv_per_stage[s - 1] - 1, encoder.conv_op, 2 * skip, skip, ks, 1, *common),
                    PoolGNNBlocks(skip, img_shape_list[level], level - self.no_pool_gnn_stage_num,
                                  self.no_pool_gnn_stage_num, **gnn_kw),
                    SwinGNNBlocks(skip, img_shape_list[level], level - self.n_conv_stages, **gnn_kw)))
            elif s < (n_enc - self.n_conv_stages):  # unreachable while n_swin_gnn_stages == 0
                stages.append(nn.Sequential(
                    StackedConvBlocks(n_conv_per_stage[s - 1] - 1, encoder.conv_op, 2 * skip, skip, ks, 1, *common),
                    SwinGNNBlocks(skip, img_shape_list[level], level - self.n_conv_stages, **gnn_kw)))
            else:
                stages.append(StackedConvBlocks(n_conv_per_stage[s - 1], encoder.conv_op, 2 * skip, skip, ks, 1, *common))
            # always built so checkpoints load with or without deep supervision (ED:302-305)
            seg_layers.append(encoder.conv_op(skip, num_classes, 1, 1, 0, bias=True))
        self.stages = nn.ModuleList(stages)
        self.transpconvs = nn.ModuleList(transpconvs)
        self.seg_layers = nn.ModuleList(seg_layers)

    def forward(self, skips):
        low = skips[-1]
        seg_outputs = []
        last = len(self.stages) - 1
        for s in range(len(self.stages)):
            tc = self.transpconvs[s]
            skip = skips[-(s + 2)]
            cat, gap = dense.up_cat(low, tc, skip)                           # torch.cat((transpconv(low), skip), 1), ED:321-322
            first = self.stages[s][0] if isinstance(self.stages[s], nn.Sequential) else self.stages[s]
            first.convs[0].conv.in_gap = gap                                 # layout of the tensor it is about to receive
            x = self.stages[s](cat)
            if self.deep_supervision or s == last:
                head = self.seg_layers[s if self.deep_supervision else -1]
                B, spatial = x.shape[0], tuple(x.shape[2:])
                if s != last:
                    x, xs = ops.fork(x)     # segmentation head + next decoder stage
                else:
                    xs = x
                seg_outputs.append(ops.from_tokens(dense.linear_tokens(ops.as_tokens(xs), head), B, spatial))
            low = x
        seg_outputs = seg_outputs[::-1]  # highest resolution first
        return seg_outputs if self.deep_supervision else seg_outputs[0]

    def compute_conv_feature_map_size(self, input_size):
        skip_sizes = []
        for s in range(len(self.encoder.strides) - 1):
            skip_sizes.append([i // j for i, j in zip(input_size, self.encoder.strides[s])])
            input_size = skip_sizes[-1]
        assert len(skip_sizes) == len(self.stages)
        output = np.int64(0)
        for s in range(len(self.stages)):
            output += self.stages[s].compute_conv_feature_map_size(skip_sizes[-(s + 1)])
            output += np.prod([self.encoder.output_channels[-(s + 2)], *skip_sizes[-(s + 1)]], dtype=np.int64)
            if self.deep_supervision or (s == (len(self.stages) - 1)):
                output += np.prod([self.num_classes, *skip_sizes[-(s + 1)]], dtype=np.int64)
        return output
