"""kNN graph construction — API mirror of the reference's network_architecture/torch_edge.py.

Same callables, argument meaning and output format (`(2, B, N, k)` int64 edge_index: row 0 neighbour index in
ascending distance, row 1 centre index), but the arithmetic runs in the fused CUDA kernels of
csrc/knn.cu (row normalise -> distance tile -> running top-k; the N x M matrix is never materialised).
Distances are always fp32 FMA (never bf16), ties go to the lowest index.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def _tokens_from_bcn1(x: torch.Tensor) -> torch.Tensor:
    """(B, C, N, 1) -> token-major [B*N, C] (a view when x is channels-last, one transpose copy otherwise)."""
    B, C, N, _ = x.shape
    return x.squeeze(-1).transpose(1, 2).reshape(B * N, C)


def _edge_index(nn_idx: torch.Tensor) -> torch.Tensor:
    B, N, k = nn_idx.shape
    center = torch.arange(N, device=nn_idx.device).view(1, N, 1).expand(B, N, k)
    return torch.stack((nn_idx, center), dim=0)


def pairwise_distance(x):
    """x (B, N, C) -> (B, N, N) squared distances (TE:12-23).  Debug helper: materialises the matrix with torch ops;
    the hot path (dense_knn_matrix) never calls it."""
    with torch.no_grad():
        inner = -2 * torch.matmul(x, x.transpose(2, 1))
        sq = torch.sum(torch.mul(x, x), dim=-1, keepdim=True)
        return sq + inner + sq.transpose(2, 1)


def xy_pairwise_distance(x, y):
    """(B, N, C), (B, M, C) -> (B, N, M) (TE:42-55).  Debug helper, see pairwise_distance."""
    with torch.no_grad():
        inner = -2 * torch.matmul(x, y.transpose(2, 1))
        return torch.sum(x * x, -1, keepdim=True) + inner + torch.sum(y * y, -1, keepdim=True).transpose(2, 1)


def _knn(x, y, k, relative_pos, normalize):
    """x: (B, C, N, 1), y: (B, C, M, 1) or None -> (B, N, k) int64.  normalize=True folds F.normalize (TE:154-160)
    into the kernel; False reproduces dense_knn_matrix called directly on raw features."""
    B, C, N, _ = x.shape
    xt = _tokens_from_bcn1(x)
    if y is None:
        idx, _ = ops.knn_graph(xt, B, N, relpos=relative_pos, k=k, dilation=1, normalize=normalize)
    else:
        M = y.shape[2]
        idx, _ = ops.knn_graph(xt, B, N, _tokens_from_bcn1(y), M, relpos=relative_pos, k=k, dilation=1,
                               normalize=normalize)
    return idx


def dense_knn_matrix(x, k=16, relative_pos=None):
    """x (B, C, N, 1) -> edge_index (2, B, N, k)  (TE:58-90; the 10 000-row chunking is unnecessary here)."""
    with torch.no_grad():
        return _edge_index(_knn(x, None, k, relative_pos, normalize=False))


def xy_dense_knn_matrix(x, y, k=16, relative_pos=None):
    """(TE:93-110)."""
    with torch.no_grad():
        return _edge_index(_knn(x, y, k, relative_pos, normalize=False))


def draw_stochastic_columns(k: int, dilation: int, stochastic: bool, epsilon: float, training: bool):
    """Mirror of the host-RNG consumption in DenseDilated.forward (TE:126-136): `torch.rand(1)` is drawn on every call
    (also in eval), `torch.randperm(k*dilation)` only when the random branch is taken.  Returns the column
    selection (LongTensor) for the random branch or None for the regular `[::dilation]` branch."""
    if stochastic:
        if torch.rand(1) < epsilon and training:
            return torch.randperm(k * dilation)[:k]
    return None


class DenseDilated(nn.Module):
    """Pick the dilated neighbours from a (2, B, N, k*dilation) neighbour list (TE:113-136)."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k

    def forward(self, edge_index):
        cols = draw_stochastic_columns(self.k, self.dilation, self.stochastic, self.epsilon, self.training)
        if cols is not None:
            return edge_index[:, :, :, cols.to(edge_index.device)]
        return edge_index[:, :, :, ::self.dilation]


class DenseDilatedKnnGraph(nn.Module):
    """Normalise, build the dense kNN graph, dilate (TE:139-163).  x: (B, C, N, 1), y: (B, C, M, 1) or None."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)

    def forward(self, x, y=None, relative_pos=None):
        with torch.no_grad():
            idx = _knn(x, y, self.k * self.dilation, relative_pos, normalize=True)
            return self._dilated(_edge_index(idx))
