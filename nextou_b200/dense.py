"""Dense building blocks of the hot path on token-major activations: pointwise (1x1) convolutions as GEMMs,
grouped pointwise convolutions, spatial / transposed convolutions, batch / instance normalisation (+ LeakyReLU).

Every module in blocks.py / conv_blocks.py routes through this file, so it is the one place where a backend is
chosen.  `stats` counts calls per backend (bench.py reports them):

  * "native.*"  — hand-written sm_100a kernels behind the C-ABI (csrc/norm.cu; csrc/gemm_tcgen05.cu when enabled);
  * "cublas.*"  — plain library GEMMs (1x1 convolutions are exactly `tokens @ W^T + b`);
  * "cudnn.*"   — spatial / transposed convolutions that the native implicit-GEMM kernel does not cover yet.

Nothing here ever runs on the CPU.  Activations are logically (N, C, *spatial), physically channels-last; the
token view [N*prod(spatial), C] of such a tensor is free (ops.as_tokens).
"""
from __future__ import annotations

from collections import Counter
from typing import Optional

import torch
import torch.nn.functional as F

from . import ops

stats: Counter = Counter()


# ------------------------------------------------------------------------------------------------------
# pointwise convolutions == GEMMs over tokens (ED:373-381, 710-720, 833-842, 305; grouped: TN:85)
# ------------------------------------------------------------------------------------------------------
def linear_tokens(tok: torch.Tensor, conv: torch.nn.Module) -> torch.Tensor:
    """1x1 conv of `conv` (weight (Cout, Cin, 1, 1[, 1])) applied to token rows: [T, Cin] -> [T, Cout]."""
    stats["cublas.linear"] += 1
    w = conv.weight.reshape(conv.weight.shape[0], -1)
    return F.linear(tok, w, conv.bias)


def grouped_linear_tokens(tok: torch.Tensor, conv: torch.nn.Module) -> torch.Tensor:
    """Grouped 1x1 conv (groups = 6 in 3-D, 4 in 2-D; weight (Cout, Cin/g, 1, ...)): the groups are the diagonal
    blocks of one dense GEMM.  The zero blocks cost (g-1)/g wasted FLOPs of a tiny GEMM (2C <= 648) but keep the
    output token-major without a permute copy."""
    stats["cublas.grouped_linear"] += 1
    g = conv.groups
    w = conv.weight.reshape(g, conv.weight.shape[0] // g, -1)
    return F.linear(tok, torch.block_diag(*w.unbind(0)), conv.bias)


# ------------------------------------------------------------------------------------------------------
# spatial convolutions (ED:125-141, 281-300) and kernel == stride transposed convolutions (ED:273-276)
# ------------------------------------------------------------------------------------------------------
def conv_nd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride, padding, groups: int = 1):
    stats["cudnn.conv"] += 1
    f = F.conv3d if x.dim() == 5 else F.conv2d
    return f(x, weight, bias, stride=stride, padding=padding, groups=groups)


def conv_transpose_nd(x, weight, bias, stride):
    stats["cudnn.conv_transpose"] += 1
    f = F.conv_transpose3d if x.dim() == 5 else F.conv_transpose2d
    return f(x, weight, bias, stride=stride)


# ------------------------------------------------------------------------------------------------------
# normalisation (+ LeakyReLU), csrc/norm.cu
# ------------------------------------------------------------------------------------------------------
def batch_norm_tokens(tok: torch.Tensor, bn: torch.nn.modules.batchnorm._BatchNorm, act_slope: Optional[float] = None):
    """nn.BatchNorm semantics on token rows: batch statistics + running-stat update in training (or when the module
    tracks no running stats), running statistics in eval."""
    slope = 1.0 if act_slope is None else act_slope
    use_batch_stats = bn.training or bn.running_mean is None
    if use_batch_stats:
        stats["native.batch_norm"] += 1
        rm = rv = None
        momentum = 0.0
        if bn.training and bn.track_running_stats and bn.running_mean is not None:
            rm, rv = bn.running_mean, bn.running_var
            bn.num_batches_tracked.add_(1)
            momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        return ops.norm_act_tokens(tok, bn.weight, bn.bias, rm, rv, momentum, bn.eps, slope, 1)
    stats["native.affine_act"] += 1
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    scale = inv if bn.weight is None else inv * bn.weight.float()
    shift = -bn.running_mean.float() * scale
    if bn.bias is not None:
        shift = shift + bn.bias.float()
    return ops.affine_act_tokens(tok, scale, shift, slope)


def instance_norm_tokens(tok: torch.Tensor, inorm: torch.nn.modules.instancenorm._InstanceNorm, batch: int,
                         act_slope: Optional[float] = None):
    """nn.InstanceNorm (no running stats, as built by torch_nn.norm_layer): one statistics instance per batch item."""
    if inorm.track_running_stats:
        raise NotImplementedError("InstanceNorm with running statistics is never built by NexToU (TN:42-46)")
    stats["native.instance_norm"] += 1
    slope = 1.0 if act_slope is None else act_slope
    return ops.norm_act_tokens(tok, inorm.weight, inorm.bias, None, None, 0.0, inorm.eps, slope, batch)


def norm_tokens(tok, norm_mod, batch: int, act_slope: Optional[float] = None):
    if isinstance(norm_mod, torch.nn.modules.batchnorm._BatchNorm):
        return batch_norm_tokens(tok, norm_mod, act_slope)
    if isinstance(norm_mod, torch.nn.modules.instancenorm._InstanceNorm):
        return instance_norm_tokens(tok, norm_mod, batch, act_slope)
    raise NotImplementedError("normalisation module %s" % type(norm_mod).__name__)
