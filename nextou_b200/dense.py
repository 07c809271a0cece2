"""Dense (GEMM-shaped) building blocks of the hot path: spatial conv, 1x1 conv, transposed conv, norms.

Every module in blocks.py / model.py routes its contraction and normalisation work through the functions in
this file, so this is the single place where a backend is chosen.  Backends:

  * "tcgen05"  — hand-written sm_100a kernels behind the C-ABI (csrc/gemm_tcgen05.cu, csrc/norm.cu), used for
                 every shape they support;
  * "library"  — cuDNN / cuBLAS through torch.nn.functional, used for the shapes the native kernels do not
                 cover yet.  It is a GPU library call (never a CPU fallback); DESIGN.md lists exactly which
                 shapes still take it and `dense.stats` counts the calls so bench.py can report them.

Activations are logically (N, C, *spatial) and physically channels-last.
"""
from __future__ import annotations

from collections import Counter
from typing import Optional, Sequence

import torch
import torch.nn.functional as F

stats: Counter = Counter()


def _fmt(x):
    return torch.channels_last_3d if x.dim() == 5 else torch.channels_last


def conv_nd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride, padding, groups: int = 1):
    """Spatial / pointwise convolution (ED:125-141, 281-300, 305, 710-720 ...)."""
    stats["library.conv"] += 1
    f = F.conv3d if x.dim() == 5 else F.conv2d
    return f(x, weight, bias, stride=stride, padding=padding, groups=groups)


def conv_transpose_nd(x, weight, bias, stride):
    """kernel == stride transposed convolution (ED:273-276, 321)."""
    stats["library.conv_transpose"] += 1
    f = F.conv_transpose3d if x.dim() == 5 else F.conv_transpose2d
    return f(x, weight, bias, stride=stride)


def batch_norm(x, bn: torch.nn.modules.batchnorm._BatchNorm, act_slope: Optional[float] = None):
    """BatchNorm (train: batch statistics + running-stat update, eval: running stats) [+ LeakyReLU]."""
    stats["library.batch_norm"] += 1
    y = bn(x)
    if act_slope is not None:
        y = F.leaky_relu(y, act_slope, inplace=True)
    return y


def instance_norm(x, inorm: torch.nn.modules.instancenorm._InstanceNorm, act_slope: Optional[float] = None):
    stats["library.instance_norm"] += 1
    y = inorm(x)
    if act_slope is not None:
        y = F.leaky_relu(y, act_slope, inplace=True)
    return y
