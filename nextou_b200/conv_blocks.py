"""conv -> (dropout) -> norm -> nonlin stacks used by every U-Net stage.

The reference takes these from the un-vendored `dynamic_network_architectures` package
(call sites ED:125-141, 281-300; helpers ED:7-8, NX:9, TR:10-11).  They are re-implemented here with the same
constructor signature, attribute names (`convs[i].conv / .norm / .nonlin / .all_modules`) and therefore the
same state_dict keys, so nnU-Net checkpoints load; the arithmetic goes through nextou_b200.dense.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import dense, ops


def convert_conv_op_to_dim(conv_op) -> int:
    table = {nn.Conv1d: 1, nn.Conv2d: 2, nn.Conv3d: 3}
    if conv_op not in table:
        raise ValueError("Unknown dimension. Only 1d 2d and 3d conv are supported. got %s" % str(conv_op))
    return table[conv_op]


def convert_dim_to_conv_op(dim: int):
    table = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
    if dim not in table:
        raise ValueError("Unknown dimension. Only 1, 2 and 3 are supported")
    return table[dim]


def maybe_convert_scalar_to_list(conv_op, scalar):
    if isinstance(scalar, (tuple, list, np.ndarray)):
        return scalar
    return [scalar] * convert_conv_op_to_dim(conv_op)


def get_matching_convtransp(conv_op):
    return {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}[convert_conv_op_to_dim(conv_op)]


def get_matching_batchnorm(conv_op):
    return {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[convert_conv_op_to_dim(conv_op)]


def get_matching_pool_op(conv_op, adaptive=False, pool_type="avg"):
    d = convert_conv_op_to_dim(conv_op)
    name = ("Adaptive" if adaptive else "") + {"avg": "Avg", "max": "Max"}[pool_type] + f"Pool{d}d"
    return getattr(nn, name)


class InitWeights_He:
    """kaiming_normal_(a=neg_slope) on conv / transposed-conv weights, zero bias (TR:88)."""

    def __init__(self, neg_slope: float = 1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


class ConvDropoutNormReLU(nn.Module):
    def __init__(self, conv_op, input_channels, output_channels, kernel_size, stride, conv_bias=False, norm_op=None,
                 norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None, nonlin_kwargs=None,
                 nonlin_first=False):
        super().__init__()
        self.input_channels = input_channels
        self.output_channels = output_channels
        stride = maybe_convert_scalar_to_list(conv_op, stride)
        kernel_size = maybe_convert_scalar_to_list(conv_op, kernel_size)
        self.stride = stride
        self.nonlin_first = nonlin_first
        ops = []
        self.conv = conv_op(input_channels, output_channels, kernel_size, stride,
                            padding=[(k - 1) // 2 for k in kernel_size], dilation=1, bias=conv_bias)
        ops.append(self.conv)
        if dropout_op is not None:
            self.dropout = dropout_op(**(dropout_op_kwargs or {}))
            ops.append(self.dropout)
        if norm_op is not None:
            self.norm = norm_op(output_channels, **(norm_op_kwargs or {}))
            ops.append(self.norm)
        if nonlin is not None:
            self.nonlin = nonlin(**(nonlin_kwargs or {}))
            ops.append(self.nonlin)
        if nonlin_first and norm_op is not None and nonlin is not None:
            ops[-1], ops[-2] = ops[-2], ops[-1]
        self.all_modules = nn.Sequential(*ops)

    def forward(self, x):
        mods = list(self.all_modules)
        B, spatial = x.shape[0], tuple(x.shape[2:])
        tok = ops.as_tokens(x)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if m is self.conv and isinstance(nxt, (nn.modules.batchnorm._BatchNorm, nn.modules.instancenorm._InstanceNorm)):
                # conv -> norm (-> LeakyReLU): one call, so that inference can fold the norm into the conv epilogue
                act = mods[i + 2] if i + 2 < len(mods) else None
                fuse = isinstance(act, nn.LeakyReLU)
                tok, spatial = dense.conv_norm_act_tokens(tok, B, spatial, m, nxt, act.negative_slope if fuse else None)
                i += 2 if fuse else 1
            elif m is self.conv:
                tok, spatial = dense.conv_tokens(tok, B, spatial, m)
            elif isinstance(m, (nn.modules.batchnorm._BatchNorm, nn.modules.instancenorm._InstanceNorm)):
                fuse = isinstance(nxt, nn.LeakyReLU)
                tok = dense.norm_tokens(tok, m, B, nxt.negative_slope if fuse else None)
                i += 1 if fuse else 0
            else:  # dropout / other activations: through the logical (N, C, *spatial) view
                tok = ops.as_tokens(m(ops.from_tokens(tok, B, spatial)))
            i += 1
        return ops.from_tokens(tok, B, spatial)

    def compute_conv_feature_map_size(self, input_size):
        assert len(input_size) == len(self.stride)
        out = [i // j for i, j in zip(input_size, self.stride)]
        return np.prod([self.output_channels, *out], dtype=np.int64)


class StackedConvBlocks(nn.Module):
    def __init__(self, num_convs, conv_op, input_channels, output_channels, kernel_size, initial_stride,
                 conv_bias=False, norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None,
                 nonlin=None, nonlin_kwargs=None, nonlin_first=False):
        super().__init__()
        if not isinstance(output_channels, (tuple, list)):
            output_channels = [output_channels] * num_convs
        common = (conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs,
                  nonlin_first)
        blocks = [ConvDropoutNormReLU(conv_op, input_channels, output_channels[0], kernel_size, initial_stride,
                                      *common)]
        for i in range(1, num_convs):
            blocks.append(ConvDropoutNormReLU(conv_op, output_channels[i - 1], output_channels[i], kernel_size, 1,
                                              *common))
        self.convs = nn.Sequential(*blocks)
        self.output_channels = output_channels[-1]
        self.initial_stride = maybe_convert_scalar_to_list(conv_op, initial_stride)

    def forward(self, x):
        return self.convs(x)

    def compute_conv_feature_map_size(self, input_size):
        assert len(input_size) == len(self.initial_stride)
        output = self.convs[0].compute_conv_feature_map_size(input_size)
        after = [i // j for i, j in zip(input_size, self.initial_stride)]
        for b in self.convs[1:]:
            output += b.compute_conv_feature_map_size(after)
        return output
