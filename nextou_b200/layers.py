"""Small layer factories — API mirror of the reference's network_architecture/torch_nn.py (TN:13-115).

`BasicConv` keeps the reference's parameter layout (an nn.Sequential of grouped 1x1 conv, norm, activation:
state_dict keys `nn.0.*`, `nn.1.*`) but runs through nextou_b200.dense.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import Linear as Lin
from torch.nn import Sequential as Seq

from . import dense


def act_layer(act, inplace=True, neg_slope=1e-2, n_prelu=1):
    """Activation by name (TN:13-29)."""
    name = act.lower()
    if name == "relu":
        return nn.ReLU(inplace)
    if name == "leakyrelu":
        return nn.LeakyReLU(neg_slope, inplace)
    if name == "prelu":
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    if name == "gelu":
        return nn.GELU()
    if name == "hswish":
        return nn.Hardswish(inplace)
    raise NotImplementedError("activation layer [%s] is not found" % name)


_NORMS = {
    ("batch", nn.Conv2d): nn.BatchNorm2d, ("batch", nn.Conv3d): nn.BatchNorm3d,
    ("instance", nn.Conv2d): nn.InstanceNorm2d, ("instance", nn.Conv3d): nn.InstanceNorm3d,
}


def norm_layer(norm, nc, conv_op):
    """Affine batch / instance norm matching the conv dimensionality (TN:32-51)."""
    name = norm.lower()
    if name not in ("batch", "instance"):
        raise NotImplementedError("normalization layer [%s] is not found" % name)
    if (name, conv_op) not in _NORMS:
        raise NotImplementedError("conv operation [%s] is not found" % conv_op)
    return _NORMS[(name, conv_op)](nc, affine=True)


class MLP(Seq):
    """Unused by NexToU; kept for API parity (TN:54-63)."""

    def __init__(self, channels, act="relu", norm=None, bias=True, conv_op=nn.Conv3d):
        m = []
        for i in range(1, len(channels)):
            m.append(Lin(channels[i - 1], channels[i], bias))
            if act is not None and act.lower() != "none":
                m.append(act_layer(act))
            if norm is not None and norm.lower() != "none":
                m.append(norm_layer(norm, channels[-1], conv_op))
        super().__init__(*m)


def groups_for(conv_op) -> int:
    """4 groups in 2-D, 6 in 3-D (TN:74-82)."""
    if conv_op == nn.Conv2d:
        return 4
    if conv_op == nn.Conv3d:
        return 6
    raise NotImplementedError("conv operation [%s] is not found" % conv_op)


class BasicConv(Seq):
    """Grouped 1x1 conv -> norm -> activation, per entry of `channels` (TN:66-92)."""

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0.0, conv_op=nn.Conv3d, dropout_op=None):
        self.conv_op = conv_op
        self.groups_num = groups_for(conv_op)
        self.batch_norm = nn.BatchNorm2d if conv_op == nn.Conv2d else nn.BatchNorm3d
        self.instance_norm = nn.InstanceNorm2d if conv_op == nn.Conv2d else nn.InstanceNorm3d
        m = []
        for i in range(1, len(channels)):
            m.append(conv_op(channels[i - 1], channels[i], 1, bias=bias, groups=self.groups_num))
            if norm is not None and norm.lower() != "none":
                m.append(norm_layer(norm, channels[-1], conv_op))
            if act is not None and act.lower() != "none":
                m.append(act_layer(act))
        super().__init__(*m)

    def forward_tokens(self, tok, batch: int):
        """Token-major fast path: [rows, Cin] -> [rows, Cout]; `batch` = number of instance-norm instances."""
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, (nn.Conv2d, nn.Conv3d)) and isinstance(nxt, (nn.modules.batchnorm._BatchNorm,
                                                                          nn.modules.instancenorm._InstanceNorm)):
                act = mods[i + 2] if i + 2 < len(mods) else None
                slope = act.negative_slope if isinstance(act, nn.LeakyReLU) else None
                tok = dense.linear_norm_act_tokens(tok, m, nxt, batch, slope)
                i += 2 if slope is not None else 1
            elif isinstance(m, (nn.Conv2d, nn.Conv3d)):
                tok = dense.grouped_linear_tokens(tok, m) if m.groups > 1 else dense.linear_tokens(tok, m)
            elif isinstance(m, (nn.modules.batchnorm._BatchNorm, nn.modules.instancenorm._InstanceNorm)):
                slope = nxt.negative_slope if isinstance(nxt, nn.LeakyReLU) else None
                tok = dense.norm_tokens(tok, m, batch, slope)
                if slope is not None:
                    i += 1
            else:
                tok = m(tok)
            i += 1
        return tok

    def forward(self, x):
        from . import ops
        B, spatial = x.shape[0], tuple(x.shape[2:])
        return ops.from_tokens(self.forward_tokens(ops.as_tokens(x), B), B, spatial)


def batched_index_select(x, idx):
    """x (B, C, M, 1), idx (B, N, k) -> (B, C, N, k) neighbour features (TN:94-115).  Compatibility helper that
    materialises the gathered tensor; the hot path uses ops.mrconv_gather instead and never builds it."""
    B, C = x.shape[:2]
    _, N, k = idx.shape
    flat = x.reshape(B, C, -1).gather(2, idx.reshape(B, 1, N * k).expand(B, C, N * k))
    return flat.reshape(B, C, N, k)
