"""Topological-interaction losses — API mirror of the reference's loss/bti_loss.py, loss/ti_loss.py,
loss/compound_bti_loss.py and loss/compound_ti_loss.py.

`BTI_Loss.forward(x, y)`: x logits (b, c, *spatial), y label map (b, 1, *spatial) -> fp64 scalar
= mean_b sum_voxels CE(x, y) * critical_map, where the critical map marks voxels whose argmax class violates an
inclusion / exclusion interaction inside the 3^d (or cross) neighbourhood (BTI:76-145).  The reference evaluates
this with 2 fp64 convolutions per interaction; here it is one bit-mask morphology kernel (csrc/bti.cu).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
from torch import nn

from . import ops


def _class_bits(v, singleton: bool) -> int:
    """A / C entry of an interaction -> class bit mask.  Entries are python ints, 0-d tensors (single label) or
    1-d tensors / lists (label sets, BTI only) — the trainers' make_tensors format (…_BTI_Synapse.py:9-15, 43-47)."""
    if isinstance(v, torch.Tensor):
        arr = v.detach().cpu().reshape(-1).tolist()
    else:
        arr = np.asarray(v).reshape(-1).tolist()
    if singleton and len(arr) != 1:
        raise ValueError("TI_Loss interactions take single labels, got %r (use BTI_Loss for label sets)" % (arr,))
    bits = 0
    for c in arr:
        c = int(c)
        if not 0 <= c < 32:
            raise ValueError("class label %d outside [0, 32): the bit-mask kernel supports up to 32 classes" % c)
        bits |= 1 << c
    return bits


class _InteractionLoss(nn.Module):
    _singleton = False

    def __init__(self, dim=3, connectivity=26, inclusion=[], exclusion=[], min_thick=1):
        """
        :param dim: 2 if 2D; 3 if 3D
        :param connectivity: 4 or 8 for 2D; 6 or 26 for 3D
        :param inclusion: list of [A,B] classes where A is completely surrounded by B.
        :param exclusion: list of [A,C] classes where A and C exclude each other.
        :param min_thick: minimum separation between the two classes (only with connectivity 8 / 26)
        """
        super().__init__()
        self.dim = dim
        self.connectivity = connectivity
        self.min_thick = min_thick
        self.interaction_list = []
        self.sum_dim_list = [1, 2, 3] if dim == 2 else [1, 2, 3, 4]
        self.apply_nonlin = lambda x: torch.nn.functional.softmax(x, 1)
        self.set_kernel()
        for inc in inclusion:
            self.interaction_list.append([True, inc[0], inc[1]])
        for exc in exclusion:
            self.interaction_list.append([False, exc[0], exc[1]])
        self._bits_memo = {}

    def set_kernel(self):
        """Connectivity structuring element (BTI:52-73); only its shape parameters reach the CUDA kernel."""
        k = 2 * self.min_thick + 1
        valid = {2: (4, 8), 3: (6, 26)}
        if self.dim not in valid or self.connectivity not in valid[self.dim]:
            raise ValueError("connectivity %r is not valid for dim %r" % (self.connectivity, self.dim))
        if self.connectivity in (8, 26):
            np_kernel = np.ones((k,) * self.dim)
        else:
            np_kernel = np.zeros((3,) * self.dim)
            centre = (1,) * self.dim
            np_kernel[centre] = 1
            for ax in range(self.dim):
                for off in (0, 2):
                    pos = list(centre)
                    pos[ax] = off
                    np_kernel[tuple(pos)] = 1
        self.kernel = torch.from_numpy(np_kernel[None, None])

    def interaction_table(self):
        """(maskA, maskC, inclusion flag) lists, one entry per interaction, bit c = class c."""
        # rebuilt from the public `interaction_list` on every call, like the reference re-reads it (BTI:84-98): a few
        # dozen small ints; label tensors are converted once and remembered by identity (they may live on the device)
        memo = self._bits_memo
        def bits(v):
            if not isinstance(v, torch.Tensor):
                return _class_bits(v, self._singleton)
            key = (id(v), v._version)
            hit = memo.get(key)
            if hit is None or hit[0] is not v:
                hit = memo[key] = (v, _class_bits(v, self._singleton))
            return hit[1]
        ma = [bits(it[1]) for it in self.interaction_list]
        mc = [bits(it[2]) for it in self.interaction_list]
        inc = [1 if it[0] else 0 for it in self.interaction_list]
        if len(memo) > 4 * max(len(self.interaction_list), 8):
            memo.clear()
        return (ma, mc, inc)

    def critical_voxels_map(self, P: torch.Tensor) -> torch.Tensor:
        """P: discrete segmentation (b, 1, *spatial) or (b, *spatial) -> double map like the reference (BTI:76-117)."""
        lab = P.reshape(P.shape[0], *P.shape[-self.dim:]).to(torch.uint8)
        crit = ops.bti_critical_map(lab, *self.interaction_table(), self.connectivity, self.min_thick)
        return crit.unsqueeze(1).double()

    def forward(self, x, y):
        """x: logits (b, c, *spatial); y: labels (b, 1, *spatial) in [0, c) -> fp64 scalar (BTI:120-145)."""
        if not self.interaction_list:
            raise ValueError("no interactions configured")  # the reference fails with UnboundLocalError here
        return ops.bti_loss(x, y, *self.interaction_table(), self.connectivity, self.min_thick)


class BTI_Loss(_InteractionLoss):
    """Binary topological interaction loss: A and C are label SETS (torch.isin, BTI:90-98)."""
    _singleton = False

    def binary_topological_interaction_module(self, P):
        return self.critical_voxels_map(P)


class TI_Loss(_InteractionLoss):
    """Topological interaction loss: A and C are single labels (P == label, loss/ti_loss.py:89-98)."""
    _singleton = True

    def topological_interaction_module(self, P):
        return self.critical_voxels_map(P)


# --------------------------------------------------------------------------------------------------
# compound losses (loss/compound_bti_loss.py:8-61, loss/compound_ti_loss.py:8-61)
# --------------------------------------------------------------------------------------------------
try:  # inside an nnU-Net checkout use the upstream pieces, exactly like the reference does
    from nnunetv2.training.loss.dice import SoftDiceLoss  # type: ignore
    from nnunetv2.training.loss.robust_ce_loss import RobustCrossEntropyLoss  # type: ignore
    from nnunetv2.utilities.helpers import softmax_helper_dim1  # type: ignore
except Exception:  # standalone: minimal equivalents of the upstream losses
    def softmax_helper_dim1(x: torch.Tensor) -> torch.Tensor:
        return torch.softmax(x, 1)

    class RobustCrossEntropyLoss(nn.CrossEntropyLoss):
        def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
            if target.ndim == input.ndim:
                assert target.shape[1] == 1
                target = target[:, 0]
            return super().forward(input, target.long())

    class SoftDiceLoss(nn.Module):
        """Soft Dice with the upstream kwargs (batch_dice, do_bg, smooth, ddp); memory-efficient formulation."""

        def __init__(self, apply_nonlin=None, batch_dice: bool = False, do_bg: bool = True, smooth: float = 1.,
                     ddp: bool = True, clip_tp: float = None):
            super().__init__()
            self.do_bg, self.batch_dice, self.apply_nonlin, self.smooth, self.ddp = do_bg, batch_dice, apply_nonlin, smooth, ddp

        def forward(self, x, y, loss_mask=None):
            if self.apply_nonlin is not None:
                x = self.apply_nonlin(x)
            axes = tuple(range(2, x.ndim))
            with torch.no_grad():
                if x.ndim != y.ndim:
                    y = y.view((y.shape[0], 1, *y.shape[1:]))
                if x.shape == y.shape:
                    y_onehot = y
                else:
                    y_onehot = torch.zeros(x.shape, device=x.device, dtype=torch.bool)
                    y_onehot.scatter_(1, y.long(), 1)
                if not self.do_bg:
                    y_onehot = y_onehot[:, 1:]
                sum_gt = y_onehot.sum(axes) if loss_mask is None else (y_onehot * loss_mask).sum(axes)
            if not self.do_bg:
                x = x[:, 1:]
            if loss_mask is None:
                intersect = (x * y_onehot).sum(axes)
                sum_pred = x.sum(axes)
            else:
                intersect = (x * y_onehot * loss_mask).sum(axes)
                sum_pred = (x * loss_mask).sum(axes)
            if self.batch_dice:
                if self.ddp and torch.distributed.is_available() and torch.distributed.is_initialized():
                    from torch.distributed.nn.functional import all_gather
                    intersect = torch.stack(all_gather(intersect)).sum(0)
                    sum_pred = torch.stack(all_gather(sum_pred)).sum(0)
                    sum_gt = torch.stack(all_gather(sum_gt.float())).sum(0)
                intersect, sum_pred, sum_gt = intersect.sum(0), sum_pred.sum(0), sum_gt.sum(0)
            dc = (2 * intersect + self.smooth) / torch.clip(sum_gt + sum_pred + self.smooth, 1e-8)
            return -dc.mean()

MemoryEfficientSoftDiceLoss = SoftDiceLoss


class _CompoundInteractionLoss(nn.Module):
    _ti_class = BTI_Loss

    def __init__(self, soft_dice_kwargs, ce_kwargs, ti_kwargs, weight_ce=1, weight_dice=1, weight_ti=1e-6,
                 ignore_label=None, dice_class=SoftDiceLoss):
        """weight_ce * CE + weight_dice * Dice + weight_ti * (B)TI; weights need not sum to one."""
        super().__init__()
        if ignore_label is not None:
            ce_kwargs['ignore_index'] = ignore_label
        self.weight_dice = weight_dice
        self.weight_ce = weight_ce
        self.weight_ti = weight_ti
        self.ignore_label = ignore_label
        self.ce = RobustCrossEntropyLoss(**ce_kwargs)
        self.dc = dice_class(apply_nonlin=softmax_helper_dim1, **soft_dice_kwargs)
        self.ti = self._ti_class(**ti_kwargs)
        self.fused = True   # False: compose the three terms from separate ops like the reference does

    def _fusable(self, net_output, target) -> bool:
        """The one-kernel path covers the configuration every reference trainer builds (…_BTI_Synapse.py:50-58):
        hard labels, no ignore label, default CE arguments, softmax Dice without tp clipping."""
        ce, dc = self.ce, self.dc
        return (self.fused and net_output.is_cuda and self.ignore_label is None and target.shape[1] == 1
                and target.ndim == net_output.ndim and net_output.shape[1] in ops.SEG_LOSS_CLASSES
                and ce.weight is None and ce.reduction == "mean" and ce.label_smoothing == 0.0 and ce.ignore_index == -100
                and type(dc).__name__ in ("SoftDiceLoss", "MemoryEfficientSoftDiceLoss")
                and getattr(dc, "clip_tp", None) is None and dc.apply_nonlin is softmax_helper_dim1
                and (self.weight_ti == 0 or bool(self.ti.interaction_list)))

    def forward(self, net_output: torch.Tensor, target: torch.Tensor):
        """target must be (b, 1, *spatial)."""
        if self._fusable(net_output, target):
            ti = self.ti
            return ops.seg_loss(net_output, target, self.weight_ce, self.weight_dice, self.weight_ti, self.dc.batch_dice,
                                self.dc.do_bg, self.dc.smooth, getattr(self.dc, "ddp", False),
                                ti.interaction_table() if self.weight_ti != 0 else None, ti.connectivity, ti.min_thick)
        if self.ignore_label is not None:
            assert target.shape[1] == 1, 'ignore label is not implemented for one hot encoded target variables ' \
                                         '(DC_and_CE_loss)'
            mask = (target != self.ignore_label).bool()
            target_dice = torch.clone(target)
            target_dice[target == self.ignore_label] = 0
            num_fg = mask.sum()
        else:
            target_dice = target
            mask = None
        dc_loss = self.dc(net_output, target_dice, loss_mask=mask) if self.weight_dice != 0 else 0
        ce_loss = self.ce(net_output, target[:, 0].long()) \
            if self.weight_ce != 0 and (self.ignore_label is None or num_fg > 0) else 0
        ti_loss = self.ti(net_output, target) if self.weight_ti != 0 else 0
        return self.weight_ce * ce_loss + self.weight_dice * dc_loss + self.weight_ti * ti_loss


class DC_and_CE_and_BTI_Loss(_CompoundInteractionLoss):
    _ti_class = BTI_Loss


class DC_and_CE_and_TI_Loss(_CompoundInteractionLoss):
    _ti_class = TI_Loss


class DeepSupervisionWrapper(nn.Module):
    """sum_i w_i * loss(out_i, target_i), skipping zero weights (upstream nnunetv2.training.loss.deep_supervision;
    used by every reference trainer, …_BTI_Synapse.py:63).  Local copy for standalone use."""

    def __init__(self, loss, weight_factors=None):
        super().__init__()
        assert any([x != 0 for x in weight_factors]), "At least one weight factor should be != 0.0"
        self.weight_factors = tuple(weight_factors)
        self.loss = loss

    def forward(self, *args):
        assert all([isinstance(i, (tuple, list)) for i in args]), \
            f"all args must be either tuple or list, got {[type(i) for i in args]}"
        weights = self.weight_factors if self.weight_factors is not None else (1,) * len(args[0])
        return sum([weights[i] * self.loss(*inputs) for i, inputs in enumerate(zip(*args)) if weights[i] != 0.0])
