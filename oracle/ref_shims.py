"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference files from /root/reference, or from the byte-identical
archive that oracle/stage_ref.py stages under oracle/_ref/ (git-ignored; it travels to the GPU box, where /root/reference
does not exist, so that `bench.py --impl reference` times the reference's own code there).

It is used by `oracle/make_golden.py` (to generate the committed fixtures under
tests/golden/), by the `-m "not gpu"` tests that pin the `oracle/` restatement
against the real reference, and by bench.py's reference arm.  Product code (nextou_b200/) never imports this.

The reference is an overlay on nnU-Net v2 and imports un-vendored packages
(`nnunetv2`, `dynamic_network_architectures`, `timm`).  We register stand-ins for
those in `sys.modules` and then load the reference files by path under their
dotted `nnunetv2.…` names.  The stand-ins for
`dynamic_network_architectures` are a restatement of that package's *public
behaviour* (its source is not in /root/reference): parity unpinned for that piece.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
from torch import nn

def _resolve_root():
    env = os.environ.get("NEXTOU_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/network_architecture/NexToU.py"):
        return "/root/reference"
    from . import stage_ref
    return stage_ref.unpack()          # byte-identical archive staged by oracle/stage_ref.py (hash-checked), or None


REF_ROOT = _resolve_root()


def reference_available() -> bool:
    return REF_ROOT is not None and os.path.isfile(os.path.join(REF_ROOT, "network_architecture", "NexToU.py"))


# --------------------------------------------------------------------------------------
# stand-ins for dynamic_network_architectures (public-API restatement)
# --------------------------------------------------------------------------------------
def _dim_of(conv_op):
    return {nn.Conv1d: 1, nn.Conv2d: 2, nn.Conv3d: 3}[conv_op]


def _conv_of(dim):
    return {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[dim]


def _scalar_to_list(conv_op, s):
    if isinstance(s, (tuple, list)):
        return s
    return [s] * _dim_of(conv_op)


def _convtransp(conv_op):
    return {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}[_dim_of(conv_op)]


def _batchnorm(conv_op):
    return {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[_dim_of(conv_op)]


def _pool(conv_op, adaptive=False, pool_type="avg"):
    d = _dim_of(conv_op)
    table = {("avg", 1): nn.AvgPool1d, ("avg", 2): nn.AvgPool2d, ("avg", 3): nn.AvgPool3d,
             ("max", 1): nn.MaxPool1d, ("max", 2): nn.MaxPool2d, ("max", 3): nn.MaxPool3d}
    return table[(pool_type, d)]


class _ConvDropoutNormReLU(nn.Module):
    def __init__(self, conv_op, cin, cout, kernel_size, stride, conv_bias=False, norm_op=None,
                 norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None,
                 nonlin_kwargs=None, nonlin_first=False):
        super().__init__()
        self.input_channels, self.output_channels = cin, cout
        stride = _scalar_to_list(conv_op, stride)
        self.stride = stride
        kernel_size = _scalar_to_list(conv_op, kernel_size)
        ops = []
        self.conv = conv_op(cin, cout, kernel_size, stride, padding=[(i - 1) // 2 for i in kernel_size],
                            dilation=1, bias=conv_bias)
        ops.append(self.conv)
        if dropout_op is not None:
            self.dropout = dropout_op(**(dropout_op_kwargs or {}))
            ops.append(self.dropout)
        if norm_op is not None:
            self.norm = norm_op(cout, **(norm_op_kwargs or {}))
            ops.append(self.norm)
        if nonlin is not None:
            self.nonlin = nonlin(**(nonlin_kwargs or {}))
            ops.append(self.nonlin)
        if nonlin_first and norm_op is not None and nonlin is not None:
            ops[-1], ops[-2] = ops[-2], ops[-1]
        self.all_modules = nn.Sequential(*ops)

    def forward(self, x):
        return self.all_modules(x)

    def compute_conv_feature_map_size(self, input_size):
        import numpy as np
        out = [i // j for i, j in zip(input_size, self.stride)]
        return np.prod([self.output_channels, *out], dtype=np.int64)


class _StackedConvBlocks(nn.Module):
    def __init__(self, num_convs, conv_op, cin, cout, kernel_size, initial_stride, conv_bias=False,
                 norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None,
                 nonlin_kwargs=None, nonlin_first=False):
        super().__init__()
        if not isinstance(cout, (tuple, list)):
            cout = [cout] * num_convs
        args = (conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs,
                nonlin_first)
        self.convs = nn.Sequential(
            _ConvDropoutNormReLU(conv_op, cin, cout[0], kernel_size, initial_stride, *args),
            *[_ConvDropoutNormReLU(conv_op, cout[i - 1], cout[i], kernel_size, 1, *args)
              for i in range(1, num_convs)])
        self.output_channels = cout[-1]
        self.initial_stride = _scalar_to_list(conv_op, initial_stride)

    def forward(self, x):
        return self.convs(x)

    def compute_conv_feature_map_size(self, input_size):
        import numpy as np
        output = self.convs[0].compute_conv_feature_map_size(input_size)
        size_after = [i // j for i, j in zip(input_size, self.initial_stride)]
        for b in self.convs[1:]:
            output += b.compute_conv_feature_map_size(size_after)
        return output


class _InitWeights_He:
    def __init__(self, neg_slope=1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


class _RobustCrossEntropyLoss(nn.CrossEntropyLoss):
    def forward(self, input, target):
        if target.ndim == input.ndim:
            assert target.shape[1] == 1
            target = target[:, 0]
        return super().forward(input, target.long())


class _SoftDiceLoss(nn.Module):
    def __init__(self, apply_nonlin=None, batch_dice=False, do_bg=True, smooth=1., ddp=True, clip_tp=None):
        super().__init__()
        self.do_bg, self.batch_dice, self.apply_nonlin, self.smooth, self.ddp = do_bg, batch_dice, apply_nonlin, smooth, ddp

    def forward(self, x, y, loss_mask=None):
        if self.apply_nonlin is not None:
            x = self.apply_nonlin(x)
        axes = tuple(range(2, x.ndim))
        with torch.no_grad():
            y_onehot = torch.zeros(x.shape, device=x.device, dtype=torch.bool)
            y_onehot.scatter_(1, y.long(), 1)
            if not self.do_bg:
                y_onehot = y_onehot[:, 1:]
            sum_gt = y_onehot.sum(axes)
        if not self.do_bg:
            x = x[:, 1:]
        intersect = (x * y_onehot).sum(axes)
        sum_pred = x.sum(axes)
        if self.batch_dice:
            intersect, sum_pred, sum_gt = intersect.sum(0), sum_pred.sum(0), sum_gt.sum(0)
        dc = (2 * intersect + self.smooth) / torch.clip(sum_gt + sum_pred + self.smooth, 1e-8)
        return -dc.mean()


class _DeepSupervisionWrapper(nn.Module):
    def __init__(self, loss, weight_factors=None):
        super().__init__()
        self.weight_factors = tuple(weight_factors)
        self.loss = loss

    def forward(self, *args):
        return sum(w * self.loss(*inputs) for w, inputs in zip(self.weight_factors, zip(*args)) if w != 0.0)


def _pkg(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # mark as package
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(_pkg(parent), child, m)
    return m


def _install_standins():
    helper = _pkg("dynamic_network_architectures.building_blocks.helper")
    helper.convert_conv_op_to_dim = _dim_of
    helper.convert_dim_to_conv_op = _conv_of
    helper.maybe_convert_scalar_to_list = _scalar_to_list
    helper.get_matching_convtransp = _convtransp
    helper.get_matching_batchnorm = _batchnorm
    helper.get_matching_pool_op = _pool
    scb = _pkg("dynamic_network_architectures.building_blocks.simple_conv_blocks")
    scb.StackedConvBlocks = _StackedConvBlocks
    scb.ConvDropoutNormReLU = _ConvDropoutNormReLU
    unet = _pkg("dynamic_network_architectures.architectures.unet")
    unet.PlainConvUNet = type("PlainConvUNet", (nn.Module,), {})
    unet.ResidualEncoderUNet = type("ResidualEncoderUNet", (nn.Module,), {})
    wi = _pkg("dynamic_network_architectures.initialization.weight_init")
    wi.InitWeights_He = _InitWeights_He
    wi.init_last_bn_before_add_to_0 = lambda m: None
    timm_layers = _pkg("timm.models.layers")
    timm_layers.DropPath = type("DropPath", (nn.Identity,), {})
    for p in ("nnunetv2", "nnunetv2.training", "nnunetv2.training.nnUNetTrainer",
              "nnunetv2.training.nnUNetTrainer.variants",
              "nnunetv2.training.nnUNetTrainer.variants.network_architecture",
              "nnunetv2.training.loss", "nnunetv2.utilities"):
        _pkg(p)
    # upstream nnU-Net loss pieces the reference's compound losses import (compound_bti_loss.py:2-5) — un-vendored, so
    # restated from their public behaviour (parity unpinned for these): softmax over dim 1, CE on (b, 1, ...)
    # targets, memory-efficient soft Dice (batch_dice / do_bg / smooth), weighted deep-supervision sum
    helpers = _pkg("nnunetv2.utilities.helpers")
    helpers.softmax_helper_dim1 = lambda x: torch.softmax(x, 1)
    rce = _pkg("nnunetv2.training.loss.robust_ce_loss")
    rce.RobustCrossEntropyLoss = _RobustCrossEntropyLoss
    dice = _pkg("nnunetv2.training.loss.dice")
    dice.SoftDiceLoss = _SoftDiceLoss
    dice.MemoryEfficientSoftDiceLoss = _SoftDiceLoss
    ds = _pkg("nnunetv2.training.loss.deep_supervision")
    ds.DeepSupervisionWrapper = _DeepSupervisionWrapper


def _load(dotted, relpath):
    if dotted in sys.modules and getattr(sys.modules[dotted], "__file__", None):
        return sys.modules[dotted]
    spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    parent, child = dotted.rsplit(".", 1)
    setattr(_pkg(parent), child, mod)
    spec.loader.exec_module(mod)
    return mod


class Ref:
    """Namespace holding the loaded reference modules."""


_REF = None


def load_reference() -> Ref:
    """Load the reference modules (cached).  Raises if /root/reference is absent."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT} (stage it with `python -m oracle.stage_ref` where /root/reference exists)")
    _install_standins()
    na = "nnunetv2.training.nnUNetTrainer.variants.network_architecture."
    r = Ref()
    r.torch_nn = _load(na + "torch_nn", "network_architecture/torch_nn.py")
    r.torch_edge = _load(na + "torch_edge", "network_architecture/torch_edge.py")
    r.pos_embed = _load(na + "pos_embed", "network_architecture/pos_embed.py")
    r.ED = _load(na + "NexToU_Encoder_Decoder", "network_architecture/NexToU_Encoder_Decoder.py")
    r.NX = _load(na + "NexToU", "network_architecture/NexToU.py")
    r.bti = _load("nnunetv2.training.loss.bti_loss", "loss/bti_loss.py")
    r.ti = _load("nnunetv2.training.loss.ti_loss", "loss/ti_loss.py")
    r.compound_bti = _load("nnunetv2.training.loss.compound_bti_loss", "loss/compound_bti_loss.py")
    r.compound_ti = _load("nnunetv2.training.loss.compound_ti_loss", "loss/compound_ti_loss.py")
    r.DeepSupervisionWrapper = _DeepSupervisionWrapper
    r.SoftDiceLoss = _SoftDiceLoss
    r.InitWeights_He = _InitWeights_He
    _REF = r
    return r


# ---- canonical configs ------------------------------------------------
SYNAPSE_EXCLUSION = [[[1, 3, 5, 7, 8, 11, 13], [2, 4, 6, 9, 10, 12]], [[1, 3, 11, 13], [5, 7, 8]],
                     [[1, 3], [11, 13]], [1, 3], [11, 13], [[5, 8], [7]], [5, 8],
                     [[4, 6, 10], [2, 9, 12]], [[4, 6], [10]], [4, 6], [[9, 12], [2]], [9, 12]]


def make_tensors(lists):
    """Same nesting rule as the trainers' make_tensors (…_BTI_Synapse.py:9-15)."""
    if not lists:
        return lists
    if isinstance(lists[0], list):
        return [make_tensors(s) for s in lists]
    return torch.tensor(lists)


def build_ref_3d(patch=(64, 224, 192), feats=(33, 66, 132, 264, 324, 324), num_classes=14, in_ch=1,
                 deep_supervision=True, seed=0):
    r = load_reference()
    n = len(feats)
    ks = [[1, 3, 3]] + [[3, 3, 3]] * (n - 1)
    st = [[1, 1, 1], [1, 2, 2]] + [[2, 2, 2]] * (n - 2)
    torch.manual_seed(seed)
    m = r.NX.NexToU(in_ch, list(patch), n, list(feats), nn.Conv3d, ks, st, 2, num_classes, 2, conv_bias=True,
                    norm_op=nn.BatchNorm3d, norm_op_kwargs={"eps": 1e-5, "affine": True},
                    nonlin=nn.LeakyReLU, nonlin_kwargs={"inplace": True}, deep_supervision=deep_supervision)
    m.apply(_InitWeights_He(1e-2))
    return m


def build_ref_2d(patch=(256, 256), feats=(33, 66, 132, 264, 512, 512, 512), num_classes=3, in_ch=1,
                 deep_supervision=True, seed=0):
    r = load_reference()
    n = len(feats)
    torch.manual_seed(seed)
    m = r.NX.NexToU(in_ch, list(patch), n, list(feats), nn.Conv2d, [[3, 3]] * n, [[1, 1]] + [[2, 2]] * (n - 1),
                    2, num_classes, 2, conv_bias=True, norm_op=nn.BatchNorm2d,
                    norm_op_kwargs={"eps": 1e-5, "affine": True}, nonlin=nn.LeakyReLU,
                    nonlin_kwargs={"inplace": True}, deep_supervision=deep_supervision)
    m.apply(_InitWeights_He(1e-2))
    return m
