"""nnU-Net v2 trainer plug-ins — API mirror of the reference's nnUNetTrainer/*.py.

Each class keeps the reference's name (nnU-Net discovers trainers by class name, README.md:81-92), hook signatures
and loss configuration; the network they build is nextou_b200.model.NexToU and the losses are nextou_b200.losses.*.
nnU-Net itself is not a dependency of this package: when `nnunetv2` is not importable the classes derive from a
small stand-in base (enough to build the network / loss outside a full nnU-Net checkout and to unit-test the hooks).
"""
from __future__ import annotations

from itertools import combinations

import numpy as np
import torch
from torch import nn

from .conv_blocks import InitWeights_He, convert_dim_to_conv_op, get_matching_batchnorm
from .losses import (DC_and_CE_and_BTI_Loss, DC_and_CE_and_TI_Loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss)
from .model import NexToU

try:  # inside nnU-Net
    from nnunetv2.training.nnUNetTrainer.nnUNetTrainer import nnUNetTrainer as _Base  # type: ignore
    from nnunetv2.training.loss.deep_supervision import DeepSupervisionWrapper  # type: ignore  # noqa: F811
    from nnunetv2.training.loss.dice import MemoryEfficientSoftDiceLoss  # type: ignore  # noqa: F811
except Exception:  # standalone
    class _Base:  # minimal stand-in of the attributes the hooks below read
        def __init__(self, plans_manager=None, configuration_manager=None, dataset_json=None, device="cuda",
                     is_ddp=False, ignore_label=None):
            self.plans_manager, self.configuration_manager, self.dataset_json = plans_manager, configuration_manager, dataset_json
            self.device, self.is_ddp = torch.device(device), is_ddp
            self.label_manager = type("LabelManager", (), {"ignore_label": ignore_label})()

        def _get_deep_supervision_scales(self):
            pool = np.vstack(self.configuration_manager.pool_op_kernel_sizes)
            return list(list(i) for i in 1 / np.cumprod(pool, axis=0))[:-1]

        def print_to_log_file(self, *args, **kwargs):
            print(*args)

        def configure_rotation_dummyDA_mirroring_and_inital_patch_size(self):
            return None, False, tuple(self.configuration_manager.patch_size), (0, 1, 2)


class nnUNetTrainer_NexToU(_Base):
    """Builds NexToU from the plans (nnUNetTrainer/nnUNetTrainer_NexToU.py:17-91)."""

    @staticmethod
    def build_network_architecture(plans_manager, dataset_json, configuration_manager, num_input_channels,
                                   enable_deep_supervision: bool = True) -> nn.Module:
        num_stages = len(configuration_manager.conv_kernel_sizes)
        dim = len(configuration_manager.conv_kernel_sizes[0])
        conv_op = convert_dim_to_conv_op(dim)
        label_manager = plans_manager.get_label_manager(dataset_json)
        model = NexToU(
            input_channels=num_input_channels,
            patch_size=configuration_manager.patch_size,
            n_stages=num_stages,
            features_per_stage=[min(configuration_manager.UNet_base_num_features * 2 ** i,
                                    configuration_manager.unet_max_num_features) for i in range(num_stages)],
            conv_op=conv_op,
            kernel_sizes=configuration_manager.conv_kernel_sizes,
            strides=configuration_manager.pool_op_kernel_sizes,
            num_classes=label_manager.num_segmentation_heads,
            deep_supervision=enable_deep_supervision,
            n_conv_per_stage=configuration_manager.n_conv_per_stage_encoder,
            n_conv_per_stage_decoder=configuration_manager.n_conv_per_stage_decoder,
            conv_bias=True,
            norm_op=get_matching_batchnorm(conv_op),
            norm_op_kwargs={'eps': 1e-5, 'affine': True},
            dropout_op=None, dropout_op_kwargs=None,
            nonlin=nn.LeakyReLU, nonlin_kwargs={'inplace': True},
        )
        model.apply(InitWeights_He(1e-2))
        return model

    # ---- B200 execution settings (no counterpart in the reference: it inherits these hooks from upstream nnU-Net) ----
    #: autocast dtype of the stock train_step / validation_step / predictor.  Upstream opens `torch.autocast("cuda")` without a
    #: dtype, i.e. fp16 + GradScaler; the tcgen05 kernels of this package take bf16 operands (fp16 inputs are widened to fp32
    #: and fall back to the library GEMMs), so the trainer switches the process-wide CUDA autocast default to bf16 — same
    #: 8-bit-exponent range as fp32, which makes the GradScaler a no-op.  Set NEXTOU_AUTOCAST=fp16 to keep upstream's dtype.
    autocast_dtype = torch.bfloat16

    def initialize(self):
        import os
        if os.environ.get("NEXTOU_AUTOCAST", "bf16").lower() not in ("fp16", "float16", "half") and torch.cuda.is_available():
            torch.set_autocast_dtype("cuda", self.autocast_dtype)
            self.print_to_log_file("nextou_b200: CUDA autocast dtype set to %s (tcgen05 tensor-core path)" % self.autocast_dtype)
        parent = getattr(super(), "initialize", None)
        if parent is not None:
            parent()

    def configure_optimizers(self):
        """Upstream: torch.optim.SGD(lr, weight_decay, momentum 0.99, nesterov) + PolyLRScheduler.  Same hyper-parameters through
        nextou_b200.optim.FusedSGD (one multi-tensor update that also refreshes the bf16 operand packs); gradient clipping stays
        in upstream's train_step (clip_grad_norm_), so max_grad_norm is left unset here."""
        from .optim import FusedSGD
        params = [p for p in self.network.parameters() if p.requires_grad]
        lr, wd = getattr(self, "initial_lr", 1e-2), getattr(self, "weight_decay", 3e-5)
        if params and all(p.is_cuda and p.dtype == torch.float32 for p in params):
            optimizer = FusedSGD(params, lr, weight_decay=wd, momentum=0.99, nesterov=True)
        else:
            optimizer = torch.optim.SGD(params, lr, weight_decay=wd, momentum=0.99, nesterov=True)
        try:
            from nnunetv2.training.lr_scheduler.polylr import PolyLRScheduler  # type: ignore
            scheduler = PolyLRScheduler(optimizer, lr, getattr(self, "num_epochs", 1000))
        except Exception:
            scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda e, n=getattr(self, "num_epochs", 1000): (1 - e / n) ** 0.9)
        return optimizer, scheduler


class nnUNetTrainer_NexToU_NoMirroring(nnUNetTrainer_NexToU):
    """No mirror augmentation / TTA (nnUNetTrainer_NexToU_NoMirroring.py:4-10)."""

    def configure_rotation_dummyDA_mirroring_and_inital_patch_size(self):
        rotation_for_DA, do_dummy_2d_data_aug, initial_patch_size, mirror_axes = \
            super().configure_rotation_dummyDA_mirroring_and_inital_patch_size()
        mirror_axes = None
        self.inference_allowed_mirroring_axes = None
        return rotation_for_DA, do_dummy_2d_data_aug, initial_patch_size, mirror_axes


class _InteractionLossMixin:
    """Shared `_build_loss`: deep-supervision weights 1, 1/2, 1/4, ... with the lowest resolution dropped, Dice + CE +
    lambda * (B)TI with lambda = 1e-6 (3-D) / 1e-4 (2-D) and 26- / 8-connectivity (…_BTI_Synapse.py:17-64)."""
    _compound = DC_and_CE_and_BTI_Loss

    def make_tensors(self, lists, device):
        if not lists:
            return lists
        if isinstance(lists[0], list):
            return [self.make_tensors(sub, device) for sub in lists]
        return torch.tensor(lists).to(device)

    def _interaction_lists(self):
        raise NotImplementedError

    def _build_loss(self):
        deep_supervision_scales = self._get_deep_supervision_scales()
        weights = np.array([1 / (2 ** i) for i in range(len(deep_supervision_scales))])
        weights[-1] = 0
        weights = weights / weights.sum()
        dim = len(self.configuration_manager.patch_size)
        connectivity, lambda_ti = (26, 1e-6) if dim == 3 else (8, 1e-4)
        inclusion_list, exclusion_list = self._interaction_lists()
        inclusion_list = self.make_tensors(inclusion_list, self.device)
        exclusion_list = self.make_tensors(exclusion_list, self.device)
        loss = self._compound(
            {'batch_dice': self.configuration_manager.batch_dice, 'smooth': 1e-5, 'do_bg': False, 'ddp': self.is_ddp}, {},
            {'dim': dim, 'connectivity': connectivity, 'inclusion': inclusion_list, 'exclusion': exclusion_list,
             'min_thick': 1},
            weight_ce=1, weight_dice=1, weight_ti=lambda_ti, ignore_label=self.label_manager.ignore_label,
            dice_class=MemoryEfficientSoftDiceLoss)
        self.print_to_log_file("dim: %s" % str(dim))
        self.print_to_log_file("connectivity: %s" % str(connectivity))
        self.print_to_log_file("lambda_ti: %s" % str(lambda_ti))
        self.print_to_log_file("inclusion_list: %s" % str(inclusion_list))
        self.print_to_log_file("exclusion_list_len: %s" % str(len(exclusion_list)))
        self.print_to_log_file("exclusion_list: %s" % str(exclusion_list))
        return DeepSupervisionWrapper(loss, weights)


class nnUNetTrainer_NexToU_BTI_Synapse(_InteractionLossMixin, nnUNetTrainer_NexToU):
    """Binary-tree exclusion sets of the Synapse / BTCV label hierarchy (…_BTI_Synapse.py:43-44)."""

    def _interaction_lists(self):
        return [], [[[1, 3, 5, 7, 8, 11, 13], [2, 4, 6, 9, 10, 12]], [[1, 3, 11, 13], [5, 7, 8]], [[1, 3], [11, 13]], [1, 3],
                    [11, 13], [[5, 8], [7]], [5, 8], [[4, 6, 10], [2, 9, 12]], [[4, 6], [10]], [4, 6], [[9, 12], [2]], [9, 12]]


class nnUNetTrainer_NexToU_BTI_ICA_NoMirroring(_InteractionLossMixin, nnUNetTrainer_NexToU_NoMirroring):
    """Intracranial-artery label tree (…_BTI_ICA_NoMirroring.py:43)."""

    def _interaction_lists(self):
        return [], [[[7, 9, 11, 12, 14, 15, 16, 17, 18], [1, 2, 3, 4, 5, 6, 8, 10, 13]], [[7, 9, 11, 12], [14, 15, 16, 17, 18]],
                    [[7, 9], [11, 12]], [7, 9], [11, 12], [[14, 15], [16, 17, 18]], [14, 15], [[16, 17], [18]], [16, 17],
                    [[3, 8, 10, 13], [1, 2, 4, 5, 6]], [[3, 10], [8, 13]], [3, 10], [8, 13], [[1, 6], [2, 4, 5]], [1, 6],
                    [[2, 4], [5]], [2, 4]]


class nnUNetTrainer_NexToU_BTI_RAVIR(_InteractionLossMixin, nnUNetTrainer_NexToU):
    """Artery / vein exclusion (…_BTI_RAVIR.py:43)."""

    def _interaction_lists(self):
        return [], [[1, 2]]


class nnUNetTrainer_NexToU_TI(_InteractionLossMixin, nnUNetTrainer_NexToU):
    """All C(n, 2) foreground label pairs exclude each other (…_TI.py:10-13, 48)."""
    _compound = DC_and_CE_and_TI_Loss

    def generate_combinations(self, n):
        return [list(comb) for comb in combinations([i + 1 for i in range(n)], 2)]

    def _interaction_lists(self):
        return [], self.generate_combinations(max(self.dataset_json["labels"].values()))


class nnUNetTrainer_NexToU_TI_NoMirroring(_InteractionLossMixin, nnUNetTrainer_NexToU_NoMirroring):
    """TI loss without mirroring (…_TI_NoMirroring.py)."""
    _compound = DC_and_CE_and_TI_Loss

    def generate_combinations(self, n):
        return [list(comb) for comb in combinations([i + 1 for i in range(n)], 2)]

    def _interaction_lists(self):
        return [], self.generate_combinations(max(self.dataset_json["labels"].values()))
