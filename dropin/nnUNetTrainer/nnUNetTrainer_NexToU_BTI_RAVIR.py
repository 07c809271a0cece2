"""Drop-in for nnunetv2/training/nnUNetTrainer/nnUNetTrainer_NexToU_BTI_RAVIR.py."""
from nextou_b200.trainers import nnUNetTrainer_NexToU_BTI_RAVIR  # noqa: F401
