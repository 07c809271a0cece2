"""Drop-in for nnunetv2/training/nnUNetTrainer/nnUNetTrainer_NexToU_TI_NoMirroring.py."""
from nextou_b200.trainers import nnUNetTrainer_NexToU_TI_NoMirroring  # noqa: F401
