"""Drop-in for nnunetv2/training/nnUNetTrainer/nnUNetTrainer_NexToU.py."""
from nextou_b200.trainers import nnUNetTrainer_NexToU  # noqa: F401
