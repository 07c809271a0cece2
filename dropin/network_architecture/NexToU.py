"""Drop-in for nnunetv2/training/nnUNetTrainer/variants/network_architecture/NexToU.py."""
from nextou_b200.model import NexToU  # noqa: F401
