"""Drop-in for nnunetv2/training/nnUNetTrainer/nnUNetTrainer_NexToU_BTI_Synapse.py."""
from nextou_b200.trainers import nnUNetTrainer_NexToU_BTI_Synapse  # noqa: F401
