"""Drop-in for nnunetv2/training/loss/bti_loss.py."""
from nextou_b200.losses import BTI_Loss  # noqa: F401
