"""Drop-in for nnunetv2/training/loss/ti_loss.py."""
from nextou_b200.losses import TI_Loss  # noqa: F401
