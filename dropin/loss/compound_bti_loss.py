"""Drop-in for nnunetv2/training/loss/compound_bti_loss.py."""
from nextou_b200.losses import DC_and_CE_and_BTI_Loss  # noqa: F401
