"""Drop-in for nnunetv2/training/loss/compound_ti_loss.py."""
from nextou_b200.losses import DC_and_CE_and_TI_Loss  # noqa: F401
