"""Shared test helpers (test infrastructure; imports the oracle, never imported by nextou_b200)."""
from __future__ import annotations

import os

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from nextou_b200.factory import FULL3D, MINI2D, MINI3D, build_nextou  # noqa: E402  (configs + constructor live in the package)


def build_product(cfg, deep_supervision=True, in_ch=1, seed=0):
    """nextou_b200.NexToU with the kwargs nnUNetTrainer_NexToU.build_network_architecture passes (TR:52-58, 74-87)."""
    return build_nextou(cfg, deep_supervision, in_ch, seed)


def golden_model(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_state_dict(npz):
    return {k[3:]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith("sd/")}


def golden_knn_list(npz):
    n = len([k for k in npz.files if k.startswith("knn/")])
    return [torch.from_numpy(npz[f"knn/{i}"].astype(np.int64)) for i in range(n)]


def load_golden_into(model, npz):
    """Copy the golden (reference-generated) parameters into a product model; relative_pos tables are rebuilt by the
    product's own init and are checked separately."""
    sd = golden_state_dict(npz)
    own = model.state_dict()
    missing = [k for k in sd if k not in own]
    assert not missing, f"golden keys missing from product state_dict: {missing[:5]}"
    for k, v in sd.items():
        assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
    model.load_state_dict(sd, strict=False)
    return sd


def full_state_dict_for_oracle(model):
    """state_dict (incl. relative_pos) as plain CPU fp32 tensors for oracle.torch_oracle.nextou_forward."""
    return {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}


def knn_inputs(case):
    """Same deterministic inputs as oracle/make_golden.py::knn_inputs."""
    name, B, N, M, C, k, d, rp = case
    g = torch.Generator().manual_seed(1234 + sum(map(ord, name)))
    x = torch.randn(B, N, C, generator=g)
    y = torch.randn(B, M, C, generator=g) if M else None
    relpos = 0.2 * torch.randn(1, N, M or N, generator=g) if rp else None
    return x, y, relpos


KNN_CASES = [
    ("swin_s2_like", 6, 168, 0, 132, 7, 1, True),
    ("pool_s2_like", 1, 1344, 168, 132, 14, 1, True),
    ("pool_s3_like", 1, 1536, 192, 264, 28, 1, True),
    ("pool_s4", 1, 1344, 0, 324, 32, 1, True),
    ("swin_s5", 1, 168, 0, 324, 28, 1, True),
    ("dilated_343", 2, 343, 0, 132, 9, 2, True),
    ("norelpos_ragged", 3, 77, 50, 36, 5, 3, False),
    ("k_eq_m", 2, 16, 0, 12, 8, 2, False),
]
