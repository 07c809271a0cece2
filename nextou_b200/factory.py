"""Named network configurations and the constructor call nnUNetTrainer_NexToU.build_network_architecture makes
(reference nnUNetTrainer/nnUNetTrainer_NexToU.py:52-58, 74-91): conv_bias=True, BatchNorm(eps 1e-5, affine),
LeakyReLU(inplace), 2 convs per stage, He initialisation with the LeakyReLU slope.  bench.py, __graft_entry__.smoke()
and the tests build their models through this one function."""
from __future__ import annotations

import torch
from torch import nn

# 3d_fullres_nextou of the reference plans (nnUNetPlans.json:426-435): BASELINE.json config 2
FULL3D = dict(patch=(64, 224, 192), feats=(33, 66, 132, 264, 324, 324), num_classes=14,
              strides=[[1, 1, 1], [1, 2, 2]] + [[2, 2, 2]] * 4, kernels=[[1, 3, 3]] + [[3, 3, 3]] * 5)
# scaled-down twins with the same stage structure (GNN blocks from stage 2 on) for parity tests and smoke()
MINI3D = dict(patch=(32, 96, 128), feats=(6, 12, 24, 36, 48, 48), num_classes=5,
              strides=[[1, 1, 1], [1, 2, 2]] + [[2, 2, 2]] * 4, kernels=[[1, 3, 3]] + [[3, 3, 3]] * 5)
MINI2D = dict(patch=(64, 64), feats=(8, 16, 32, 32, 32), num_classes=3,
              strides=[[1, 1]] + [[2, 2]] * 4, kernels=[[3, 3]] * 5)


def build_nextou(cfg, deep_supervision: bool = True, in_ch: int = 1, seed: int = 0):
    """nextou_b200.NexToU with the kwargs the reference trainer passes (TR:52-58, 74-87), He-initialised (TR:88)."""
    from .conv_blocks import InitWeights_He
    from .model import NexToU
    dim = len(cfg["patch"])
    conv = nn.Conv3d if dim == 3 else nn.Conv2d
    bn = nn.BatchNorm3d if dim == 3 else nn.BatchNorm2d
    torch.manual_seed(seed)
    m = NexToU(in_ch, list(cfg["patch"]), len(cfg["feats"]), list(cfg["feats"]), conv, cfg["kernels"], cfg["strides"], 2,
               cfg["num_classes"], 2, conv_bias=True, norm_op=bn, norm_op_kwargs={"eps": 1e-5, "affine": True},
               nonlin=nn.LeakyReLU, nonlin_kwargs={"inplace": True}, deep_supervision=deep_supervision)
    m.apply(InitWeights_He(1e-2))
    return m
