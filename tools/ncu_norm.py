"""Standalone launches of the normalisation kernels at 3d_fullres_nextou shapes for ncu (raw C-ABI calls, bf16)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nextou_b200 import _lib, ops
from nextou_b200._lib import cf, cstream, ll, ptr

L = _lib.lib()
DEV = "cuda"
shapes = [(40, 2752512), (136, 86016)]
for C, rows in shapes:
    x = (torch.randn(rows, C, device=DEV) * 2 + 0.5).bfloat16()
    dy = torch.randn(rows, C, device=DEV).bfloat16()
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    partial = ops._norm_partial(C, rows, 1, x.device)
    mean, invstd = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    y, dx = torch.empty_like(x), torch.empty_like(x)
    sums, dxs = torch.empty(2 * C, device=DEV), torch.empty(C, device=DEV)
    for _ in range(1):
        ops.check(L.nextou_norm_stats_tracked(ptr(x), 1, C, C, ll(rows), 1, cf(1e-5), ptr(partial), ptr(mean), ptr(invstd), ptr(None),
                                              ptr(None), cf(0.1), ptr(None), cstream()), "stats")
        ops.check(L.nextou_norm_apply_res(ptr(x), 1, C, C, ll(rows), 1, ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), cf(0.01),
                                          ptr(None), ptr(y), cstream()), "apply")
        ops.check(L.nextou_norm_bwd_colsum(ptr(x), ptr(dy), 1, C, C, ll(rows), 1, ptr(mean), ptr(invstd), ptr(gamma), ptr(beta),
                                           cf(0.01), ptr(partial), ptr(sums), ptr(dx), ptr(dxs), cstream()), "bwd")
    torch.cuda.synchronize()
