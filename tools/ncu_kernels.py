"""A few representative launches of the hot kernels at 3d_fullres_nextou shapes, for `ncu --set full` (diagnostic).
    ncu --set full --import-source on --clock-control none -o gpurun_out/kernels python tools/ncu_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from nextou_b200 import ops

dev = "cuda"
which = set(sys.argv[1:]) or {"conv", "wgrad_halo", "wgrad", "gemm", "knn", "norm"}


def vol(sp, c):
    V = sp[0] * sp[1] * sp[2]
    return torch.randn(V, ops.pad8(c), device=dev).bfloat16()[:, :c]


torch.cuda.profiler.stop()
sp0, sp1 = (64, 224, 192), (64, 112, 96)
x33, d33 = vol(sp0, 33), vol(sp0, 33)
w33 = torch.randn(33, 33, 1, 3, 3, device=dev) * 0.05
x66, d66 = vol(sp1, 66), vol(sp1, 66)
w66 = torch.randn(66, 66, 3, 3, 3, device=dev) * 0.05
T = 86016
a132 = torch.randn(T, 136, device=dev).bfloat16()[:, :132]
dy528 = torch.randn(T, 528, device=dev).bfloat16()
b528 = torch.randn(528, 136, device=dev).bfloat16()[:, :132]
xs = torch.randn(512 * 168, 132, device=dev)
rp = torch.randn(1, 168, 168, device=dev) * 0.1
torch.cuda.synchronize()
torch.cuda.profiler.start()
if "conv" in which:
    ops.conv_ndhwc_bf16(x33, 1, sp0, 33, ops.pack_conv_weight(w33), 33, (1, 3, 3), None, halo=True)
    ops.conv_ndhwc_bf16(x66, 1, sp1, 66, ops.pack_conv_weight(w66), 66, (3, 3, 3), None, halo=True)
if "wgrad_halo" in which:
    ops.conv_wgrad_bf16(d33, x33, 1, sp0, 33, 33, (1, 3, 3), halo=True)
    ops.conv_wgrad_bf16(d66, x66, 1, sp1, 66, 66, (3, 3, 3), halo=True)
if "wgrad" in which:
    ops.conv_wgrad_bf16(dy528, a132, 1, (T,), 132, 528, (1,))
if "gemm" in which:
    ops.gemm_bf16_tn(a132, b528, None)
if "knn" in which:
    ops.knn_graph(xs, 512, 168, relpos=rp, k=7)
if "knn_pool" in which:      # Pool-GNN stage 3: 10 752 query tokens x 1 344 pooled candidates, 264 channels, k = 28
    ops.knn_graph(torch.randn(10752, 264, device=dev), 1, 10752, y_tok=torch.randn(1344, 264, device=dev), m=1344,
                  relpos=torch.randn(1, 10752, 1344, device=dev) * 0.1, k=28)
if "upconv" in which:        # last decoder up-sampling: ConvTranspose3d(66 -> 33, k = s = (1, 2, 2)) from 64x112x96 to 64x224x192
    wt = torch.randn(66, 33, 1, 2, 2, device=dev) * 0.05
    _, wb = ops.pack_weight_pair(wt, conv=True, flip_b=False)
    cat = torch.empty(sp0[0] * sp0[1] * sp0[2], 80, device=dev, dtype=torch.bfloat16)
    ops.convtranspose_fwd_bf16(x66, 1, sp1, 66, wb, 33, (1, 2, 2), torch.zeros(33, device=dev), out=cat, store_cols=40)
if "seghead" in which:       # full-resolution segmentation head 33 -> 14 and its data gradient 14 -> 33
    ws = torch.randn(14, 40, device=dev).bfloat16()[:, :33]
    ops.gemm_bf16_tn(x33, ws, torch.zeros(14, device=dev))
    dl = torch.randn(sp0[0] * sp0[1] * sp0[2], 16, device=dev).bfloat16()[:, :14]
    wsT = torch.randn(33, 16, device=dev).bfloat16()[:, :14]
    ops.gemm_bf16_tn(dl, wsT, None)
if "norm" in which:
    g = torch.ones(136, device=dev)
    ops.norm_act_tokens(a132, g[:132], g[:132], None, None, 0.1, 1e-5, 0.01, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
