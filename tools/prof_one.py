"""Run each tcgen05 kernel once on two representative layer shapes (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nextou_b200 import ops
dev = "cuda"
for name, sp, cin, cout, ks in [("enc s0 conv1", (64, 224, 192), 33, 33, (1, 3, 3)), ("dec st3 conv0", (64, 112, 96), 132, 66, (3, 3, 3))]:
    V = sp[0] * sp[1] * sp[2]
    x = torch.randn(V, ops.pad8(cin), device=dev).bfloat16()[:, :cin]
    dy = torch.randn(V, ops.pad8(cout), device=dev).bfloat16()[:, :cout]
    w = torch.randn(cout, cin, *ks, device=dev) * 0.05
    wp = ops.pack_conv_weight(w)
    for _ in range(2):
        ops.conv_ndhwc_bf16(x, 1, sp, cin, wp, cout, ks, None, halo=True)
        ops.conv_wgrad_bf16(dy, x, 1, sp, cin, cout, ks, halo=True)
    torch.cuda.synchronize()
a = torch.randn(86016, 136, device=dev).bfloat16()[:, :132]
b = torch.randn(528, 136, device=dev).bfloat16()[:, :132]
for _ in range(2):
    ops.gemm_bf16_tn(a, b, None)
torch.cuda.synchronize()
