"""GPU time of the first convolution (1 -> 33, 1x3x3, 64x224x192): CUDA-core kernels (csrc/conv_small.cu) vs the tensor-core
halo kernels with Cin padded to one 64-channel slab.  Raw C-ABI calls in a CUDA graph of 20 repetitions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nextou_b200 import _lib, ops
from tools.norm_ab import graph_time

DEV = "cuda"
B, sp, cin, cout, ks = 1, (64, 224, 192), 1, 33, (1, 3, 3)
V = sp[0] * sp[1] * sp[2]
x1 = torch.randn(V, 1, device=DEV).bfloat16()
x8 = torch.zeros(V, 8, device=DEV, dtype=torch.bfloat16)
x8[:, :1] = x1
w = torch.randn(cout, cin, *ks, device=DEV) / 3
bias = torch.randn(cout, device=DEV)
dy = torch.randn(V, 40, device=DEV).bfloat16()[:, :cout]
wp, _ = ops.pack_weight_pair(w, conv=True, flip_b=True)
mb = (V * 40 * 2 + V * 2) / 1e6
t = graph_time(lambda: ops.conv_small_fwd(x1, B, sp, w, bias))
print(f"small fwd   {t:7.1f} us  ({mb / t * 1e3 / 1e3:5.2f} TB/s algorithmic)")
t = graph_time(lambda: ops.conv_ndhwc_bf16(x8, B, sp, cin, wp, cout, ks, bias))
print(f"halo  fwd   {t:7.1f} us")
t = graph_time(lambda: ops.conv_small_wgrad(dy, x1, B, sp, cin, cout, ks))
print(f"small wgrad {t:7.1f} us  ({mb / t * 1e3 / 1e3:5.2f} TB/s algorithmic)")
t = graph_time(lambda: ops.conv_wgrad_bf16(dy, x8[:, :1], B, sp, cin, cout, ks))
print(f"halo  wgrad {t:7.1f} us")
