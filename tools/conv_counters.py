"""Per-role cycle split of the halo conv kernel (CTA 0) on representative layer shapes (diagnostic)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nextou_b200 import ops, _lib
dev = "cuda"
L = _lib.lib()
buf = torch.zeros(16, dtype=torch.int64, device=dev)
L.nextou_debug_set_conv_counters(ctypes.c_void_p(buf.data_ptr()))
names = ["mma total", "mma wait tmem_empty", "mma wait fullA", "mma wait fullB", "tiles", "epi total", "epi wait tmem_full"]
for name, sp, cin, cout, ks in [("enc s0 conv1", (64, 224, 192), 33, 33, (1, 3, 3)), ("enc s1 conv1", (64, 112, 96), 66, 66, (3, 3, 3)),
                                ("dec st3 conv0", (64, 112, 96), 132, 66, (3, 3, 3))]:
    V = sp[0] * sp[1] * sp[2]
    x = torch.randn(V, ops.pad8(cin), device=dev).bfloat16()[:, :cin]
    w = torch.randn(cout, cin, *ks, device=dev) * 0.05
    wp = ops.pack_conv_weight(w)
    for _ in range(2):
        buf.zero_()
        ops.conv_ndhwc_bf16(x, 1, sp, cin, wp, cout, ks, None, halo=True)
        torch.cuda.synchronize()
    v = buf.tolist()
    print(name)
    for n, c in zip(names, v):
        print(f"   {n:24s} {c:12d}" + (f"  per tile {c / max(v[4], 1):9.0f}" if n != "tiles" else ""))
