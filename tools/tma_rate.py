import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nextou_b200 import _lib
L = _lib.lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
V = 64 * 224 * 192
print("mode cin pitch ctas  cycles/box   bytes/cycle/SM(smem)", flush=True)
for mode, cin, pitch in [(0, 33, 40), (0, 64, 64), (0, 40, 40), (1, 33, 40), (1, 64, 64), (1, 132, 136)]:
    x = torch.zeros(V, pitch, dtype=torch.bfloat16, device="cuda")
    for ctas in (1, 148):
        for _ in range(2):
            _lib.check(L.nextou_debug_tma_rate(ctypes.c_void_p(x.data_ptr()), mode, cin, pitch, 256, ctas, ctypes.c_void_p(out.data_ptr()), None), "probe")
            torch.cuda.synchronize()
        cyc = out[:ctas].float().mean().item() / 256
        bb = 180 * 128 if mode == 0 else 128 * 128
        print(f"{mode:4d} {cin:3d} {pitch:5d} {ctas:4d} {cyc:11.1f} {bb / cyc:10.1f}", flush=True)
