import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nextou_b200 import _lib
L = _lib.lib()
buf = torch.zeros(4, dtype=torch.int64, device="cuda")
print("flags: 1 = 4 warps streaming tcgen05.ld, 2 = 4 warps spinning on try_wait, 4 = halo-style shifted A rows", flush=True)
print("   N flags  reps  cyc per MMA (issue)  cyc per MMA (complete)", flush=True)
for n in (48, 80, 144):
    for flags in (0, 4, 1, 2, 5):
        reps = 512
        for _ in range(2):
            _lib.check(L.nextou_debug_mma_rate(n, reps, flags, ctypes.c_void_p(buf.data_ptr()), None), "probe")
            torch.cuda.synchronize()
        a, b, _, _ = buf.tolist()
        print(f"{n:4d} {flags:5d} {reps:5d} {a / reps:14.1f} {b / reps:20.1f}", flush=True)
