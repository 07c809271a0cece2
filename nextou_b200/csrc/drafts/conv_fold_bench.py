"""DRAFT timing: conv_fold (standalone draft build) vs the product's conv_halo at the two dominant small-Cout shapes."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, HERE)
import torch

from conv_fold import conv_fold, pack_fold_weight
from nextou_b200 import ops

lib = ctypes.CDLL(sys.argv[1])
lib.nextou_last_error.restype = ctypes.c_char_p


def t(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for cin, cout, ks, sp in [(33, 33, (1, 3, 3), (64, 224, 192)), (66, 66, (3, 3, 3), (64, 112, 96)), (66, 33, (1, 3, 3), (64, 224, 192)),
                          (132, 66, (3, 3, 3), (64, 112, 96))]:
    V = sp[0] * sp[1] * sp[2]
    x = torch.randn(V, ops.pad8(cin), device="cuda").bfloat16()[:, :cin]
    w = torch.randn(cout, cin, *ks, device="cuda") * 0.05
    wf, wp = pack_fold_weight(w), ops.pack_conv_weight(w)
    f = t(lambda: conv_fold(lib, x, 1, sp, cin, wf, cout, ks[0]))
    h = t(lambda: ops.conv_ndhwc_bf16(x, 1, sp, cin, wp, cout, ks, None, halo=True))
    print(f"{cin}->{cout} {ks} {sp}: fold {f:.3f} ms, halo {h:.3f} ms, speed-up {h / f:.2f}x", flush=True)
