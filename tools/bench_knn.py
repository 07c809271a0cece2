"""Per-site timing of the fused kNN kernel at the 3d_fullres_nextou graph shapes (diagnostic).
    python tools/bench_knn.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from nextou_b200 import ops

SITES = [  # name, graphs, N, M, C, k
    ("Pool s2", 1, 10752, 168, 132, 14), ("Swin s2", 512, 168, 168, 132, 7), ("Pool s3", 1, 10752, 1344, 264, 28),
    ("Swin s3", 64, 168, 168, 264, 14), ("Pool s4", 1, 1344, 1344, 324, 32), ("Swin s4", 8, 168, 168, 324, 14),
    ("Pool s5", 1, 168, 168, 324, 32), ("Swin s5", 1, 168, 168, 324, 28),
]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"{'site':10s} {'B':>4s} {'N':>6s} {'M':>5s} {'C':>4s} {'k':>3s} {'norm ms':>8s} {'topk ms':>8s} {'TF/s':>6s} {'Mvox/s':>8s}")
tot = 0.0
for name, B, N, M, C, k in SITES:
    x = torch.randn(B * N, C, device="cuda")
    y = torch.randn(B * M, C, device="cuda") if M != N else None
    rp = torch.randn(1, N, M, device="cuda") * 0.1
    xn, sqx = ops.knn_normalize(x, B, N)
    yn, sqy = (ops.knn_normalize(y, B, M) if y is not None else (None, None))
    tn = timeit(lambda: ops.knn_normalize(x, B, N))
    tk = timeit(lambda: ops.knn_topk(xn, sqx, yn, sqy, rp, k, 1))
    tot += tk + tn
    print(f"{name:10s} {B:4d} {N:6d} {M:5d} {C:4d} {k:3d} {tn:8.3f} {tk:8.3f} {2.0 * B * N * M * C / tk / 1e9:6.1f} {B * N / (tk + tn) / 1e3:8.1f}")
print(f"sum over the 8 site shapes: {tot:.3f} ms  (one forward = enc + dec = 14 sites)")
