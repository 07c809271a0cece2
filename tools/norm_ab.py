"""GPU time of the normalisation kernel groups (raw C-ABI calls captured into a CUDA graph of 20 repetitions, so no host
overhead is in the numbers): forward = stats + finalize + apply, backward = reduce + finalize + apply + bias-gradient
finalize, with the achieved fraction of the measured HBM peak for their streaming passes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nextou_b200 import _lib, ops
from nextou_b200._lib import cf, cstream, dtype_code, ll, ptr

DEV = "cuda"
L = _lib.lib()
REPS = 20
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def graph_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / REPS * 1e3      # us per call


shapes = [] if __name__ != "__main__" else [(40, 2752512), (72, 688128), (528, 86016), (264, 86016), (136, 86016), (264, 10752), (1056, 10752), (328, 1344), (1296, 168)]
for C, rows in shapes:
    x = (torch.randn(rows, C, device=DEV) * 2 + 0.5).bfloat16()
    dy = torch.randn(rows, C, device=DEV).bfloat16()
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    partial = ops._norm_partial(C, rows, 1, x.device)
    mean, invstd = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    y, dx = torch.empty_like(x), torch.empty_like(x)
    sums, dxs = torch.empty(2 * C, device=DEV), torch.empty(C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)

    def fwd():
        ops.check(L.nextou_norm_stats_tracked(ptr(x), 1, C, C, ll(rows), 1, cf(1e-5), ptr(partial), ptr(mean), ptr(invstd), ptr(rm),
                                              ptr(rv), cf(0.1), ptr(None), cstream()), "stats")
        ops.check(L.nextou_norm_apply_res(ptr(x), 1, C, C, ll(rows), 1, ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), cf(0.01),
                                          ptr(None), ptr(y), cstream()), "apply")

    def bwd():
        ops.check(L.nextou_norm_bwd_colsum(ptr(x), ptr(dy), 1, C, C, ll(rows), 1, ptr(mean), ptr(invstd), ptr(gamma), ptr(beta),
                                           cf(0.01), ptr(partial), ptr(sums), ptr(dx), ptr(dxs), cstream()), "bwd")
    mb = x.numel() * 2 / 1e6
    tf, tb = graph_time(fwd), graph_time(bwd)
    print(f"C={C:5d} rows={rows:8d} {mb:6.1f} MB   fwd {tf:6.1f} us = {3 * mb / tf * 1e3 / PEAK:4.2f} of the measured HBM peak (3 passes)   "
          f"bwd {tb:6.1f} us = {5 * mb / tb * 1e3 / PEAK:4.2f} (5 passes)", flush=True)
